"""Device spike-train analyses (hhd_an_* through the reference-named modules) against the reference's own golden vectors
and the CPU oracle.  Integer counts and everything the `codes/` formulas derive from them: bit-exact.  The revision's
pandas/np.corrcoef coefficients and the ISI moments: |d| <= 1e-12 (different but exact-in-integers summation order)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import analysis_oracle as ao  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200 import analysis as an  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200.codes import AuxiliarEquations as AE  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision import AnalysesTools as AT  # noqa: E402

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(ROOT, "tests", "golden", "analysis_golden.npz"))
TRAINS = [G["train_%d" % i] for i in range(int(G["n"]))]
GROUP = [TRAINS[i] for i in (0, 1, 2, 4, 5, 6)]


def test_codes_functions_match_the_reference_bit_for_bit():
    t0, t1, t_bin, lag0, lag1 = G["codes_args"]
    assert np.array_equal(AE.set_binary_vector(TRAINS[5], t0, t1, t_bin), G["codes_binary_5"])
    lags, corr = AE.Pearson_correlation_group(GROUP, t0, t1, t_bin, lag0, lag1)
    assert np.array_equal(lags, G["codes_group_lags"]) and np.array_equal(corr, G["codes_group_corr"])
    assert np.array_equal(np.asarray(AE.Zero_lag_Pearson_correlation_group(GROUP, t0, t1, t_bin)), G["codes_zero_lag_group"])
    assert np.array_equal(AE.Pearson_correlation(TRAINS[0], TRAINS[1], t0, t1, t_bin, lag0, lag1)[1], G["codes_pair01_corr"])
    t, r = AE.pop_rate(TRAINS, 0.0, 400.0, 5.0)
    assert np.array_equal(t, G["codes_pop_rate_t"]) and np.array_equal(r, G["codes_pop_rate"])
    assert np.array_equal(AE.pop_rate(TRAINS, 0.0, 400.0, 5.0, sigma=2.0)[1], G["codes_pop_rate_sigma2"])


def test_revision_functions_match_the_reference():
    t0, t1, dt = G["rev_corr_args"]
    res = AT.get_correlations_from_spike_trains(TRAINS, (int(t0), float(t1)), idcs=list(G["rev_corr_idcs"]), delta_t=int(dt),
                                                lags=list(G["rev_corr_lags_in"]))
    for name, v in zip(("lags", "auto_mean", "auto_std", "cross_mean", "cross_std"), res):
        assert np.allclose(v, G["rev_corr_" + name], rtol=0, atol=1e-12, equal_nan=True), name
    mean, cv = AT.get_ISI_stats_from_spike_trains(TRAINS, neuron_idcs=[0, 1, 2, 4, 5, 6], tlim=(5.0, 395.0))
    assert np.allclose(mean, G["rev_isi_mean"], rtol=1e-13, atol=0, equal_nan=True)
    assert np.allclose(cv, G["rev_isi_cv"], rtol=1e-12, atol=0, equal_nan=True)
    arr = ao.binned_trains_revision(TRAINS, [0, 1, 2, 4, 5, 6], (int(t0), float(t1)), int(dt))
    direct = AT.ISIcorrelations(arr, int(dt), list(G["rev_corr_lags_in"]))
    for a, b in zip(direct, res):
        assert np.array_equal(a, b, equal_nan=True)


def test_original_model_spike_trains_as_original_isi_py_analyses_them():
    """The reference's own data and script: codes/Original_spiketime.txt through codes/Original_ISI.py:20-57 — off-grid spike
    times, 30 000 bins, 30 lags, 2256 ordered pairs; the golden values are the reference functions' output."""
    off = G["orig_offsets"]
    trains = [G["orig_trains"][off[q]:off[q + 1]] for q in range(int(G["orig_n"]))]
    lags, corr = AE.Pearson_correlation_group(trains, 1000, 61000, 2, -30, 30)
    assert np.array_equal(lags, G["orig_corr_lags"]) and np.array_equal(corr, G["orig_corr"])
    mean, cv = AT.ISIstats(trains)
    assert np.allclose(mean, G["orig_isi_mean"], rtol=1e-13, atol=0) and np.allclose(cv, G["orig_isi_cv"], rtol=1e-11, atol=0)


def _random_trains(n, t_ms, rate_hz, dt, seed, noise=False):
    rng = np.random.default_rng(seed)
    trains = []
    for i in range(n):
        k = rng.poisson(rate_hz * (1 + 3 * rng.random()) * t_ms / 1000.0) if i % 17 else 0
        steps = np.sort(rng.choice(np.arange(int(t_ms / dt)), size=min(k, int(t_ms / dt)), replace=False))
        t = steps * dt
        if noise:
            t = (steps * (dt * 1e-3)) / 1e-3                      # what `SpikeMonitor.t / ms` looks like: off-grid in the last bit
        trains.append(t)
    return trains


@pytest.mark.parametrize("n,n_ms,lags", [(150, 9000.0, [0, 1, 7, 64, 129]), (33, 700.0, [0, 3]), (1, 300.0, [0, 2])])
def test_counts_exact_at_size(n, n_ms, lags):
    trains = _random_trains(n, n_ms, 8.0, 0.05, seed=n, noise=True)
    whole = 2.5 + 1.5 * int((n_ms - 3.0) / 1.5)          # the reference's set_binary_vector needs a whole number of bins
    for rule, tlim, bin_ms in ((an.BIN_FLOORDIV, (3, n_ms - 1.7), 2), (an.BIN_TRUNC, (2.5, whole), 1.5)):
        if rule == an.BIN_FLOORDIV:
            arr = ao.binned_trains_revision(trains, list(range(n)), tlim, bin_ms)
            t1 = np.ceil(tlim[1]).astype(int)
        else:
            arr = np.stack([ao.set_binary_vector(s, tlim[0], tlim[1], bin_ms) for s in trains]).astype(int)
            t1 = tlim[1]
        neuron, t = an.spikes_from_trains(trains, list(range(n)))
        bits, bins, counts = an.binned_trains(neuron, t, np.arange(n), tlim[0], t1, bin_ms, rule, arr.shape[1], bins=True, counts=True)
        assert np.array_equal(bins, arr)
        assert np.array_equal(counts, arr.sum(axis=0))
        assert np.array_equal(bits, an.pack_bits(arr))
        use = [l for l in lags if l < arr.shape[1]]
        sxy, head, tail = an.lagged_counts(bits, arr.shape[1], use)
        osxy, ohead, otail = ao.lagged_counts(arr, use)
        assert np.array_equal(sxy, osxy) and np.array_equal(head, ohead) and np.array_equal(tail, otail)


def test_group_correlations_and_isi_at_size():
    trains = _random_trains(60, 6000.0, 6.0, 0.05, seed=5)
    lags, corr = AE.Pearson_correlation_group(trains, 100.0, 5900.0, 2.0, -20.0, 22.0)
    olags, ocorr = ao.pearson_group(trains, 100.0, 5900.0, 2.0, -20.0, 22.0)
    assert np.array_equal(lags, olags) and np.array_equal(corr, ocorr)
    idcs = [q for q in range(60) if len(trains[q]) > 2]
    res = AT.get_correlations_from_spike_trains(trains, (100, 5900.0), idcs=idcs, delta_t=2, lags=[-5, 0, 2])
    ores = ao.correlations_revision(trains, (100, 5900.0), idcs, 2, [-5, 0, 2])
    for a, b in zip(res, ores):
        assert np.allclose(a, b, rtol=0, atol=1e-12, equal_nan=True)
    mean, cv = AT.get_ISI_stats_from_spike_trains(trains, tlim=(50.0, 5990.0))
    omean, ocv = ao.isi_stats(trains, (50.0, 5990.0))
    assert np.allclose(mean, omean, rtol=1e-13, atol=0, equal_nan=True) and np.allclose(cv, ocv, rtol=1e-11, atol=0, equal_nan=True)


def test_analysis_of_a_simulated_column():
    """End to end: spikes of the engine -> device analyses, against the oracle on the same spikes."""
    from helpers import make_engine
    from oracle_binding import load_golden
    g = load_golden("net_n200_s0")
    e = make_engine("gpu", g, method="euler", accum="fixed", seed=2)
    e.run(20000)
    i, s = e.spikes()
    t = s.astype(np.float64) * 0.05
    n = e.n_neurons
    trains = [t[i == q] for q in range(n)]
    active = [q for q in range(n) if len(trains[q]) > 1]
    assert len(active) > 10
    mean, cv, count = an.isi_stats(i, t, active)
    omean, ocv = ao.isi_stats([trains[q] for q in active])
    assert np.allclose(mean, omean, rtol=1e-13) and np.allclose(cv, ocv, rtol=1e-11, equal_nan=True)
    assert np.array_equal(count, [len(trains[q]) - 1 for q in active])
    tt, rate = AE.pop_rate(trains, 0.0, 1000.0, 5.0)
    assert np.array_equal(rate, ao.pop_rate(trains, 0.0, 1000.0, 5.0)[1])
