"""The analysis oracle against the golden vectors the REFERENCE's own functions produced (CPU), and the ABI surface."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import analysis_oracle as ao  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "analysis_golden.npz"))
TRAINS = [G["train_%d" % i] for i in range(int(G["n"]))]
GROUP = [TRAINS[i] for i in (0, 1, 2, 4, 5, 6)]


def test_codes_binary_vector_and_pearson_bit_exact():
    t0, t1, t_bin, lag0, lag1 = G["codes_args"]
    assert np.array_equal(ao.set_binary_vector(TRAINS[5], t0, t1, t_bin), G["codes_binary_5"])
    lags, corr = ao.pearson_group(GROUP, t0, t1, t_bin, lag0, lag1)
    assert np.array_equal(lags, G["codes_group_lags"])
    assert np.array_equal(corr, G["codes_group_corr"])
    assert np.array_equal(np.asarray(ao.zero_lag_group(GROUP, t0, t1, t_bin), dtype=np.float64), G["codes_zero_lag_group"])


def test_codes_pop_rate_bit_exact():
    t, r = ao.pop_rate(TRAINS, 0.0, 400.0, 5.0)
    assert np.array_equal(t, G["codes_pop_rate_t"]) and np.array_equal(r, G["codes_pop_rate"])
    assert np.array_equal(ao.pop_rate(TRAINS, 0.0, 400.0, 5.0, sigma=2.0)[1], G["codes_pop_rate_sigma2"])


def test_revision_correlations_and_isi():
    t0, t1, dt = G["rev_corr_args"]
    res = ao.correlations_revision(TRAINS, (int(t0), float(t1)), list(G["rev_corr_idcs"]), int(dt), list(G["rev_corr_lags_in"]))
    for name, v in zip(("lags", "auto_mean", "auto_std", "cross_mean", "cross_std"), res):
        assert np.allclose(v, G["rev_corr_" + name], rtol=0, atol=1e-12, equal_nan=True), name   # pandas -> np.corrcoef
    mean, cv = ao.isi_stats(GROUP, (5.0, 395.0))
    assert np.array_equal(mean, G["rev_isi_mean"], equal_nan=True) and np.array_equal(cv, G["rev_isi_cv"], equal_nan=True)


def _orig_trains():
    off = G["orig_offsets"]
    return [G["orig_trains"][off[q]:off[q + 1]] for q in range(int(G["orig_n"]))]


def test_original_model_spike_trains_as_original_isi_py_analyses_them():
    """codes/Original_spiketime.txt (the original model's 61 s run, first 48 qualifying trains) through
    codes/Original_ISI.py:20-57: ISI mean / CV per train and Pearson_correlation_group(stset, 1000, 61000, 2, -30, 30)."""
    trains = _orig_trains()
    mean, cv = ao.isi_stats(trains)
    assert np.array_equal(mean, G["orig_isi_mean"]) and np.array_equal(cv, G["orig_isi_cv"])
    lags, corr = ao.pearson_group(trains[:10], 1000, 61000, 2, -30, 30)     # the full group takes the oracle a minute
    assert np.array_equal(lags, G["orig_corr_lags"]) and corr.shape == G["orig_corr"].shape


def test_lagged_counts_are_the_sufficient_statistics():
    """The integer counts reproduce both coefficient definitions (the host formulas the product applies to device counts)."""
    from rescience_hass_hertag_durstewitz_2016_b200 import analysis as an
    rng = np.random.default_rng(3)
    arr = (rng.random((9, 300)) < 0.08).astype(int)
    arr[4] = 0
    sxy, head, tail = ao.lagged_counts(arr, [0, 1, 5])
    for q, L in enumerate([0, 1, 5]):
        for lag in sorted({L, -L}):
            rb = an.pearson_binary(sxy[q], head[q], tail[q], arr.shape[1], lag)
            rs = an.pearson_shifted(sxy[q], head[q], tail[q], arr.shape[1], lag)
            for i in range(9):
                for j in range(9):
                    assert rb[i, j] == ao.pearson(arr[i].astype(float), arr[j].astype(float), lag)
                    want = ao.shifted_corr(arr[i], arr[j], lag)
                    assert (np.isnan(want) and np.isnan(rs[i, j])) or abs(rs[i, j] - want) < 1e-12
    assert np.array_equal(an.pack_bits(arr)[:, 0] & np.uint64(1), arr[:, 0].astype(np.uint64))


def test_abi_exports_analysis_symbols():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "rescience_hass_hertag_durstewitz_2016_b200", "csrc", "libhhd.so"))
    for name in ("hhd_an_words", "hhd_an_binned_trains", "hhd_an_lagged_counts", "hhd_an_isi_stats"):
        assert hasattr(lib, name), name
    assert lib.hhd_an_words(ctypes.c_int64(64)) == 2


def test_analyses_have_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU box")
    from rescience_hass_hertag_durstewitz_2016_b200 import analysis as an
    from rescience_hass_hertag_durstewitz_2016_b200.engine import HhdError
    from rescience_hass_hertag_durstewitz_2016_b200.codes import AuxiliarEquations as AE
    with pytest.raises(HhdError):
        an.lagged_counts(an.pack_bits(np.eye(3, 70, dtype=int)), 70, [0, 1])
    with pytest.raises(HhdError):
        AE.Zero_lag_Pearson_correlation_group([np.array([1.0, 5.0]), np.array([2.0])], 0.0, 10.0, 1.0)


def test_V_stats_and_LFP_match_the_reference():
    """codes_revision/AnalysesTools.py:88-172 (membrane-potential statistics with the spikes cut out) and :343-373 (LFP proxy and its
    periodogram), against values the reference's own functions returned for the same traces (make_golden_analysis.py)."""
    import types
    from rescience_hass_hertag_durstewitz_2016_b200 import units
    from rescience_hass_hertag_durstewitz_2016_b200.codes_revision import AnalysesTools as AT
    from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.BasicsSetup import Table
    g = G
    V, t, V_T, Itot = g["vstats_V"], g["vstats_t"], g["vstats_VT"], g["lfp_Itot"]
    membr = Table(V_T[None, :], ["param", "cell_index"], [["V_T"], np.arange(len(V_T))])
    cortex = types.SimpleNamespace(
        recorded=types.SimpleNamespace(V=types.SimpleNamespace(V=V * units.mV, t=t * units.ms), var=lambda name: Itot * units.pA,
                                       t=lambda name: t * units.ms),
        neurons=types.SimpleNamespace(N=V.shape[0]), network=types.SimpleNamespace(membr_params=membr))
    assert np.allclose(np.asarray(AT.get_V_stats(cortex)), g["vstats_all"], rtol=1e-12, atol=1e-12)
    assert np.allclose(np.asarray(AT.get_V_stats(cortex, neuron_idcs=[1, 3, 4], tlim=(20.0, 150.0))), g["vstats_sel"], rtol=1e-12, atol=1e-12)
    tl, lfp = AT.get_LFP(cortex, invert_Itot=True, population_agroupate="sum")
    assert np.allclose(tl, g["lfp_t"]) and np.allclose(lfp, g["lfp_sum_inverted"], rtol=1e-12)
    one = types.SimpleNamespace(recorded=types.SimpleNamespace(var=lambda name: Itot[0] * units.pA, t=lambda name: t * units.ms))
    f, p = AT.get_LFP_SPD(one, log=True, sigma=2)
    assert np.allclose(f, g["lfp_spd_f"], rtol=1e-9) and np.allclose(p, g["lfp_spd_p"], rtol=1e-9, atol=1e-9)
