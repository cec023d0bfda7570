"""The integrator + threshold + w_crossing logic against the reference's own closed-form simpAdEx firing rate.

tests/golden/fI_golden.npz holds 610 (parameters, I, V0, w0) -> rate cases computed by the REFERENCE's `FRsimpAdEx`
(/root/reference/codes/NetworkSetup.py:444-553; generator: tests/golden/make_golden_fI.py) on 64 cells of the
reference's default column and the four parameter sets of codes/SingleNeuron.py:190-294.  A neuron with constant
current and no synapses, started at (V0, w0), must fire its first spike 1000 / rate ms later.

Tolerance (stated here, the reference states none): |t_first - 1000/rate| <= 2 dt + 3e-4 * 1000/rate for rk4 at
dt = 0.05 and 0.01 ms (spike times are quantised to dt; a spike found in step k is stamped k*dt); the first-order
methods are checked at dt = 0.01 with 1 % (their global error is O(dt)).
"""
import numpy as np
import pytest

from oracle_binding import OracleEngine, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine, load_codes_network

NOSYN = (np.zeros((7, 0)), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))


def subset(g, min_rate):
    """Cases whose closed-form rate is at least min_rate (bounds the run at fine dt)."""
    keep = g["rate_hz"] >= min_rate
    return {k: (v[..., keep] if k != "STypPar" else v) for k, v in g.items()}


def first_spike_times(cls, g, dt, method, **kw):
    P, I = g["params"], g["I"]
    m = P.shape[1]
    e = cls(m, dt, method=method, seed=0, **kw)
    load_codes_network(e, P, np.stack([g["V0"], g["w0"]]), g["STypPar"], *NOSYN, I)
    e.run(int((1000.0 / g["rate_hz"].min() * 1.01 + 5.0) / dt))
    i, s = e.spikes()
    nonfinite = e.counters()["nonfinite"]
    e.close()
    t = np.full(m, np.nan)
    first = np.unique(i, return_index=True)
    t[first[0]] = s[first[1]] * dt
    return t, nonfinite


def check(t, g, dt, rel):
    want = 1000.0 / g["rate_hz"]
    assert not np.isnan(t).any(), "%d cases never spiked" % np.isnan(t).sum()
    err = np.abs(t - want)
    tol = 2 * dt + rel * want
    worst = np.argmax(err / tol)
    assert (err <= tol).all(), "case %d (cell %d kind %d): first spike %.3f ms, closed form %.3f ms" % (
        worst, g["origin"][worst], g["kind"][worst], t[worst], want[worst])
    return float(np.max(err / want))


@pytest.mark.parametrize("dt", [0.05, 0.01])
def test_oracle_rk4_matches_reference_closed_form(dt):
    g = subset(load_golden("fI_golden"), 0.5 if dt == 0.05 else 2.0)
    t, nonfinite = first_spike_times(OracleEngine, g, dt, "rk4")
    assert nonfinite == 0
    check(t, g, dt, 3e-4)


@pytest.mark.parametrize("method", ["euler", "rk2", "gsl_rk2"])
def test_oracle_other_methods_match_reference_closed_form(method):
    g = subset(load_golden("fI_golden"), 2.0)
    t, _ = first_spike_times(OracleEngine, g, 0.01, method)
    check(t, g, 0.01, 1e-2 if method == "euler" else 1e-3)


def test_oracle_silent_below_rheobase():
    """I_SN = g_L (V_T - E_L - delta_T) (codes/CortexNetwork.py:339, content.tex:130): no spikes below it."""
    g = load_golden("fI_golden")
    P = g["params"]
    first = np.unique(g["origin"], return_index=True)[1]
    P = P[:, first]
    i_sn = P[1] * (P[9] - P[2] - P[3])
    keep = i_sn > 1.0
    P, i_sn = P[:, keep], i_sn[keep]
    for fac, spiking in ((0.98, False), (1.05, True)):
        e = OracleEngine(P.shape[1], 0.05, method="rk4", seed=0)
        load_codes_network(e, P, np.stack([P[2], np.zeros(P.shape[1])]), g["STypPar"], *NOSYN, i_sn * fac)
        e.run(100000)
        i, _ = e.spikes()
        e.close()
        assert (len(np.unique(i)) == P.shape[1]) if spiking else (len(i) == 0), (fac, len(np.unique(i)), P.shape[1])


@pytest.mark.gpu
@pytest.mark.parametrize("plan", ["graph", "resident"])
@pytest.mark.parametrize("dt,method", [(0.05, "rk4"), (0.01, "rk4"), (0.01, "rk2"), (0.01, "gsl_rk2")])
def test_gpu_matches_reference_closed_form(dt, method, plan):
    g = subset(load_golden("fI_golden"), 0.5 if dt == 0.05 else 2.0)
    t, nonfinite = first_spike_times(Engine, g, dt, method, plan=plan)
    assert nonfinite == 0
    check(t, g, dt, 3e-4 if method == "rk4" else 1e-3)
    to, _ = first_spike_times(OracleEngine, g, dt, method)
    assert np.array_equal(t, to)                  # and bit-identical spike steps with the oracle
