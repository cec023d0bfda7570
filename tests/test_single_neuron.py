"""codes/SingleNeuron.py façade (on the CPU oracle): the four hand-picked parameter sets of the reference's script against its own
closed-form onset rate (tests/golden/fI_golden.npz, kind/origin mark the SingleNeuron.py sets)."""
import numpy as np
import pytest

from oracle_binding import OracleEngine, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.codes.SingleNeuron import SingleNeuron
from rescience_hass_hertag_durstewitz_2016_b200.units import ms, mV, pA


def cases():
    g = load_golden("fI_golden")
    pick = [i for i in range(len(g["I"])) if g["origin"][i] < 0 and g["w0"][i] == 0 and g["rate_hz"][i] > 5][:6]
    return g, pick


def test_first_spike_follows_the_closed_form_rate():
    g, pick = cases()
    assert len(pick) >= 3
    for i in pick:
        n = SingleNeuron(g["params"][:, i], I=float(g["I"][i]), V0=float(g["V0"][i]), w0=0.0, method="rk4", time_step=0.05, _engine_cls=OracleEngine)
        want = 1000.0 / g["rate_hz"][i]
        n.run((want * 1.05 + 5) * ms)
        t = n.spikemonitor.t / ms
        assert len(t) >= 1 and abs(t[0] - want) <= 0.1 + 3e-4 * want
        V = n.neuronmonitor.V[0] / mV
        assert V.shape[0] == int(round((want * 1.05 + 5) / 0.05)) and V.max() <= g["params"][4, i] + 60 and V[0] == g["V0"][i]
        assert n.neuronmonitor.w.shape == (1, V.shape[0]) and n.neuronmonitor.w[0, 0] / pA == 0.0
        assert n.I_SN == pytest.approx(g["params"][1, i] * (g["params"][9, i] - g["params"][2, i] - g["params"][3, i]))
        n.close()


def test_reset_writes_last_spike_and_nullclines():
    g, pick = cases()
    i = pick[0]
    n = SingleNeuron(g["params"][:, i], I=float(g["I"][i]), _engine_cls=OracleEngine)
    n.run(2 * 1000.0 / g["rate_hz"][i] * ms + 20 * ms)
    t = n.spikemonitor.t / ms
    assert len(t) >= 1
    assert n.engine.read_state("last_spike")[0] == pytest.approx(t[-1], abs=1e-9)          # ms; the network's neuron without synapses keeps -1e8
    V, wV, lo, hi = n.nullclines()
    assert np.all(lo < wV) and np.all(wV < hi) or np.any(wV <= 0)
    assert abs(V[np.argmin(wV)] - g["params"][9, i]) < (V[1] - V[0]) * 1.01                 # the V-nullcline bottoms out at V_T
    n.close()
