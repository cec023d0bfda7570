"""Two independent CPU restatements of the reference's simulation loop against each other.

oracle/hhd_oracle.c   C, the reference's own units (ms, mV, pA, nS), hhd_exp / Philox from csrc/hhd_math.h (shared with the CUDA
                      engine), per-synapse queue, increments of a step summed per target first;
oracle/numpy_oracle.py vectorised numpy in SI units the way Brian's numpy target evaluates the equation strings
                      (codes/CortexNetwork.py:233-311, 343-387), numpy's exp, its own Philox, `g_post += inc` one synapse at
                      a time.  Shares no code with the C oracle or the engine.
Agreement (identical spike lists, |dV| <= 1e-9 mV) therefore checks the shared math header, every unit conversion and the C
oracle's reading of Brian's schedule; the engine is bit-identical to the C oracle (tests/test_gpu_parity.py).
The first version of this comparison found a real difference: `int(t - last_spike < 5*ms)` decides differently on
millisecond values (k*0.05 - k0*0.05 < 5.0) than on Brian's SI values for about one spike step in ten — engine and C oracle
now evaluate the clock in seconds (hhd_time_s).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from helpers import first_divergence
from numpy_oracle import NumpyOracle, philox_u01, timestep
from oracle_binding import OracleEngine, default_idc, load_golden, oracle_lib
from rescience_hass_hertag_durstewitz_2016_b200.engine import load_codes_network

STATE = ["V", "w", "g_AMPA_on", "g_AMPA_off", "g_GABA_on", "g_GABA_off", "g_NMDA_on", "g_NMDA_off"]


def both(g, method, seed=3, dt=0.05, idc=None, **kw):
    out = []
    for cls in (OracleEngine, NumpyOracle):
        e = cls(g["NeuPar"].shape[1], dt, method=method, seed=seed, **kw)
        load_codes_network(e, g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], g["source_arr"], g["target_arr"],
                           default_idc(g) if idc is None else idc)
        out.append(e)
    return out


def test_numpy_philox_equals_the_c_implementation():
    f = oracle_lib().hhdo_philox_u01
    f.restype = C.c_double
    f.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    idx = np.array([0, 1, 2, 12345, (1 << 32) + 7, (1 << 40) + 3], dtype=np.uint64)
    for seed in (0, 1, 0xDEADBEEFCAFE):
        for step in (0, 1, 99999, (1 << 33) + 5):
            for stream in (0, 1):
                ours = philox_u01(seed, idx, step, stream)
                ref = np.array([f(seed, int(i), step, stream) for i in idx])
                assert np.array_equal(ours, ref)
    t = oracle_lib().hhdo_timestep
    t.restype = C.c_int64
    t.argtypes = [C.c_double, C.c_double]
    assert timestep(5e-3, 5e-5) == t(5.0, 0.05) == 100 and timestep(5e-3, 1e-5) == t(5.0, 0.01) == 500


@pytest.mark.parametrize("method", ["euler", "rk2", "rk4"])
def test_c_and_numpy_oracles_agree_on_the_reference_column(method):
    """207-neuron column built by the reference's NetworkSetup, 200 ms: ~600 spikes, ~40 000 synaptic events with STP,
    release failures, delays, w_crossing projections and depolarisation-block episodes."""
    g = load_golden("net_n200_s0")
    c, n = both(g, method)
    for e in (c, n):
        e.run(4000)
    ci, cs = c.spikes()
    ni, ns = n.spikes()
    assert len(ci) > 300
    assert first_divergence(ci, cs, ni, ns) == len(ci) == len(ni)
    for v in STATE:
        assert np.max(np.abs(c.read_state(v) - n.read_state(v))) <= 1e-9, v
    assert np.max(np.abs(c.read_state("last_spike") - n.read_state("last_spike"))) <= 1e-9
    for v in ("u", "R", "a"):
        assert np.max(np.abs(c.read_syn_state(v) - n.read_syn_state(v))) <= 1e-11, v
    assert np.array_equal(c.read_syn_state("delay_steps"), n.read_syn_state("delay_steps"))
    cc, nc = c.counters(), n.counters()
    assert cc["syn_events"] == nc["syn_events"] > 10000 and cc["w_crossings"] == nc["w_crossings"] and cc["spikes"] == nc["spikes"]


def test_oracles_agree_on_the_default_column_and_its_lfp():
    """BASELINE config 1 (1003 neurons, 293 599 synapses, rk4), first 75 ms, with the summed-I_tot monitor."""
    g = load_golden("net_n1000_s0_full")
    c, n = both(g, "rk4", seed=0)
    mc = c.add_state_monitor("I_tot", None, "sum")
    n.add_state_monitor("I_tot", None, "sum")
    for e in (c, n):
        e.run(1500)
    ci, cs = c.spikes()
    ni, ns = n.spikes()
    assert len(ci) > 100 and first_divergence(ci, cs, ni, ns) == len(ci) == len(ni)
    assert np.max(np.abs(c.read_state("V") - n.read_state("V"))) <= 1e-9
    lc, ln = c.read_monitor(mc)[:, 0], n.read_monitor(0)[:, 0]
    assert np.allclose(lc, ln, rtol=1e-10, atol=1e-6)


def test_depolarisation_block_window_is_evaluated_in_seconds():
    """`int(t - last_spike < 5*ms)` (codes/CortexNetwork.py:251) on Brian's SI clock: for a spike at step k0 the gate at
    step k0 + 100 is open when k*dt - k0*dt rounds below 0.005 s.  Both oracles must take the same side for every k0; on
    millisecond values (k*0.05 - k0*0.05 < 5.0) a tenth of the steps land on the other side."""
    dt_ms = 0.05
    dts = dt_ms * 1e-3
    differs = sum(((k0 + 100) * dts - k0 * dts < 5 * 1e-3) != ((k0 + 100) * dt_ms - k0 * dt_ms < 5.0) for k0 in range(20000))
    assert differs > 1000       # the unit matters ...
    g = load_golden("net_n200_s0")
    idc = default_idc(g) * 2.0           # ... and strongly driven cells sit in the block term after every spike
    c, n = both(g, "euler", idc=idc)
    for e in (c, n):
        e.run(3000)
    ci, cs = c.spikes()
    ni, ns = n.spikes()
    assert len(ci) > 600 and first_divergence(ci, cs, ni, ns) == len(ci) == len(ni)
    assert np.max(np.abs(c.read_state("V") - n.read_state("V"))) <= 1e-9


def test_oracles_agree_on_the_revisions_model_variant():
    """A network drawn by the revision's own generator (codes_revision/NetworkSetup.py port, bit-identical to the reference's)
    with the revision's constants — NMDA Mg factor 1, last_spike = -1000 ms, intra-stripe delays 1 ms longer — on both oracles."""
    from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.NetworkSetup import network_setup
    net = network_setup(150, 1, None, 3)
    P = net.membr_params.values
    n = P.shape[1]
    NeuPar = np.zeros((12, n))
    NeuPar[[0, 1, 2, 3, 4, 5, 7, 8, 9]] = P
    NeuPar[10], NeuPar[11] = net.refractory_current.values, P[7]
    sp = net.syn_params
    SynPar = np.stack([sp["channel"].values[0] + 1, sp["STSP_params"].values[0], sp["STSP_params"].values[1], sp["STSP_params"].values[2],
                       sp["spiking"].values[0], sp["delay"].values[0], sp["spiking"].values[1]])
    STyp = load_golden("net_n200_s0")["STypPar"].copy()
    assert net.channel_constants()["mg_fac"] == 1.0
    STyp[5, 2] = 1.0
    idc = np.full(n, 200.0)
    idc[net.neuron_idcs(("PC", 0))] = 250.0
    g = dict(NeuPar=NeuPar, V0=np.stack([P[2], np.zeros(n)]), STypPar=STyp, SynPar=SynPar, source_arr=net.syn_pairs[1], target_arr=net.syn_pairs[0])
    c, m = both(g, "rk4", idc=idc, last_spike0_ms=-1000.0)
    for e in (c, m):
        e.run(6000)
    ci, cs = c.spikes()
    ni, ns = m.spikes()
    assert len(ci) > 200 and first_divergence(ci, cs, ni, ns) == len(ci) == len(ni)
    for v in STATE:
        assert np.max(np.abs(c.read_state(v) - m.read_state(v))) <= 1e-9, v
    assert c.counters()["syn_events"] == m.counters()["syn_events"] > 5000
