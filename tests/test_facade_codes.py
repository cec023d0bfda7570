"""The codes/ façade (CortexNetwork look-alike) on the CPU oracle: constructor, monitors, inputs, helpers."""
import numpy as np
import pytest

from rescience_hass_hertag_durstewitz_2016_b200.engine import HhdError

from oracle_binding import OracleEngine, golden_groups, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.codes import CortexNetwork as CN
from rescience_hass_hertag_durstewitz_2016_b200.codes.CortexNetwork import CortexNetwork, ms, mV, pA, sourcetarget

NOSCALE = ([], [], [], [], [], [])


def build(g, method="rk4", dt=0.05, seed=3, spm=None, scales2=NOSCALE, **kw):
    groups = [[list(map(int, s)) for s in grp] for grp in golden_groups(g)]
    nstripes = int(g["nstripes"])
    cc = [[250, 200, 200, 200, 200, 200, 200, 250, 200, 200, 200, 200, 200, 200] for _ in range(nstripes)]
    spm = spm if spm is not None else (g["target_arr"], g["source_arr"])
    return CortexNetwork(g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], spm, groups, cc, [], scales2, False, method, dt,
                         seed, "", _engine_cls=OracleEngine, **kw)


def dense_spmtx(g):
    n = g["NeuPar"].shape[1]
    S = np.zeros((n, n, 2))
    seen = {}
    for k, (t, s) in enumerate(zip(g["target_arr"], g["source_arr"])):
        layer = 1 if (t, s) in seen else 0
        seen[(t, s)] = True
        S[t, s, layer] = k + 1
    return S


def test_sourcetarget_matches_reference_output():
    g = load_golden("net_n200_s0")
    t, s = sourcetarget(dense_spmtx(g))
    assert np.array_equal(t, g["target_arr"]) and np.array_equal(s, g["source_arr"])


def test_run_and_spikemonitor_interface():
    g = load_golden("net_n200_s0")
    c = build(g, spm=dense_spmtx(g))
    c.run(200 * ms)
    sm = c.spikemonitor
    assert sm.num_spikes > 100 and len(sm.i) == len(sm.t) == sm.num_spikes
    assert sm.i.min() >= 0 and sm.i.max() < 207
    assert np.all(np.diff(sm.t) >= 0) and sm.t.max() < 0.2
    assert (sm.t / ms).max() > 100                      # `spikemonitor.t/ms` as the scripts write it
    assert sm.count.sum() == sm.num_spikes and len(sm.count) == 207
    trains = sm.spike_trains()
    assert len(trains) == 207 and sum(len(v) for v in trains.values()) == sm.num_spikes
    assert c.group.N == 207 and c.group.t == pytest.approx(0.2)
    # same network through the engine API directly gives the same spikes
    from helpers import make_engine
    e = make_engine("cpu", g, method="rk4", seed=3)
    e.run(4000)
    i, s = e.spikes()
    assert np.array_equal(i, sm.i) and np.allclose(s * 0.05 * ms, sm.t)


def test_neuron_monitors_shapes_and_units():
    g = load_golden("net_n200_s0")
    c = build(g, method="rk2")
    c.neuron_monitors([["V", ["V"], [["all", "all"]]], ["I", ["I_tot", "w"], [["PC_L23", 0]]]])
    c.run(50 * ms)
    vm = c.neuronmonitors["V"]
    assert vm.V.shape == (207, 1000) and len(vm.t) == 1000 and np.array_equal(vm.record, np.arange(207))
    assert -0.5 < vm.V.min() < vm.V.max() < 0.0          # volts
    assert (vm.V / mV).min() > -500
    im = c.neuronmonitors["I"]
    npc = len(golden_groups(g)[0][0])
    assert im.I_tot.shape == (npc, 1000) and im.w.shape == (npc, 1000)
    assert np.allclose(im.I_tot[:, 0] / pA, 250.0)       # at t=0 only the background current flows
    assert np.array_equal(im.record, np.sort(golden_groups(g)[0][0]))


def test_group_helpers():
    g = load_golden("net_n60x5_s1")
    c = build(g)
    assert c.group_name_to_group_idc("PC") == [0, 7]
    assert c.group_idc_to_group_name(12) == "IN_CC_L5"
    allidc = c.group_name_to_neuron_idc([["all", "all"]])
    assert np.array_equal(allidc, np.arange(g["NeuPar"].shape[1]))
    pc0 = c.group_name_to_neuron_idc([["PC_L23", 0]])
    assert np.array_equal(pc0, np.sort(golden_groups(g)[0][0]))
    assert c.neuron_idc_to_group_name(int(pc0[0])) == ("PC_L23", 0)
    syn = c.group_name_to_synapses_idc([["PC_L5", 1, "PC_L23", 1]])
    assert len(syn) and np.all(np.isin(c.target_arr[syn], golden_groups(g)[7][1])) and np.all(np.isin(c.source_arr[syn], golden_groups(g)[0][1]))
    a = c.group_name_to_AMPAsynapses_idc([["all", "all", "PC", 0]])
    n = c.group_name_to_NMDAsynapses_idc([["all", "all", "PC", 0]])
    gab = c.group_name_to_GABAsynapses_idc([["all", "all", "PC", 0]])
    assert len(a) == len(n) > 0 and len(gab) == 0
    k = int(syn[0])
    q = int(c.target_arr[k])        # incoming + outgoing, as the reference counts them (codes/CortexNetwork.py:838-842)
    assert c.get_connections_count(q) == np.sum(c.target_arr == q) + np.sum(c.source_arr == q)
    assert np.array_equal(c.get_source_from_target(np.array([q, q + 1])), c.source_arr[np.isin(c.target_arr, [q, q + 1])])


def test_scales2_rescales_gmax():
    g = load_golden("net_n200_s0")
    c = build(g, scales2=([["PC", 0, 2.0]], [], [], [], [["PC_L5", 0, 0.5]], []))
    S0 = g["SynPar"]
    ampa_from_pc = (S0[0] == 1) & np.isin(g["source_arr"], np.concatenate([golden_groups(g)[0][0], golden_groups(g)[7][0]]))
    gaba_to_pc5 = (S0[0] == 2) & np.isin(g["target_arr"], golden_groups(g)[7][0])
    assert np.allclose(c.SynPar[4, ampa_from_pc], 2.0 * S0[4, ampa_from_pc])
    assert np.allclose(c.SynPar[4, gaba_to_pc5], 0.5 * S0[4, gaba_to_pc5])
    untouched = ~(ampa_from_pc | gaba_to_pc5)
    assert np.array_equal(c.SynPar[4, untouched], S0[4, untouched])


def test_regular_and_poisson_input_drive_the_targets():
    g = load_golden("net_n200_s0")
    for kind in ("regular", "poisson"):
        np.random.seed(5)
        c = build(g, dt=0.01)
        c.I_DC[:] = 0
        c.engine.set_idc(np.zeros(207))
        if kind == "regular":
            c.regular_input([[1, 1, 50, 2.0, 0.0, 5.0, 10.0, [[0, 0, 0.5]]]])
        else:
            c.poisson_input([[20, 1, 300.0, 2.0, 0.3, 5.0, 15.0, [[0, 0, 0.2]]]])
        c.neuron_monitors([["g", ["g_AMPA", "g_NMDA", "g_GABA"], [["PC_L23", 0]]]])
        c.run(20 * ms)
        gm = c.neuronmonitors["g"]
        assert gm.g_AMPA[:, :400].max() == 0 and gm.g_AMPA.max() > 0 and gm.g_NMDA.max() > 0
        hit = (gm.g_AMPA.max(axis=1) > 0).sum()
        assert 0 < hit <= gm.g_AMPA.shape[0]
        assert c.engine.counters()["ext_events"] > 0


def test_unsupported_variants_fail_loudly():
    g = load_golden("net_n200_s0")
    groups = golden_groups(g)
    with pytest.raises(HhdError):            # stochastic equations need euler, as in Brian
        CortexNetwork(g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], (g["target_arr"], g["source_arr"]), groups, [], [],
                      NOSCALE, 0.5, "rk4", 0.05, 0, "", _engine_cls=OracleEngine)
    with pytest.raises(ValueError):
        build(g, method="heun")


def test_fluctuating_current():
    g = load_golden("net_n200_s0")
    groups = [[list(map(int, s)) for s in grp] for grp in golden_groups(g)]
    AC = [["" for _ in range(14)]]
    AC[0][0] = "50*sin(2*pi*t/20)"
    AC[0][7] = "30"
    c = CortexNetwork(g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], (g["target_arr"], g["source_arr"]), groups,
                      [[0] * 14], [5, 25, AC], NOSCALE, False, "rk4", 0.05, 1, "", _engine_cls=OracleEngine)
    assert c.ta.shape == (500, 2)
    c.neuron_monitors([["I", ["I_inj"], [["all", "all"]]]])
    c.run(40 * ms)
    I = c.neuronmonitors["I"].I_inj / pA
    pc23, pc5, other = groups[0][0][0], groups[7][0][0], groups[1][0][0]
    assert np.all(I[pc23, :100] == 0) and I[pc23, 100:500].max() == pytest.approx(50.0, rel=1e-3)
    assert np.all(I[pc5, 100:] == 30.0) and np.all(I[other] == 0)
    assert np.all(I[pc23, 500:] == I[pc23, 499])                      # TimedArray holds its last value


def test_noise_variant():
    """`noise = D` (codes/Main.py:31): noise_D = D, nCon[i] = incoming + outgoing synapses of i, Euler-Maruyama on g_X_off."""
    g = load_golden("net_n200_s0")
    groups = [[list(map(int, s)) for s in grp] for grp in golden_groups(g)]
    args = (g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], (g["target_arr"], g["source_arr"]), groups, [[0] * 14], [], NOSCALE)
    quiet = CortexNetwork(*args, False, "euler", 0.05, 1, "", _engine_cls=OracleEngine)
    noisy = CortexNetwork(*args, 0.0005, "euler", 0.05, 1, "", _engine_cls=OracleEngine)
    for c in (quiet, noisy):
        c.neuron_monitors([["g", ["g_AMPA"], [["all", "all"]]]])
        c.run(20 * ms)
    n = g["NeuPar"].shape[1]
    ncon = np.bincount(g["target_arr"], minlength=n) + np.bincount(g["source_arr"], minlength=n)
    gq, gn = quiet.neuronmonitors["g"].g_AMPA, noisy.neuronmonitors["g"].g_AMPA
    assert not gq.any() and gn[ncon > 0].std() > 0
    assert not gn[ncon == 0].any()                                    # nCon = 0: no noise
