"""The codes_revision façade (Cortex, Network) on the CPU oracle."""
import numpy as np
import pytest

from oracle_binding import OracleEngine, golden_groups, load_golden
from rescience_hass_hertag_durstewitz_2016_b200 import units
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.CortexSetup import Cortex
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.NetworkSetup import Network, network_setup

CONST = [[("PC", 0), 250], [("IN", 0), 200]]


def net_from_golden(name="net_n200_s0"):
    g = load_golden(name)
    groups = [[list(map(int, s)) for s in grp] for grp in golden_groups(g)]
    return Network.from_codes(g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], g["target_arr"], g["source_arr"], groups, seed=0), g


def test_network_container_queries(tmp_path):
    net, g = net_from_golden("net_n60x5_s1")
    assert net.basics.struct.Ncells_total == 340 and net.basics.struct.Nstripes == 5
    pc0 = net.neuron_idcs(("PC", 0))
    assert sorted(pc0) == sorted(list(golden_groups(g)[0][0]) + list(golden_groups(g)[7][0]))
    assert len(net.neuron_idcs([("ALL", 0), ("ALL", 1)])) == 68 * 2
    assert np.array_equal(net.membr_params.loc[dict(param="g_L")].values, g["NeuPar"][1])
    assert np.array_equal(net.membr_params.loc[dict(param="V_T", cell_index=[3, 5])].values, g["NeuPar"][9, [3, 5]])
    syn = net.syn_idcs_from_groups(("PC_L5", 1), ("PC_L23", 1), channel="NMDA")
    assert len(syn) and np.all(g["SynPar"][0, syn] == 3)
    assert np.allclose(net.rheobase([0]), g["NeuPar"][1, 0] * (g["NeuPar"][9, 0] - g["NeuPar"][2, 0] - g["NeuPar"][3, 0]))
    net.save(str(tmp_path / "Network1"))
    back = Network.load(str(tmp_path / "Network1"))
    assert np.array_equal(back.syn_pairs, net.syn_pairs) and back.group_distr == net.group_distr and back.seed == 0
    assert np.array_equal(back.syn_params["STSP_params"].values, net.syn_params["STSP_params"].values)


def test_cortex_run_transient_and_monitor_schedule():
    net, g = net_from_golden()
    c = Cortex(net, CONST, "rk4", 0.05, transient=50, seed=4, _engine_cls=OracleEngine)
    c.set_neuron_monitors("V", "V", ("PC_L23", 0), start=60, stop=80)
    c.set_longrun_neuron_monitors("I_tot", "I_tot", ("ALL", 0), 5000, start=50, stop=None, population_agroupate="sum")
    c.run(100)
    assert c.neurons.t / units.ms == pytest.approx(100.0) and c.neurons.N == 207
    sm = c.spikemonitor
    assert sm.num_spikes > 0 and (sm.t / units.ms).min() >= 50.0          # nothing recorded during the transient
    V = c.recorded.var("V")
    tV = c.recorded.t("V") / units.ms
    assert V.shape[1] == 400 and tV[0] == pytest.approx(60.0) and tV[-1] == pytest.approx(79.95)
    I = c.recorded.var("I_tot")
    tI = c.recorded.t("I_tot") / units.ms
    assert I.shape == (1000,) and tI[0] == pytest.approx(50.0)
    assert I.min() / units.pA > 0.5 * 207 * 200 * 0.5                      # summed currents, order of N x I_DC
    c.run(20)                                                              # continue-run
    assert c.neurons.t / units.ms == pytest.approx(120.0) and c.recorded.var("I_tot").shape == (1400,)
    cnt = c.spiking_count()
    assert cnt.sum() == sm.num_spikes
    assert c.spiking_count(tlim=(50, 120)).sum() == sm.num_spikes
    binned = c.spiking_count(neuron_idcs=np.arange(207), tlim=(50, 120), delta_t=10)
    assert binned.shape == (207, 7) and binned.sum() == sm.num_spikes
    rate = c.spiking_rate()
    assert rate.shape == (207,) and rate.max() > 10
    assert len(c.spiking_idcs((np.greater, 0.0))) == np.count_nonzero(cnt)
    assert len(c.spiking_idcs([(np.greater, 0.0)], groupstripe=("PC", 0))) <= len(c.neuron_idcs(("PC", 0)))


def test_cortex_matches_engine_level_run_with_revision_defaults():
    """Same arrays through the façade and through the engine API (last_spike0 = -1000 ms as the revision sets it)."""
    net, g = net_from_golden()
    c = Cortex(net, CONST, "rk2", 0.05, seed=9, accum="fixed", _engine_cls=OracleEngine)
    c.run(150)
    from helpers import make_engine
    e = make_engine("cpu", g, method="rk2", seed=9, accum="fixed", last_spike0_ms=-1000.0)
    e.run(3000)
    i, s = e.spikes()
    assert len(i) > 100 and np.array_equal(c.spikemonitor.i, i)


def test_stimuli():
    net, g = net_from_golden()
    np.random.seed(3)
    c = Cortex(net, None, "rk2", 0.05, seed=3, _engine_cls=OracleEngine)
    tgt = c.neuron_idcs(("PC_L23", 0))
    c.set_regular_stimuli("reg", 1, ["AMPA", "NMDA"], tgt, 0.5, 10000, 5, 10, 2.0, 0.0)
    c.set_poisson_stimuli("poi", 10, "GABA", tgt, 0.1, 500, 20, 40, 1.0, 0.3)
    c.set_neuron_monitors("g", ["g_AMPA", "g_NMDA", "g_GABA"], ("PC_L23", 0))
    c.run(50)
    gm = c.neuron_monitors["g"]
    assert gm.g_AMPA[:, :100].max() == 0 and gm.g_AMPA.max() > 0 and gm.g_NMDA.max() > 0
    assert gm.g_GABA.max() > 0 and c.engine.counters()["ext_events"] > 50
    with pytest.raises(ValueError):
        c.set_regular_stimuli("bad", 1, "AMPA", tgt, 0.5, 1e6, 60, 70, 1.0, 0.0)


def test_network_setup_draws_the_codes_network():
    net = network_setup(200, 1, None, 0, generator="codes", workers=4)
    g = load_golden("net_n200_s0")
    assert np.array_equal(net.syn_pairs[0], g["target_arr"]) and np.array_equal(net.syn_pairs[1], g["source_arr"])
    assert np.array_equal(net.membr_params.values[1], g["NeuPar"][1])
    assert np.array_equal(net.refractory_current.values, g["NeuPar"][10])


def test_monitor_times_follow_the_active_intervals_and_spike_cache_is_invalidated():
    """A monitor switched off and on again has a gap in `.t` (advisor finding, round 1); clearing the spike monitor shows at once."""
    from rescience_hass_hertag_durstewitz_2016_b200.monitors import SpikeMonitorView, StateMonitorView
    from helpers import make_engine
    g = load_golden("net_n200_s0")
    e = make_engine("cpu", g, method="euler", seed=2)
    sv = StateMonitorView(e, "V", [0, 5])
    sm = SpikeMonitorView(e, 207)
    e.run(40)
    sv.set_active(False)
    e.run(100)
    sv.set_active(True)
    e.run(20)
    t = sv.t / units.ms
    assert sv.V.shape == (2, 60) and len(t) == 60
    assert np.allclose(t[:40], np.arange(40) * 0.05) and np.allclose(t[40:], (140 + np.arange(20)) * 0.05)
    e.run(3000)
    n = sm.num_spikes
    assert n > 0
    e.clear_monitor(-1)
    assert sm.num_spikes == 0
