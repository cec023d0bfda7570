"""The revision's own network generator (codes_revision/NetworkSetup.py:15-170) — our port against networks the REFERENCE's
code drew in this container for the same seeds (tests/golden/rev_net_*.npz, see tests/golden/make_golden_revision.py):
every membrane parameter, I_ref, group membership after re-sorting, every synapse and every synaptic parameter bit-identical."""
import os

import numpy as np
import pytest

from oracle_binding import OracleEngine
from rescience_hass_hertag_durstewitz_2016_b200 import units
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.CortexSetup import Cortex
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision.NetworkSetup import Network, network_setup
from test_basics_setup import scales_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def assert_same(net, g):
    assert np.array_equal(net.membr_params.values, g["membr_params"])
    assert np.array_equal(net.refractory_current.values, g["I_ref"])
    assert np.array_equal(net.syn_pairs, g["syn_pairs"])
    for k in ("channel", "spiking", "STSP_params", "delay"):
        assert np.array_equal(net.syn_params[k].values, g["syn_" + k]), k
    assert [x for st in net.group_distr for grp in st for x in grp] == list(g["group_flat"])
    assert list(np.cumsum([0] + [len(grp) for st in net.group_distr for grp in st])) == list(g["group_offs"])


@pytest.mark.parametrize("n,seed,scaled", [(150, 3, False), (300, 11, True), (400, 0, False)])
def test_bit_identical_to_the_reference_generator(n, seed, scaled):
    g = np.load(os.path.join(GOLDEN, "rev_net_n%d_s%d%s.npz" % (n, seed, "_scaled" if scaled else "")))
    net = network_setup(n, 1, scales_case() if scaled else None, seed)
    assert_same(net, g)
    assert net.basics.scales is (None if not scaled else net.basics.scales) and net.seed == seed


def test_what_the_revision_generator_does_differently_from_codes():
    """The differences a caller switching between the two generators sees (module docstring of NetworkSetup.py)."""
    net = network_setup(150, 1, None, 3)
    gi = net.basics.struct.groups.idcs
    # pCon[PC_L5 <- IN_CC_L5] = 0 in the revision: no such synapse
    assert len(net.syn_idcs_from_groups(("PC_L5", 0), ("IN_CC_L5", 0))) == 0
    assert len(net.syn_idcs_from_groups(("PC_L5", 0), ("IN_CC_L23", 0))) > 0
    # intra-stripe delays carry the extra 1 ms: PC_L23 -> PC_L23 mean 1.5465 + 1
    d = net.syn_params["delay"].values[0][net.syn_idcs_from_groups(("PC_L23", 0), ("PC_L23", 0), channel="AMPA")]
    assert 2.2 < d.mean() < 2.9
    # AMPA and NMDA twins share target, source and delay, not gmax / STSP parameters
    a = net.syn_idcs_from_groups(("PC_L5", 0), ("PC_L23", 0), channel="AMPA")
    m = net.syn_idcs_from_groups(("PC_L5", 0), ("PC_L23", 0), channel="NMDA")
    assert len(a) == len(m) and np.array_equal(net.syn_pairs[:, a], net.syn_pairs[:, m])
    assert np.array_equal(net.syn_params["delay"].values[0][a], net.syn_params["delay"].values[0][m])
    assert not np.array_equal(net.syn_params["spiking"].values[0][a], net.syn_params["spiking"].values[0][m])
    assert gi["PC_L5"] == 7 and net.channel_constants()["mg_fac"] == 1.0


def test_save_and_load_keep_the_scaled_tables(tmp_path):
    net = network_setup(150, 1, scales_case(), 3)
    net.save(str(tmp_path / "scaled"))
    back = Network.load(str(tmp_path / "scaled"))
    assert np.array_equal(back.basics.struct.conn.pCon.values, net.basics.struct.conn.pCon.values)
    assert np.array_equal(back.basics.syn.spiking.params.gmax["mean"].values, net.basics.syn.spiking.params.gmax["mean"].values)
    assert back.basics.scales == scales_case() and np.array_equal(back.syn_params["spiking"].values, net.syn_params["spiking"].values)


def test_more_than_one_stripe_is_refused_like_the_reference_fails():
    with pytest.raises(NotImplementedError):
        network_setup(100, 3, None, 0)
    net = network_setup(60, 3, None, 1, generator="codes", workers=2)
    assert net.basics.struct.Ncells_total == net.membr_params.values.shape[1] and len(net.group_distr) == 3
    assert net.channel_constants()["mg_fac"] == 0.33


def test_cortex_on_a_revision_network_runs_and_round_trips(tmp_path):
    net = network_setup(150, 1, None, 3)
    net.save(str(tmp_path / "net"))
    back = Network.load(str(tmp_path / "net"))
    assert np.array_equal(back.syn_pairs, net.syn_pairs) and back.channel_constants()["mg_fac"] == 1.0
    c = Cortex(back, [[("PC", 0), 250], [("IN", 0), 200]], "rk4", 0.05, seed=5, accum="fixed", _engine_cls=OracleEngine)
    c.run(200)
    q = Cortex(net, [[("PC", 0), 250], [("IN", 0), 200]], "rk4", 0.05, seed=5, accum="fixed", reference_quirks=True, _engine_cls=OracleEngine)
    q.run(200)
    assert c.spikemonitor.num_spikes > 20 and c.neurons.t / units.ms == pytest.approx(200.0)
    assert not np.array_equal(c.spikemonitor.t, q.spikemonitor.t)          # tau_fac <- tau_rec changes the dynamics
