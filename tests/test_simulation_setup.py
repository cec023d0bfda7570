"""codes/SimulationSetup.py: the compact npz form of a saved network and reading of a directory in the reference's text form."""
import os

import numpy as np

from oracle_binding import golden_groups, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.codes import SimulationSetup as SS
from rescience_hass_hertag_durstewitz_2016_b200.codes.NetworkSetup import sourcetarget


def golden(name="net_n200_s0"):
    g = load_golden(name)
    groups = [[list(map(int, s)) for s in grp] for grp in golden_groups(g)]
    return g, groups


def test_npz_round_trip(tmp_path, monkeypatch):
    g, groups = golden("net_n60x5_s1")
    monkeypatch.chdir(tmp_path)
    d = SS.set_dir()
    assert d == "Simulation_0" and SS.set_dir() == "Simulation_1"
    SS.save_setup(g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], (g["target_arr"], g["source_arr"]), groups, d)
    assert os.path.getsize(os.path.join(d, "setup.npz")) < 2_000_000
    NeuPar, V0, STypPar, SynPar, SPMtx, gd = SS.set_setup(0)
    assert np.array_equal(NeuPar, g["NeuPar"]) and np.array_equal(V0, g["V0"]) and np.array_equal(SynPar, g["SynPar"])
    assert np.array_equal(STypPar, g["STypPar"]) and gd == groups
    assert SPMtx.shape == (340, 340, 2)
    t, s = sourcetarget(SPMtx)                                  # the dense form indexes the synapses as the reference's does
    assert np.array_equal(t, g["target_arr"]) and np.array_equal(s, g["source_arr"])
    # beyond the dense limit the connectivity comes back as the pair of arrays
    _, _, _, _, pair, _ = SS.set_setup(0, dense_limit=100)
    assert np.array_equal(pair[0], g["target_arr"]) and np.array_equal(pair[1], g["source_arr"])
    assert SS.set_group_distr(0) == groups


def test_reads_a_directory_in_the_references_text_form(tmp_path, monkeypatch):
    """What codes/SimulationSetup.py:70-80 writes: np.savetxt of the four tables and of the two dense layers of SPMtx."""
    g, groups = golden()
    monkeypatch.chdir(tmp_path)
    os.mkdir("Simulation_3")
    n = g["NeuPar"].shape[1]
    dense = SS._dense(g["target_arr"].astype(np.int64), g["source_arr"].astype(np.int64), n, 2)
    for name, arr in (("NeuPar", g["NeuPar"]), ("V0", g["V0"]), ("STypPar", g["STypPar"]), ("SynPar", g["SynPar"]),
                      ("SPMtx0", dense[:, :, 0]), ("SPMtx1", dense[:, :, 1])):
        np.savetxt("Simulation_3/%s.txt" % name, arr, delimiter=",")
    SS.save_group_distr(groups, "Simulation_3")
    NeuPar, V0, STypPar, SynPar, SPMtx, gd = SS.set_setup(3)
    assert np.array_equal(NeuPar, g["NeuPar"]) and np.array_equal(SynPar, g["SynPar"]) and gd == groups
    t, s = sourcetarget(SPMtx)
    assert np.array_equal(t, g["target_arr"]) and np.array_equal(s, g["source_arr"])
