"""Connectivity / parameter construction must be BIT-IDENTICAL to the reference's own NetworkSetup for the same numpy
seed (north_star: "connectivity ... must match bit-exactly").  The golden fixtures were produced by importing
/root/reference/codes/NetworkSetup.py (tests/golden/make_golden_network.py)."""
import contextlib
import hashlib
import io
import warnings

import numpy as np
import pytest

from oracle_binding import load_golden
from rescience_hass_hertag_durstewitz_2016_b200.codes.CortexNetwork import sourcetarget
from rescience_hass_hertag_durstewitz_2016_b200.codes.NetworkSetup import FRsimpAdEx, NetworkSetup, SetCon, rand_par, sparse_network
from rescience_hass_hertag_durstewitz_2016_b200.codes.ParamSetup import ParamSetup

SCALES = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(n1, nstripes, seed, sparse=False, workers=4):
    warnings.filterwarnings("ignore")
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        out = (sparse_network if sparse else NetworkSetup)(n1, nstripes, SCALES, True, workers=workers)
    return out, np.random.random()


@pytest.mark.parametrize("name", ["net_n200_s0", "net_n60x5_s1", "net_n40x7_s2", "net_n40x9_s3"])
def test_bit_identical_to_the_reference(name):
    """1, 5, 7 and 9 stripes: the stripe counts for which the reference's inter-stripe offsets (+-1, +-2, +-4 with wrap-around)
    do not collide."""
    g = load_golden(name)
    (NeuPar, V0, STypPar, SynPar, SPMtx, group_distr), rng_after = build(int(g["n1"]), int(g["nstripes"]), int(g["seed"]))
    assert np.array_equal(NeuPar, g["NeuPar"])
    assert np.array_equal(V0, g["V0"])
    assert np.array_equal(STypPar, g["STypPar"])
    assert SynPar.shape == g["SynPar"].shape and np.array_equal(SynPar, g["SynPar"])
    t, s = sourcetarget(SPMtx)
    assert np.array_equal(t, g["target_arr"]) and np.array_equal(s, g["source_arr"])
    assert rng_after == float(g["rng_after"])                      # the RNG stream was consumed identically
    ns = int(g["nstripes"])
    flat = [int(x) for grp in group_distr for st in grp for x in st]
    assert np.array_equal(flat, g["group_flat"])
    assert [len(st) for grp in group_distr for st in grp] == list(np.diff(g["group_offs"]))
    assert sha(NeuPar) == str(g["sha_NeuPar"]) and sha(SynPar) == str(g["sha_SynPar"])


def test_sparse_form_equals_dense_form():
    g = load_golden("net_n60x5_s1")
    (_, _, _, SynPar, (t, s), _), _ = build(int(g["n1"]), int(g["nstripes"]), int(g["seed"]), sparse=True)
    assert np.array_equal(SynPar, g["SynPar"])
    assert np.array_equal(t, g["target_arr"]) and np.array_equal(s, g["source_arr"])


def test_serial_and_parallel_setup_agree():
    (a, _, _, sa, _, _), ra = build(60, 1, 3, workers=1)
    (b, _, _, sb, _, _), rb = build(60, 1, 3, workers=4)
    assert np.array_equal(a, b) and np.array_equal(sa, sb) and ra == rb


def test_full_column_digest():
    """The reference's default 1003-neuron column, seed 0 (fixture net_n1000_s0: digests + thin sample)."""
    g = load_golden("net_n1000_s0")
    (NeuPar, V0, _, SynPar, (t, s), _), rng_after = build(1000, 1, 0, sparse=True, workers=8)
    assert sha(NeuPar) == str(g["sha_NeuPar"]) and sha(V0) == str(g["sha_V0"]) and sha(SynPar) == str(g["sha_SynPar"])
    assert sha(t.astype(np.int64)) == str(g["sha_target"]) and sha(s.astype(np.int64)) == str(g["sha_source"])
    assert rng_after == float(g["rng_after"])


def test_building_blocks():
    np.random.seed(1)
    X, idc = SetCon(10, 7, 5, 0.4)
    assert X.shape == (7, 5) and len(idc) == 14 and idc[0] == 11 and np.count_nonzero(X) == 14
    assert sorted(X[X > 0]) == list(idc)
    r = rand_par(1000, 1.0, 0.5, 0.0, 2.0, 0)
    assert r.min() >= 0 and r.max() <= 2 and abs(r.mean() - 1) < 0.1
    assert np.all(rand_par(5, 3.0, 0, 0, 0, 0) == 3.0)
    par = np.array([200.0, 10.0, -70.0, 2.0, -40.0, 100.0, 0.0, 50.0, -60.0, -50.0])
    assert FRsimpAdEx(par, 100.0) is None                        # below rheobase (180 pA): no spiking
    f_on, f_ss = FRsimpAdEx(par, 400.0, 0, -70.0, []), FRsimpAdEx(par, 400.0)
    assert f_on > f_ss > 0                                        # adaptation lowers the steady-state rate
    out = ParamSetup(1000, 1, SCALES)
    assert int(out[0].sum()) == 1003 and out[14].shape == (14, 14)


def test_get_gmax_conversion_factors():
    """codes/ParamSetup.py:603-613: one PSP -> g_max factor per target group, pyramidal and interneuron sources apart."""
    from rescience_hass_hertag_durstewitz_2016_b200.codes.ParamSetup import get_gmax
    assert get_gmax(1.0, 0, 0) == 1.0569 and get_gmax(1.0, 0, 7) == 1.0569 and get_gmax(1.0, 0, 3) == 2.3859
    assert get_gmax(0.5, 13, 0) == 0.5 * 0.6360 and get_gmax(2.0, 7, 12) == 2.0 * 3.5816
