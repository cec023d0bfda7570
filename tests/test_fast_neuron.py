"""The builder's fast path: codes/fast_neuron.py against the exact scipy path that is bit-identical with the
reference, and the many-column builder of scaled.py on top of it."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from oracle_binding import golden_groups, load_golden
from rescience_hass_hertag_durstewitz_2016_b200 import scaled
from rescience_hass_hertag_durstewitz_2016_b200.codes import NetworkSetup as NS
from rescience_hass_hertag_durstewitz_2016_b200.codes import fast_neuron as FN

ONES = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))


def test_reference_current_matches_the_references_fmin_values():
    """I_ref of every neuron of the three golden networks (produced by the reference's own fmin + quad): the bisection lands
    within 1e-4 pA, well inside the 1e-4 tolerances Nelder-Mead itself stops at."""
    for name in ("net_n1000_s0_full", "net_n200_s0", "net_n60x5_s1"):
        P = load_golden(name)["NeuPar"]
        assert np.abs(FN.reference_current(P) - P[10]).max() < 1e-4


def test_rates_latencies_and_adaptation_ratio_match_the_exact_path():
    P = load_golden("net_n200_s0")["NeuPar"]
    idx = np.r_[0:12, 95:110, 150:207]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exact = np.array([NS._per_neuron(P[:, i].copy()) for i in idx])
        Iref, t_lat, t_lif, ratio = FN.per_neuron(P[:, idx])
        assert np.abs(Iref - exact[:, 0]).max() < 1e-4
        assert np.allclose(t_lat, exact[:, 1], rtol=1e-10) and np.allclose(t_lif, exact[:, 2], rtol=1e-13, equal_nan=True)
        assert np.abs(ratio - exact[:, 3]).max() < 1e-4
        for I in (60.0, 175.0, 420.0):
            fast = FN.firing_rate(P[:, idx], I, None, None)
            ref = np.array([np.nan if r is None else float(np.asarray(r).reshape(-1)[0]) for r in (NS.FRsimpAdEx(P[:, i], I, [], [], []) for i in idx)])
            assert np.array_equal(np.isfinite(fast), np.isfinite(ref))
            ok = np.isfinite(ref)
            assert np.abs(fast[ok] / ref[ok] - 1).max() < 1e-6


@pytest.mark.parametrize("name,n1,stripes,seed", [("net_n200_s0", 200, 1, 0), ("net_n60x5_s1", 60, 5, 1), ("net_n40x7_s2", 40, 7, 2), ("net_n40x9_s3", 40, 9, 3)])
def test_fast_builder_draws_the_references_network(name, n1, stripes, seed):
    """sparse_network(fast=True) = the reference's network — every membrane parameter, group membership after re-typing, every
    synapse and synaptic parameter bit-identical — except I_ref, which agrees to 1e-4 pA."""
    g = load_golden(name)
    np.random.seed(seed)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        NeuPar, V0, STyp, SynPar, (tgt, src), gd = NS.sparse_network(n1, stripes, ONES, True, fast=True)
    rows = [r for r in range(12) if r != 10]
    assert np.array_equal(NeuPar[rows], g["NeuPar"][rows]) and np.abs(NeuPar[10] - g["NeuPar"][10]).max() < 1e-4
    assert np.array_equal(SynPar, g["SynPar"]) and np.array_equal(tgt, g["target_arr"]) and np.array_equal(src, g["source_arr"])
    ref_groups = golden_groups(g)
    assert all(list(gd[a][s]) == list(ref_groups[a][s]) for a in range(14) for s in range(stripes))


def test_distinct_columns_are_independent_reference_draws(tmp_path):
    cols = scaled.build_distinct_columns(200, 5, 10, seed=4, workers=2, cache_dir=str(tmp_path))
    again = scaled.build_distinct_columns(200, 6, 7, seed=4, workers=1, cache_dir=str(tmp_path))       # from the cache
    assert np.array_equal(again[0]["gmax"], cols[1]["gmax"]) and np.array_equal(again[0]["NeuPar"], cols[1]["NeuPar"])
    # column 7 = NetworkSetup(200, 1) after np.random.seed(column_seed(4, 7))
    np.random.seed(scaled.column_seed(4, 7))
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        NeuPar, V0, STyp, SynPar, (tgt, src), gd = NS.sparse_network(200, 1, ONES, True, fast=True)
    c = cols[2]
    assert np.array_equal(c["NeuPar"], NeuPar) and np.array_equal(c["tgt"], tgt) and np.array_equal(c["delay"], SynPar[5])
    assert not np.array_equal(cols[0]["NeuPar"], cols[1]["NeuPar"])
    net = scaled.distinct_columns_network(cols)
    n1 = c["NeuPar"].shape[1]
    assert net["NeuPar"].shape == (12, 5 * n1) and net["I_DC"].shape == (5 * n1,)
    pc = np.concatenate([net["groups"][0], net["groups"][7]])
    assert np.all(net["I_DC"][pc] == 250.0) and np.sum(net["I_DC"][:n1] == 200.0) == n1 - len(pc)
    syn = scaled.distinct_synapses_device(cols, 5, 40, "cpu", inter_column=True, seed=1)
    n_intra = sum(len(x["src"]) for x in cols)
    assert syn[0].numel() > n_intra and int(syn[1].min()) >= 5 * n1 and int(syn[1].max()) < 10 * n1      # targets are local columns
    assert int(syn[0][:n_intra].min()) >= 5 * n1 and int(syn[0][:n_intra].max()) < 10 * n1               # intra-column sources too
    # the groups that project across columns keep their ids in every draw (they are not re-typed)
    offs = np.concatenate([[0], np.cumsum(cols[0]["group_sizes"])])
    for g in (0, 7, 12, 13):
        assert all(np.array_equal(x["group_flat"][offs[g]:offs[g + 1]], cols[0]["group_flat"][offs[g]:offs[g + 1]]) for x in cols)


def test_inter_column_count_is_what_the_generator_draws():
    g = load_golden("net_n200_s0")
    groups = [grp[0] for grp in golden_groups(g)]
    n1 = g["NeuPar"].shape[1]
    syn = scaled.inter_column_synapses_device(groups, n1, 40, "cpu", 0, 40, seed=1)
    assert syn[0].numel() == scaled.inter_column_count(groups, 40)


def test_partitioned_build_is_the_same_network(tmp_path):
    """Every rank builds only the columns it owns (and the inter-column synapses that target them); the union over the ranks is the
    network a single rank builds — the column seeds and the chunked random stream of the inter-column draws do not depend on the
    partition."""
    import torch
    ncol = 10
    full_cols = scaled.build_distinct_columns(200, 0, ncol, seed=2, workers=2, cache_dir=str(tmp_path))
    full = scaled.distinct_synapses_device(full_cols, 0, ncol, "cpu", inter_column=True, seed=1)
    parts = []
    for rank in range(2):
        lo, hi = rank * 5, rank * 5 + 5
        cols = scaled.build_distinct_columns(200, lo, hi, seed=2, workers=1, cache_dir=None)      # rebuilt, not read from the cache
        assert all(np.array_equal(a["NeuPar"], b["NeuPar"]) and np.array_equal(a["gmax"], b["gmax"]) for a, b in zip(cols, full_cols[lo:hi]))
        parts.append(scaled.distinct_synapses_device(cols, lo, ncol, "cpu", inter_column=True, seed=1))
        n1 = cols[0]["NeuPar"].shape[1]
        assert int(parts[-1][1].min()) >= lo * n1 and int(parts[-1][1].max()) < hi * n1               # a rank holds the synapses onto ITS neurons
    union = [torch.cat([p[q] for p in parts]) for q in range(9)]
    assert union[0].numel() == full[0].numel()
    order = lambda t: np.lexsort(tuple(x.numpy() for x in (t[4], t[3], t[8], t[2], t[0], t[1])))      # by target, source, channel, delay, g_max, U
    a, b = order(union), order(full)
    for q in range(9):
        assert np.array_equal(union[q].numpy()[a], full[q].numpy()[b]), q
