"""codes_revision/BasicsSetup.py port: every table, the equation strings and the basics_scales selectors against what the
REFERENCE's own module assembled (tests/golden/rev_basics.npz, written by tests/golden/make_golden_revision.py)."""
import os

import numpy as np
import pytest

from rescience_hass_hertag_durstewitz_2016_b200.codes_revision import BasicsSetup as B

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rev_basics.npz"))


def scales_case():
    """The selectors used when the golden file was made (make_golden_revision.SCALES_CASE)."""
    S = B.GROUPS_SETS
    return {"gmax_mean": [(dict(target=S["PC"], source=S["PC"]), 2.0), (dict(target=S["IN_L23"], source=["PC_L23"]), 0.5)],
            "pCon": [(dict(target=["PC_L5"], source=S["IN_L5"]), 0.75)],
            "membr_param_std": [(dict(param=["C", "g_L"], group=S["IN"]), 0.8)]}


def tables(b):
    g = b.syn.spiking.params.gmax
    d = b.syn.delay
    t = {"Ncells_per_group": b.struct.Ncells_per_group.values, "Ncells": b.struct.Ncells, "pCon": b.struct.conn.pCon.values,
         "cluster": b.struct.conn.cluster.values, "gmax_mean": g["mean"].values, "gmax_sigma": g["sigma"].values,
         "gmax_min": g["min"].values, "gmax_max": g["max"].values, "pfail": b.syn.spiking.params.pfail.values,
         "delay_mean": d.delay_mean.values, "delay_sigma": d.delay_sigma.values, "delay_min": d.delay_min.values,
         "delay_max": d.delay_max.values, "stsp_kinds": b.syn.STSP.kinds.values, "stsp_sets": b.syn.STSP.sets.values,
         "stsp_distr": b.syn.STSP.distr.values, "syn_kinds": b.syn.kinds.values, "channel_params": b.syn.channels.params.values,
         "gsyn_factor": b.syn.channels.gsyn_factor.values}
    for k in ("mean", "covariance", "k", "std", "min", "max", "tau_m_min", "tau_m_max"):
        t["membr_" + k] = b.membr[k].values
    return t


@pytest.mark.parametrize("prefix,args", [("default__", (1000, 1, None)), ("scaled__", (300, 1, "scales"))])
def test_tables_equal_the_reference(prefix, args):
    n, stripes, sc = args
    b = B.basics_setup(n, stripes, scales_case() if sc else None, disp=False)
    for name, mine in tables(b).items():
        ref = GOLD[prefix + name]
        mine = np.asarray(mine)
        if ref.dtype.kind in "US":
            assert np.array_equal(mine.astype(ref.dtype), ref), name
        else:
            assert mine.shape == ref.shape and np.array_equal(mine.astype(np.float64), ref.astype(np.float64), equal_nan=True), name


def test_scales_do_not_leak_into_the_module_tables():
    before = B.PCON.values.copy()
    B.basics_setup(300, 1, scales_case(), disp=False)
    assert np.array_equal(B.PCON.values, before)
    assert np.array_equal(B.basics_setup(300, 1, None, disp=False).struct.conn.pCon.values, before)


def test_equation_strings_equal_the_reference():
    b = B.basics_setup(1000, 1, None, disp=False)
    assert b.equations.membr_model == str(GOLD["equations_membr_model"])
    assert b.equations.syn_model == str(GOLD["equations_syn_model"])
    assert [p["eq"] for p in b.equations.syn_pathway] == list(GOLD["equations_syn_pathway"])
    assert [p["eq"] for p in b.equations.ext_syn_pathway] == list(GOLD["equations_ext_syn_pathway"])
    assert [p["order"] for p in b.equations.syn_pathway] == [0, 1, 2, 3, 4, 5, 6] + [7] * 6
    assert [p["delay"] for p in b.equations.syn_pathway] == [False] * 7 + [True] * 6
    assert b.equations.rheobase(2.0, -50.0, -70.0, 3.0) == 2.0 * (-50.0 + 70.0 - 3.0)


def test_the_revisions_documented_slips_are_present():
    """they are what makes the revision's numbers differ from codes/ — kept on purpose, each pinned here."""
    b = B.basics_setup(1000, 1, None, disp=False)
    pcon = b.struct.conn.pCon
    assert float(pcon.loc["PC_L5", "IN_CC_L5"]) == 0.0 and float(pcon.loc["PC_L5", "IN_CC_L23"]) == 0.7006
    assert float(b.syn.channels.params.loc["NMDA", "Mg_fac"]) == 1.0
    assert b.syn.STSP.decl_vars["tau_fac"]["value"] == "tau_rec"
    m = B.engine_model(b)
    assert m["mg_fac"] == 1.0 and m["last_spike0_ms"] == -1000.0 and m["tau_fac_from"] == "tau_rec"
    assert np.allclose(m["gsyn"], [1.0 * 10 * 1.4 / 8.6, 1.0 * 40 * 3 / 37, 1.09 * 75 * 4.3 / 70.7])


def test_table_selection_semantics():
    t = B.PCON.copy()
    assert t.loc[dict(target=["PC_L23", "PC_L5"], source=B.GROUPS_SETS["IN_L5"])].shape == (2, 6)       # orthogonal, not pairwise
    assert t.loc[dict(target="PC_L23")].shape == (14,)
    t.loc[dict(target=["PC_L23", "PC_L5"], source="PC_L23")] = [0.5, 0.25]
    assert float(t.loc["PC_L5", "PC_L23"]) == 0.25 and float(B.PCON.loc["PC_L5", "PC_L23"]) == 0.2333
    assert list(t.coords["target"]) == B.GROUP_NAMES
    b = B.basics_setup(1000, 1, None, disp=False)
    assert b.struct.Ncells == 1003 and b.struct.Ncells_total == 1003 and b["struct"]["groups"].N == 14
    assert set(b.keys()) == {"struct", "membr", "syn", "equations", "scalables", "scales"}
