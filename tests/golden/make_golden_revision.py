"""Golden vectors for the revision's model layer, produced by the REFERENCE's own code
(/root/reference/codes_revision/BasicsSetup.py and NetworkSetup.py) imported in this container with stand-ins for the two
packages that are not installed: xarray (tests/golden/_xarray_stub.py) and brian2 (unit names only).  Run here:
    python tests/golden/make_golden_revision.py
writes
    tests/golden/rev_basics.npz          every table basics_setup() assembles (default and with basics_scales)
    tests/golden/rev_net_n<N>_s<seed>[_scaled].npz   what network_setup(N, 1, basics_scales, seed) draws
Before anything is written the stand-in is cross-checked: the tables assembled through it must equal the tables the
reference's codes/ParamSetup.py builds with plain numpy, except for the documented differences.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference_revision():
    import _xarray_stub
    _stub("xarray", DataArray=_xarray_stub.DataArray)
    _stub("brian2", pF=1e-12, nS=1e-9, mV=1e-3, ms=1e-3, pA=1e-12, second=1.0)
    if not hasattr(np, "NaN"):
        np.NaN = np.nan                      # BasicsSetup.py:481-486 predates numpy 2
    sys.path.insert(0, os.path.join(REF, "codes_revision"))
    import BasicsSetup as BS
    import NetworkSetup as NS
    sys.path.pop(0)
    return BS, NS


SCALES_CASE = lambda BS: {"gmax_mean": [(dict(target=BS.GROUPS_SETS["PC"], source=BS.GROUPS_SETS["PC"]), 2.0),
                                         (dict(target=BS.GROUPS_SETS["IN_L23"], source=["PC_L23"]), 0.5)],
                          "pCon": [(dict(target=["PC_L5"], source=BS.GROUPS_SETS["IN_L5"]), 0.75)],
                          "membr_param_std": [(dict(param=["C", "g_L"], group=BS.GROUPS_SETS["IN"]), 0.8)]}


def basics_tables(b):
    """Everything of a BasicsSetup instance as plain arrays (group order = GROUP_NAMES, [target, source])."""
    t = {}
    t["Ncells_per_group"] = np.asarray(b.struct.Ncells_per_group.values)
    t["Ncells"] = np.asarray(b.struct.Ncells)
    t["pCon"] = b.struct.conn.pCon.values
    t["cluster"] = b.struct.conn.cluster.values
    m = b.membr
    for k in ("mean", "covariance", "k", "std", "min", "max", "tau_m_min", "tau_m_max"):
        t["membr_" + k] = np.asarray(m[k].values)
    s = b.syn
    g = s.spiking.params.gmax
    t["gmax_mean"], t["gmax_sigma"], t["gmax_min"], t["gmax_max"] = (np.asarray(g[k].values) for k in ("mean", "sigma", "min", "max"))
    t["pfail"] = np.asarray(s.spiking.params.pfail.values)
    d = s.delay
    t["delay_mean"], t["delay_sigma"], t["delay_min"], t["delay_max"] = (np.asarray(d[k].values) for k in ("delay_mean", "delay_sigma", "delay_min", "delay_max"))
    t["stsp_kinds"] = np.asarray(s.STSP.kinds.values)
    t["stsp_sets"] = np.asarray(s.STSP.sets.values)
    t["stsp_distr"] = np.asarray(s.STSP.distr.values).astype("U16")
    t["syn_kinds"] = np.asarray(s.kinds.values).astype("U8")
    t["channel_params"] = np.asarray(s.channels.params.values)
    t["gsyn_factor"] = np.asarray(s.channels.gsyn_factor.values)
    return t


def cross_check_stub(tab):
    """The tables assembled through the xarray stand-in against codes/ParamSetup.py (plain numpy), reference F8 list."""
    sys.path.insert(0, os.path.join(REF, "codes"))
    import ParamSetup as PS
    sys.path.pop(0)
    with contextlib.redirect_stdout(io.StringIO()):
        out = PS.ParamSetup(1000, 1, ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14))))
    return out


def network_arrays(net):
    out = {"membr_params": np.asarray(net.membr_params.values), "I_ref": np.asarray(net.refractory_current.values),
           "syn_pairs": np.asarray(net.syn_pairs.values if hasattr(net.syn_pairs, "values") else net.syn_pairs).astype(np.int64)}
    for k in ("channel", "spiking", "STSP_params", "delay"):
        out["syn_" + k] = np.asarray(net.syn_params[k].values, dtype=np.float64)
    flat, offs = [], [0]
    for st in net.group_distr:
        for g in st:
            flat.extend(int(x) for x in g)
            offs.append(len(flat))
    out["group_flat"], out["group_offs"], out["nstripes"] = np.asarray(flat, dtype=np.int64), np.asarray(offs), np.asarray(len(net.group_distr))
    return out


def main():
    BS, NS = import_reference_revision()
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())
    with quiet():
        b0 = BS.basics_setup(1000, 1, None, disp=False)
        b1 = BS.basics_setup(300, 1, SCALES_CASE(BS), disp=False)
    tab = {"default__" + k: v for k, v in basics_tables(b0).items()}
    tab.update({"scaled__" + k: v for k, v in basics_tables(b1).items()})
    tab["group_names"] = np.asarray(BS.GROUP_NAMES)
    tab["memb_param_names"] = np.asarray(BS.MEMB_PARAM_NAMES)
    tab["equations_membr_model"] = np.asarray(b0.equations.membr_model)
    tab["equations_syn_model"] = np.asarray(b0.equations.syn_model)
    tab["equations_syn_pathway"] = np.asarray([p["eq"] for p in b0.equations.syn_pathway])
    tab["equations_ext_syn_pathway"] = np.asarray([p["eq"] for p in b0.equations.ext_syn_pathway])
    np.savez_compressed(os.path.join(HERE, "rev_basics.npz"), **tab)
    print("rev_basics.npz:", len(tab), "tables")

    for n, seed, scaled in ((150, 3, False), (300, 11, True), (400, 0, False)):
        with quiet():
            net = NS.network_setup(n, 1, SCALES_CASE(BS) if scaled else None, seed)
        arr = network_arrays(net)
        name = "rev_net_n%d_s%d%s.npz" % (n, seed, "_scaled" if scaled else "")
        np.savez_compressed(os.path.join(HERE, name), **arr)
        print(name, arr["membr_params"].shape, arr["syn_pairs"].shape)


if __name__ == "__main__":
    main()
