"""`ParamSetup(N1, Nstripes, scales)` with the reference's signature and return tuple (codes/ParamSetup.py:5, 589-601).

The numeric tables of the model (cell-type fractions, Box-Cox/multivariate-normal membrane-parameter statistics,
connection probabilities, peak conductances, delays, STP classes, synapse kinetics, inter-stripe rules — Hass et al.
2016, Tables 1-4) are data: they live in ../data/pfc_constants.npz, extracted from the reference by
tools/extract_reference_constants.py.  What is code here is only what depends on the arguments:
group sizes `ceil(N1 * percent / 100)` (:9-10) and the scale factors (:212, 343-363, 474-514).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "pfc_constants.npz")
_PC = [0, 7]
_IN = [1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13]


def _constants():
    d = np.load(_DATA, allow_pickle=False)
    return {k: d[k] for k in d.files}


def get_gmax(PSP, i, j):
    """Peak conductance (nS) that gives a PSP of `PSP` mV on a cell of target group i from a source of group j: one factor per
    target group for pyramidal sources (j = 0, 7) and one for interneuron sources (codes/ParamSetup.py:603-613).  The stored
    Syngmax table already contains this conversion."""
    c = _constants()
    return PSP * float(c["gmax_per_psp_from_PC" if j in (0, 7) else "gmax_per_psp_from_IN"][i])


def ParamSetup(N1, Nstripes, scales):
    c = _constants()
    gmax_scale, pCon_scale, param_std_scale = scales
    NTypes = np.ceil(N1 * c["percent"] / 100)
    N = int(round(np.sum(NTypes), 0))
    if N != N1:
        print("WARNING: Neurons can not be distributed evenly across cell types. Number of generated neurons ({}) differs "
              "from specified number ({}).\n".format(N, N1))
    N_sig = c["N_sig"] * param_std_scale

    Syngmax = c["Syngmax"].copy()
    for rows, cols, f in ((_PC, _PC, gmax_scale[0]), (_IN, _IN, gmax_scale[3]), (_PC, _IN, gmax_scale[1]), (_IN, _PC, gmax_scale[2])):
        Syngmax[np.ix_(rows, cols)] = Syngmax[np.ix_(rows, cols)] * f
    pCon = c["pCon"].copy()
    for rows, cols, f in ((_PC, _PC, pCon_scale[0]), (_IN, _PC, pCon_scale[1]), (_PC, _IN, pCon_scale[2]), (_IN, _IN, pCon_scale[3])):
        pCon[np.ix_(rows, cols)] = f * pCon[np.ix_(rows, cols)]

    STSP_setup = {str(k): c["STSP_tables"][q] for q, k in enumerate(c["STSP_names"])}
    get_SynType = {int(k): [int(v) for v in row if v != 0] for k, row in zip(c["SynType_keys"], c["SynType_vals"])}
    get_STSP_prob = {int(k): [float(v) for v in row if not np.isnan(v)] for k, row in zip(c["STSP_prob_keys"], c["STSP_prob_vals"])}
    get_STSP_types = {int(k): [str(v) for v in row if str(v)] for k, row in zip(c["STSP_types_keys"], c["STSP_types_vals"])}
    inter, pos = [], 0
    for n in c["interStripes_len"]:
        inter.append([int(v) for v in c["interStripes_flat"][pos:pos + int(n)]])
        pos += int(n)
    ConParStripes = ([[int(a), int(b)] for a, b in c["pConStripes"]], [[float(a), float(b)] for a, b in c["coefStripes"]], inter)

    return (NTypes, c["CellPar"], c["V0Par"], c["k_trans"], c["ParCov"], N_sig, c["N_max"], c["N_min"],
            c["SynTypes"], Syngmax, c["Syndelay"], c["Synpfail"], c["STSP_prob"], c["STSP_types"], pCon, c["cluster_flag"],
            c["S_sig"], c["S_max"], c["S_min"], float(c["p_fail"]),
            STSP_setup["E1"], STSP_setup["E2"], STSP_setup["E3"], STSP_setup["I1"], STSP_setup["I2"], STSP_setup["I3"],
            STSP_setup, get_SynType, get_STSP_prob, get_STSP_types, c["STypPar"], ConParStripes)
