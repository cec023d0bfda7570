"""Dump the model constants of the reference (the numeric tables inside codes/ParamSetup.py) into a data file.

Runs only where /root/reference exists.  The constants are DATA of the published model (Hass et al. 2016, Tables 1-4);
they are extracted by calling the reference's own `ParamSetup(N1, Nstripes, scales)` with unit scales and stored in
rescience_hass_hertag_durstewitz_2016_b200/data/pfc_constants.npz, so that our ParamSetup is a loader + the scale logic,
not a copy of the reference's source.  The cell-type fractions are recovered from NTypes at N1 = 1e6.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/codes")
import warnings
warnings.filterwarnings("ignore")
import contextlib, io
from ParamSetup import ParamSetup, get_gmax  # the reference's own code

ones = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))
with contextlib.redirect_stdout(io.StringIO()):
    big = ParamSetup(10 ** 6, 1, ones)
    out = ParamSetup(1000, 1, ones)
(NTypes, CellPar, V0Par, k_trans, ParCov, N_sig, N_max, N_min, SynTypes, Syngmax, Syndelay, Synpfail, STSP_prob, STSP_types,
 pCon, cluster_flag, S_sig, S_max, S_min, p_fail, E1, E2, E3, I1, I2, I3, STSP_setup, get_SynType, get_STSP_prob,
 get_STSP_types, STypPar, ConParStripes) = out
percent = np.asarray(big[0]) / 1e4                       # NTypes = ceil(N1 * percent / 100)
assert np.array_equal(np.ceil(1000 * percent / 100), NTypes)
pConStripes, coefStripes, interStripes = ConParStripes
data = dict(
    percent=percent, CellPar=np.asarray(CellPar), V0Par=np.asarray(V0Par), k_trans=np.asarray(k_trans),
    ParCov=np.stack([np.asarray(m) for m in ParCov]), N_sig=np.asarray(N_sig), N_max=np.asarray(N_max), N_min=np.asarray(N_min),
    SynTypes=np.asarray(SynTypes), Syngmax=np.asarray(Syngmax), Syndelay=np.asarray(Syndelay), Synpfail=np.asarray(Synpfail),
    STSP_prob=np.asarray(STSP_prob), STSP_types=np.asarray(STSP_types), pCon=np.asarray(pCon), cluster_flag=np.asarray(cluster_flag),
    S_sig=np.asarray(S_sig), S_max=np.asarray(S_max), S_min=np.asarray(S_min), p_fail=p_fail,
    STSP_names=np.asarray(sorted(STSP_setup.keys())), STSP_tables=np.stack([STSP_setup[k] for k in sorted(STSP_setup.keys())]),
    SynType_keys=np.asarray(sorted(get_SynType.keys())), SynType_vals=np.asarray([get_SynType[k] + [0] * (2 - len(get_SynType[k])) for k in sorted(get_SynType.keys())]),
    STSP_prob_keys=np.asarray(sorted(get_STSP_prob.keys())),
    STSP_prob_vals=np.asarray([list(get_STSP_prob[k]) + [np.nan] * (3 - len(get_STSP_prob[k])) for k in sorted(get_STSP_prob.keys())]),
    STSP_types_keys=np.asarray(sorted(get_STSP_types.keys())),
    STSP_types_vals=np.asarray([list(get_STSP_types[k]) + [""] * (3 - len(get_STSP_types[k])) for k in sorted(get_STSP_types.keys())]),
    STypPar=np.asarray(STypPar), pConStripes=np.asarray(pConStripes), coefStripes=np.asarray(coefStripes),
    interStripes_flat=np.concatenate([np.asarray(x) for x in interStripes]), interStripes_len=np.asarray([len(x) for x in interStripes]),
    # PSP amplitude -> peak conductance factors per target group (get_gmax, codes/ParamSetup.py:603-613), read off by calling it with PSP = 1
    gmax_per_psp_from_PC=np.asarray([get_gmax(1.0, i, 0) for i in range(14)]), gmax_per_psp_from_IN=np.asarray([get_gmax(1.0, i, 1) for i in range(14)]),
)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rescience_hass_hertag_durstewitz_2016_b200", "data", "pfc_constants.npz")
np.savez_compressed(path, **data)
print("wrote", path, os.path.getsize(path), "bytes")
