"""`Network` / `network_setup` of codes_revision/NetworkSetup.py (:15-19, 867-1060) without xarray.

The container keeps the revision's field names and orientation: `syn_pairs[0]` targets, `syn_pairs[1]` sources,
`group_distr[stripe][group]` (index order swapped w.r.t. codes/), `membr_params` with the 9 names of
codes_revision/BasicsSetup.py:20-32, `syn_params` = {'channel', 'spiking' (gmax, pfail), 'STSP_params' (U, tau_rec,
tau_fac), 'delay'}.  `network_setup` draws the network with the codes/ generator (our bit-exact port): the revision's
own generator differs numerically from codes/ through four slips the survey lists (F8: tau_fac loaded from tau_rec, one
pCon entry left 0, Mg factor 1, last_spike0 = -1000 ms) and needs xarray; the two that are *model data* — Mg factor and
last_spike0 — are options of `Cortex`, the other two would change the drawn network and are not reproduced.
"""
import os

import numpy as np

GROUP_NAMES = ["PC_L23", "IN_L_L23", "IN_L_d_L23", "IN_CL_L23", "IN_CL_AC_L23", "IN_CC_L23", "IN_F_L23",
               "PC_L5", "IN_L_L5", "IN_L_d_L5", "IN_CL_L5", "IN_CL_AC_L5", "IN_CC_L5", "IN_F_L5"]
GROUPS_SETS = {
    "ALL": GROUP_NAMES, "L23": GROUP_NAMES[:7], "L5": GROUP_NAMES[7:], "PC": ["PC_L23", "PC_L5"],
    "IN": GROUP_NAMES[1:7] + GROUP_NAMES[8:], "IN_L23": GROUP_NAMES[1:7], "IN_L5": GROUP_NAMES[8:],
    "IN_L_both": ["IN_L_L23", "IN_L_d_L23", "IN_L_L5", "IN_L_d_L5"],
    "IN_CL_both": ["IN_CL_L23", "IN_CL_AC_L23", "IN_CL_L5", "IN_CL_AC_L5"],
    "IN_L_both_L23": ["IN_L_L23", "IN_L_d_L23"], "IN_L_both_L5": ["IN_L_L5", "IN_L_d_L5"],
    "IN_CL_both_L23": ["IN_CL_L23", "IN_CL_AC_L23"], "IN_CL_both_L5": ["IN_CL_L5", "IN_CL_AC_L5"],
    "IN_L": ["IN_L_L23", "IN_L_L5"], "IN_L_d": ["IN_L_d_L23", "IN_L_d_L5"], "IN_CL": ["IN_CL_L23", "IN_CL_L5"],
    "IN_CL_AC": ["IN_CL_AC_L23", "IN_CL_AC_L5"], "IN_CC": ["IN_CC_L23", "IN_CC_L5"], "IN_F": ["IN_F_L23", "IN_F_L5"],
}
GROUPS_SETS.update({g: [g] for g in GROUP_NAMES})
MEMB_PARAM_NAMES = ["C", "g_L", "E_L", "delta_T", "V_up", "tau_w", "b", "V_r", "V_T"]
CHANNEL_NAMES = ["AMPA", "GABA", "NMDA"]


class Labelled:
    """A 2-D array with named rows: the subset of xarray.DataArray the revision's consumers use
    (`.loc[dict(param=name)]`, `.loc[dict(param=name, cell_index=idc)]`, `.values`, `.coords`)."""

    def __init__(self, values, names, second="cell_index"):
        self.values = np.asarray(values)
        self.names = list(names)
        self.second = second
        self.coords = {"param": np.asarray(self.names), second: np.arange(self.values.shape[1])}

    class _Loc:
        def __init__(self, o):
            self.o = o

        def __getitem__(self, sel):
            o = self.o
            rows = sel.get("param", slice(None))
            if isinstance(rows, str):
                rows = o.names.index(rows)
            elif not isinstance(rows, slice):
                rows = [o.names.index(r) for r in rows]
            cols = sel.get(o.second, slice(None))
            out = o.values[rows][..., cols] if not isinstance(rows, list) else o.values[rows][:, cols]
            return _Val(out)

    @property
    def loc(self):
        return Labelled._Loc(self)


class _Val:
    def __init__(self, v):
        self.values = np.asarray(v)

    def __array__(self, dtype=None):
        return np.asarray(self.values, dtype=dtype)


class _Holder:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Network:
    def __init__(self, membr_params, refractory_current, syn_params, syn_pairs, group_distr, STypPar, seed=None):
        self.membr_params = membr_params
        self.refractory_current = refractory_current
        self.syn_params = syn_params
        self.syn_pairs = np.asarray(syn_pairs, dtype=np.int64)
        self.group_distr = group_distr                       # [stripe][group] -> list of neuron ids
        self.STypPar = np.asarray(STypPar, dtype=np.float64)
        self.seed = seed
        n = membr_params.values.shape[1]
        nstripes = len(group_distr)
        self.basics = _Holder(
            struct=_Holder(Ncells=n // max(nstripes, 1), Ncells_total=n, Nstripes=nstripes, groups=_Holder(names=GROUP_NAMES, sets=GROUPS_SETS)),
            syn=_Holder(channels=_Holder(names=CHANNEL_NAMES)),
            membr=_Holder(names=MEMB_PARAM_NAMES))

    # ---- construction ----------------------------------------------------------------------------------------------
    @classmethod
    def from_codes(cls, NeuPar, V0, STypPar, SynPar, target_arr, source_arr, group_distr, seed=None):
        """From the arrays codes/NetworkSetup returns (NeuPar rows C,g_L,E_L,delta_T,V_up,tau_w,a,b,V_r,V_T,I_ref,V_dep)."""
        NeuPar = np.asarray(NeuPar, dtype=np.float64)
        SynPar = np.asarray(SynPar, dtype=np.float64)
        membr = Labelled(NeuPar[[0, 1, 2, 3, 4, 5, 7, 8, 9]], MEMB_PARAM_NAMES)
        iref = _Val(NeuPar[10])
        nsyn = SynPar.shape[1]
        syn = {"channel": Labelled((SynPar[0:1] - 1).astype(int) if nsyn else np.zeros((1, 0), dtype=int), ["channel"], "syn_index"),
               "spiking": Labelled(SynPar[[4, 6]] if nsyn else np.zeros((2, 0)), ["gmax", "pfail"], "syn_index"),
               "STSP_params": Labelled(SynPar[[1, 2, 3]] if nsyn else np.zeros((3, 0)), ["U", "tau_rec", "tau_fac"], "syn_index"),
               "delay": Labelled(SynPar[5:6] if nsyn else np.zeros((1, 0)), ["delay"], "syn_index")}
        nstripes = len(group_distr[0])
        gd = [[list(map(int, group_distr[g][s])) for g in range(len(group_distr))] for s in range(nstripes)]
        return cls(membr, iref, syn, np.stack([np.asarray(target_arr), np.asarray(source_arr)]), gd, STypPar, seed)

    # ---- queries (codes_revision/NetworkSetup.py:978-1060) -----------------------------------------------------------
    def group_idcs(self, group):
        if isinstance(group, (int, np.integer)):
            return [int(group)]
        return [GROUP_NAMES.index(g) for g in GROUPS_SETS[group]]

    def neuron_idcs(self, groupstripe_list):
        if isinstance(groupstripe_list[0], (str, int)):
            groupstripe_list = [groupstripe_list]
        out = []
        for group, stripe in groupstripe_list:
            for g in self.group_idcs(group):
                out.extend(self.group_distr[stripe][g])
        return np.array(out, dtype=np.int64)

    def syn_idcs_from_neurons(self, target, source, channel=None):
        hit = np.isin(self.syn_pairs[0], np.array(target)) & np.isin(self.syn_pairs[1], np.array(source))
        if channel is not None:
            channel = [channel] if isinstance(channel, str) else channel
            ch = self.syn_params["channel"].values[0]
            hit &= np.isin(ch, [CHANNEL_NAMES.index(c) for c in channel])
        return np.arange(self.syn_pairs.shape[1])[hit]

    def syn_idcs_from_groups(self, target_groupstripe_list, source_groupstripe_list, channel=None):
        return self.syn_idcs_from_neurons(self.neuron_idcs(target_groupstripe_list), self.neuron_idcs(source_groupstripe_list), channel)

    def rheobase(self, neuron_idcs=None):
        """I_SN = g_L (V_T - E_L - delta_T), pA (content.tex:130)."""
        P = self.membr_params.values
        r = P[1] * (P[8] - P[2] - P[3])
        return r if neuron_idcs is None else r[neuron_idcs]

    # ---- persistence: one .npz instead of netCDF + npy with Windows paths (codes_revision/NetworkSetup.py:886-976) ----
    def save(self, path):
        os.makedirs(path, exist_ok=True)
        flat = [x for st in self.group_distr for g in st for x in g]
        offs = np.cumsum([0] + [len(g) for st in self.group_distr for g in st])
        np.savez_compressed(os.path.join(path, "network.npz"), membr=self.membr_params.values, iref=self.refractory_current.values,
                            channel=self.syn_params["channel"].values, spiking=self.syn_params["spiking"].values,
                            stsp=self.syn_params["STSP_params"].values, delay=self.syn_params["delay"].values,
                            syn_pairs=self.syn_pairs, STypPar=self.STypPar, group_flat=np.asarray(flat), group_offs=offs,
                            nstripes=len(self.group_distr), seed=-1 if self.seed is None else self.seed)

    @classmethod
    def load(cls, path):
        d = np.load(os.path.join(path, "network.npz"))
        ns = int(d["nstripes"])
        offs, flat = d["group_offs"], d["group_flat"]
        ng = (len(offs) - 1) // ns
        gd = [[list(map(int, flat[offs[s * ng + g]:offs[s * ng + g + 1]])) for g in range(ng)] for s in range(ns)]
        syn = {"channel": Labelled(d["channel"], ["channel"], "syn_index"), "spiking": Labelled(d["spiking"], ["gmax", "pfail"], "syn_index"),
               "STSP_params": Labelled(d["stsp"], ["U", "tau_rec", "tau_fac"], "syn_index"), "delay": Labelled(d["delay"], ["delay"], "syn_index")}
        seed = int(d["seed"])
        return cls(Labelled(d["membr"], MEMB_PARAM_NAMES), _Val(d["iref"]), syn, d["syn_pairs"], gd, d["STypPar"], None if seed < 0 else seed)


def network_setup(Ncells, Nstripes, basics_scales=None, seed=None, workers=None):
    """codes_revision/NetworkSetup.py:15-19: seed numpy, draw the network."""
    import contextlib
    import io
    import warnings
    from ..codes.NetworkSetup import sparse_network
    if basics_scales:
        raise NotImplementedError("basics_scales (xarray selectors, codes_revision/BasicsSetup.py:803-821) are not supported; "
                                  "scale the arrays of the returned Network instead")
    np.random.seed(seed)
    scales = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        NeuPar, V0, STypPar, SynPar, (t, s), gd = sparse_network(Ncells, Nstripes, scales, True, workers=workers)
    return Network.from_codes(NeuPar, V0, STypPar, SynPar, t, s, gd, seed)
