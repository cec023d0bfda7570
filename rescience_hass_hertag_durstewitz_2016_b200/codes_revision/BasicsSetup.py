"""`basics_setup` / `BasicsSetup` of codes_revision/BasicsSetup.py (:803-853, :873-1060) without xarray and brian2.

The reference assembles ~25 labelled tables (group x group, parameter x group, ...) as xarray.DataArray objects and a
tree of dataclasses around them; consumers touch them through `.loc[...]`, `.values` and `.coords`.  Here the tables are
`Table` objects (a small labelled n-d array on numpy with exactly those three access paths, orthogonal label indexing,
`.loc[...] = ...`) and the tree is the same tree of attribute holders, so `basics.struct.conn.pCon.loc['PC_L5', 'PC_L23']`,
`basics.syn.spiking.params.gmax['mean']`, `basics.equations.syn_pathway[p]['eq']`, `basics_scales` selectors
(`{'gmax_mean': [(dict(target=[...], source=[...]), 2.0)]}`, :819-823) work as in the reference.

Numbers: the model constants are DATA of the published model; they are loaded from data/pfc_constants.npz (dumped from the
reference's codes/ParamSetup.py) and turned into the revision's tables by the revision's own, documented differences
(SURVEY F8), each applied explicitly below:
  * BasicsSetup.py:427-428 assigns pCon[PC_L5 <- IN_CC_L23] twice (0.2130, then 0.7006) and never pCon[PC_L5 <- IN_CC_L5],
    which stays 0; :637-638 does the same to the gmax sigma table;
  * :591 NMDA Mg_fac = 1 (codes/: 0.33);  the Mg block multiplies every channel, with Mg_fac = 0 for AMPA / GABA;
  * :447 STSP_PARAMS loads tau_fac from the 'tau_rec' row;
  * CortexSetup.py:254 last_spike = -1000 ms (codes/: -1e8 ms); the block term relaxes to V_r (:737; codes/: V_dep = V_r).
tests/test_basics_setup.py checks every table against tests/golden/rev_basics.npz, which the reference's own module
produced (tests/golden/make_golden_revision.py).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "pfc_constants.npz")


# ---- labelled arrays -------------------------------------------------------------------------------------------------
class Table:
    """n-d array with named dimensions and labelled coordinates: `.values`, `.dims`, `.coords[dim]`, `.loc[...]`
    (dict of dim -> label | list of labels, or a tuple of labels in dimension order), `.copy()`.  Lists select
    orthogonally (one list per dimension -> the outer product), as xarray does."""

    def __init__(self, values, dims, coords, name=None):
        self.values = np.array(values)
        self.dims = [dims] if isinstance(dims, str) else list(dims)
        self.coords = {d: np.asarray(c) for d, c in zip(self.dims, coords)}
        self.name = name
        assert self.values.ndim == len(self.dims), (self.values.shape, self.dims)
        self._index = {d: {(c.item() if hasattr(c, "item") else c): i for i, c in enumerate(self.coords[d])} for d in self.dims}

    # -- selection
    def _pos(self, dim, label):
        idx = self._index[dim]
        if isinstance(label, slice):
            return label
        if isinstance(label, Table):
            label = label.values
        if np.ndim(label) == 0:
            return idx[label.item() if hasattr(label, "item") else label]
        return np.array([idx[x.item() if hasattr(x, "item") else x] for x in label], dtype=np.int64)

    def _resolve(self, key):
        if not isinstance(key, dict):
            key = dict(zip(self.dims, key if isinstance(key, tuple) else (key,)))
        sel, keep = [], []
        for d in self.dims:
            p = self._pos(d, key[d]) if d in key else slice(None)
            if isinstance(p, slice):
                p = np.arange(len(self.coords[d]))[p]
            if np.ndim(p):
                keep.append((d, self.coords[d][p]))
            sel.append(p)
        vec = [p for p in sel if np.ndim(p)]
        mesh = iter(np.ix_(*vec)) if vec else iter(())
        return tuple(next(mesh) if np.ndim(p) else p for p in sel), keep

    class _Loc:
        def __init__(self, table):
            self.t = table

        def __getitem__(self, key):
            index, keep = self.t._resolve(key)
            return Table(self.t.values[index], [d for d, _ in keep], [c for _, c in keep], self.t.name)

        def __setitem__(self, key, value):
            index, keep = self.t._resolve(key)
            v = np.asarray(value.values if isinstance(value, Table) else value)
            shape = tuple(len(c) for _, c in keep)
            self.t.values[index] = v.reshape(shape) if v.ndim and v.size == int(np.prod(shape)) else v

    @property
    def loc(self):
        return Table._Loc(self)

    # -- array protocol
    @property
    def shape(self):
        return self.values.shape

    def copy(self):
        return Table(self.values.copy(), self.dims, [self.coords[d].copy() for d in self.dims], self.name)

    def to_numpy(self):
        return self.values

    def astype(self, t):
        return Table(self.values.astype(t), self.dims, [self.coords[d] for d in self.dims], self.name)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __float__(self):
        return float(self.values)

    def __int__(self):
        return int(self.values)

    def __len__(self):
        return len(self.values)

    def __eq__(self, other):
        return self.values == (other.values if isinstance(other, Table) else other)

    __hash__ = object.__hash__

    def __repr__(self):
        return "Table(%s, dims=%s)\n%r" % (self.name, self.dims, self.values)


class BaseClass:
    """The reference's BaseClass (:873-927): attribute holder with item access, keys/values and a printed tree."""

    def __init__(self, **fields):
        self.__dict__.update(fields)

    def keys(self):
        return list(self.__dict__.keys())

    def values(self):
        return list(self.__dict__.values())

    def __getitem__(self, item):
        return getattr(self, item)

    def __setitem__(self, item, value):
        setattr(self, item, value)

    def __contains__(self, item):
        return item in self.__dict__

    def tree(self, max_depth=None, item=None, _depth=0):
        node = self if item is None else self[item]
        if _depth == 0:
            print("%s instance containing:" % type(node).__name__ if node.keys() else type(node).__name__)
        for k in node.keys():
            child = node[k]
            deeper = isinstance(child, BaseClass) and (max_depth is None or max_depth > _depth)
            print("    |" * _depth + "    |- %s: %s%s" % (k, type(child).__name__, " instance containing:" if deeper else ""))
            if deeper:
                child.tree(max_depth, None, _depth + 1)

    def view(self, item=None):
        self.tree(0, item)


def _holder(name):
    return type(name, (BaseClass,), {})


ChannelSetup, SynapseSetup, SpikeParamSetup, DelaySetup, StripeSetup, ConnectionSetup, STSPSetup, MembraneSetup, GroupSetup, \
    StructureSetup, ParamSetup, EquationsSetup, BasicsSetup = (_holder(n) for n in (
        "ChannelSetup", "SynapseSetup", "SpikeParamSetup", "DelaySetup", "StripeSetup", "ConnectionSetup", "STSPSetup",
        "MembraneSetup", "GroupSetup", "StructureSetup", "ParamSetup", "EquationsSetup", "BasicsSetup"))


def setup_view(setup):
    setup.view() if isinstance(setup, BaseClass) else print("Input is %s, not BaseClass.\n%s" % (type(setup).__name__, setup))


def setup_tree(setup):
    setup.tree() if isinstance(setup, BaseClass) else print("Input is %s, not BaseClass.\n%s" % (type(setup).__name__, setup))


# ---- names -----------------------------------------------------------------------------------------------------------
_LAYERS = ("L23", "L5")
_TYPES = ("PC", "IN_L", "IN_L_d", "IN_CL", "IN_CL_AC", "IN_CC", "IN_F")
GROUP_NAMES = ["%s_%s" % (t, l) for l in _LAYERS for t in _TYPES]
GROUP_KINDS = {g: ("exc" if g.startswith("PC") else "inh") for g in GROUP_NAMES}


def _group_sets():
    in_types = _TYPES[1:]
    sets = {"ALL": list(GROUP_NAMES)}
    for l in _LAYERS:
        sets[l] = [g for g in GROUP_NAMES if g.endswith("_" + l)]
    sets["PC"] = ["PC_L23", "PC_L5"]
    sets["IN"] = [g for g in GROUP_NAMES if g.startswith("IN")]
    for l in _LAYERS:
        sets["IN_" + l] = [g for g in sets["IN"] if g.endswith("_" + l)]
    sets["IN_L_both"] = ["IN_L_L23", "IN_L_d_L23", "IN_L_L5", "IN_L_d_L5"]
    sets["IN_CL_both"] = ["IN_CL_L23", "IN_CL_AC_L23", "IN_CL_L5", "IN_CL_AC_L5"]
    for l in _LAYERS:
        sets["IN_L_both_" + l] = ["IN_L_" + l, "IN_L_d_" + l]
        sets["IN_CL_both_" + l] = ["IN_CL_" + l, "IN_CL_AC_" + l]
    for t in in_types:
        sets[t] = ["%s_%s" % (t, l) for l in _LAYERS]
    for g in GROUP_NAMES:
        sets[g] = [g]
    return sets


GROUPS_SETS = _group_sets()

MEMB_PARAM_UNITS = dict(C=dict(unit="pF", value="C"), g_L=dict(unit="nS", value="g_L"), E_L=dict(unit="mV", value="E_L"),
                        delta_T=dict(unit="mV", value="delta_T"), V_up=dict(unit="mV", value="V_up"),
                        tau_w=dict(unit="ms", value="tau_w"), b=dict(unit="pA", value="b"), V_r=dict(unit="mV", value="V_r"),
                        V_T=dict(unit="mV", value="V_T"))
MEMB_PARAM_NAMES = list(MEMB_PARAM_UNITS)
UNITS_DICT = {"pF": 1e-12, "nS": 1e-9, "mV": 1e-3, "ms": 1e-3, "pA": 1e-12, 1: 1}          # SI floats stand in for Brian units
UNIT_MAIN_DICT = {"pF": "farad", "nS": "siemens", "mV": "volt", "ms": "second", "pA": "amp", 1: 1}

CHANNEL_KINDS = {"AMPA": "exc", "GABA": "inh", "NMDA": "exc"}
CHANNEL_NAMES = list(CHANNEL_KINDS)
CHANNEL_PARAMS = ["tau_on", "tau_off", "E", "Mg_fac", "Mg_slope", "Mg_half"]
CHANNEL_PARAMS_UNITS = dict(tau_on=("ms", "second"), tau_off=("ms", "second"), E=("mV", "volt"), Mg_fac=("1", "1"),
                            Mg_slope=("1", "1"), Mg_half=("1", "1"))

# the revision loads tau_fac from the tau_rec row (BasicsSetup.py:447) — kept as data, applied by Cortex on request
STSP_PARAMS = {"U": dict(unit=1, value="U"), "tau_rec": dict(unit="ms", value="tau_rec"), "tau_fac": dict(unit="ms", value="tau_rec")}
STSP_VARS = {"u_temp": dict(unit=1, value="U"), "u": dict(unit=1, value="U"), "R_temp": dict(unit=1, value=1),
             "R": dict(unit=1, value=1), "a_syn": dict(unit=1, value="U")}
STSP_DECL_VARS = {**STSP_PARAMS, **STSP_VARS}
STSP_PARAM_NAMES = list(STSP_PARAMS)
STSP_PARAM_DISTR = ["%s_%s" % (v, s) for s in ("mean", "std") for v in STSP_PARAM_NAMES]
SYN_PARAMS = {"gmax": dict(unit="nS", value="gmax"), "pfail": dict(unit=1, value="pfail")}
SYN_PARAM_NAMES = list(SYN_PARAMS)
SYN_AUX = {"failure": dict(unit=1), "gsyn_amp": dict(unit=1)}

# share of each short-term-plasticity kind (E1..E3, I1..I3) per projection class (Hass et al. 2016, Table 4; :481-488)
STSP_SETS_DICT = {"A": (0.45, 0.38, 0.17, None, None, None), "B": (1, None, None, None, None, None),
                  "C": (None, 1, None, None, None, None), "D": (None, None, None, 0.25, 0.5, 0.25),
                  "E": (None, None, None, None, 1, None), "F": (None, None, None, 0.29, 0.58, 0.13)}

INTERSTRIPE_SETS = {
    "A": {"pair": ["PC_L23", "PC_L23"], "coefficient_0": 0.909, "coefficient_1": 1.4587, "connections": [-4, -2, 2, 4]},
    "B": {"pair": ["PC_L5", "IN_CC_L5"], "coefficient_0": 0.909, "coefficient_1": 1.0375, "connections": [-1, 1]},
    "C": {"pair": ["PC_L5", "IN_F_L5"], "coefficient_0": 0.909, "coefficient_1": 1.0375, "connections": [-1, 1]},
}


# ---- tables ----------------------------------------------------------------------------------------------------------
def _gg(values, name=None):
    return Table(values, ["target", "source"], [GROUP_NAMES, GROUP_NAMES], name)


def _pg(values, name=None):
    return Table(values, ["param", "group"], [MEMB_PARAM_NAMES, GROUP_NAMES], name)


def _stsp_distr(c):
    """Letter of the STSP set of every projection, from codes/'s numeric coding (STSP_prob / STSP_types look-up tables)."""
    kinds = list(c["STSP_names"])
    prob = {int(k): v for k, v in zip(c["STSP_prob_keys"], c["STSP_prob_vals"])}
    types = {int(k): v for k, v in zip(c["STSP_types_keys"], c["STSP_types_vals"])}
    out = np.full((14, 14), "", dtype="U16")
    for t in range(14):
        for s in range(14):
            kp, kt = int(c["STSP_prob"][t, s]), int(c["STSP_types"][t, s])
            if kp not in prob or kt not in types or not c["pCon"][t, s] and not (t == 7 and s == 12):
                continue
            row = [None] * 6
            for name, p in zip(types[kt], prob[kp]):
                if name:
                    row[kinds.index(name)] = float(p)
            hit = [k for k, v in STSP_SETS_DICT.items() if all((a is None) == (b is None) and (a is None or abs(a - b) < 1e-12) for a, b in zip(v, row))]
            out[t, s] = hit[0] if hit else ""
    return out


def _load_tables():
    c = np.load(_DATA, allow_pickle=False)
    T = {}
    T["PCELLS_PER_GROUP"] = c["percent"].copy()
    T["CELL_MEAN"] = _pg(c["CellPar"][1:10], "Memb_param mean")
    T["CELL_STD"] = _pg(c["N_sig"][1:10], "Memb_param std")
    rows = [1, 2, 3, 4, 5, 6, 8, 9, 10]                      # codes/ keeps `a` (unused) in row 7 of the min / max tables
    T["CELL_MINIMUM"] = _pg(c["N_min"][rows], "Memb_param min")
    T["CELL_MAXIMUM"] = _pg(c["N_max"][rows], "Memb_param max")
    T["K_TRANSF"] = _pg(c["k_trans"], "Memb_param K")
    T["CELL_COVARIANCE"] = Table(c["ParCov"], ["group", "memb_par0", "memb_par1"], [GROUP_NAMES, MEMB_PARAM_NAMES, MEMB_PARAM_NAMES], "Memb_param covariance")
    T["TAU_MIN"] = Table(c["N_min"][11], "group", [GROUP_NAMES], "Memb_param tau_m min")
    T["TAU_MAX"] = Table(c["N_max"][11], "group", [GROUP_NAMES], "Memb_param tau_m max")

    pcon, gsig = c["pCon"].copy(), c["S_sig"][:, :, 0].copy()
    t, s_ok, s_typo = GROUP_NAMES.index("PC_L5"), GROUP_NAMES.index("IN_CC_L5"), GROUP_NAMES.index("IN_CC_L23")
    for tab in (pcon, gsig):                                  # F8: :427-428 and :637-638 (see module docstring)
        tab[t, s_typo] = tab[t, s_ok]
        tab[t, s_ok] = 0.0
    T["PCON"] = _gg(pcon, "pCon")
    T["CLUSTER_FLAG"] = _gg(c["cluster_flag"], "Clustering flag")
    T["SYN_GMAX"] = _gg(c["Syngmax"], "Syn gmax")
    T["GMAX_SIGMA"], T["GMAX_MIN"], T["GMAX_MAX"] = _gg(gsig), _gg(c["S_min"][:, :, 0]), _gg(c["S_max"][:, :, 0])
    T["SYN_DELAY"] = _gg(c["Syndelay"], "Syn delay")
    T["DELAY_SIGMA"], T["DELAY_MIN"], T["DELAY_MAX"] = _gg(c["S_sig"][:, :, 1]), _gg(c["S_min"][:, :, 1]), _gg(c["S_max"][:, :, 1])
    T["SYN_PFAIL"] = _gg(c["Synpfail"], "Syn pfail")
    T["SYN_KINDS"] = _gg(np.array([[GROUP_KINDS[s] for s in GROUP_NAMES] for _ in GROUP_NAMES]), "Syn kinds")

    kinds = [str(k) for k in c["STSP_names"]]
    T["STSP_KINDS"] = Table(c["STSP_tables"].reshape(len(kinds), 6), ["kind", "param"], [kinds, STSP_PARAM_DISTR], "STSP param sets")
    sets = np.array([[np.nan if v is None else v for v in row] for row in STSP_SETS_DICT.values()], dtype=np.float64)
    T["STSP_SETS"] = Table(sets, ["set", "kind"], [list(STSP_SETS_DICT), kinds], "STSP set distribution")
    T["STSP_SET_DISTR"] = _gg(_stsp_distr(c), "STSP distribution groups")

    S = c["STypPar"]
    chan = np.stack([S[1], S[2], S[3], S[5], S[6], S[7]], axis=1)      # [channel][tau_on, tau_off, E, Mg_fac, Mg_slope, Mg_half]
    chan[CHANNEL_NAMES.index("NMDA"), CHANNEL_PARAMS.index("Mg_fac")] = 1.0          # F8: :591
    T["PARAM_CHANNELS"] = Table(chan, ["channel", "param"], [CHANNEL_NAMES, CHANNEL_PARAMS])
    T["GSYN_FACTOR_CHANNELS"] = Table(S[0].reshape(3, 1), ["channel", "param"], [CHANNEL_NAMES, ["factor"]])
    return T


_T = _load_tables()
globals().update(_T)                                          # PCON, SYN_GMAX, CELL_STD, ... as module attributes


# ---- equation strings (the model as data; :694-798) -------------------------------------------------------------------
def _per_channel(template):
    return "\n".join(template.format(n) for n in CHANNEL_NAMES)


EQ_MEMB_PARAM = "\n".join("%s: %s" % (n, UNIT_MAIN_DICT[MEMB_PARAM_UNITS[n]["unit"]]) for n in MEMB_PARAM_NAMES)
EQ_CHANNEL_PARAM = "\n".join("{}_{} = {}*{}: {}".format(p, k, float(_T["PARAM_CHANNELS"].loc[k, p].values), *CHANNEL_PARAMS_UNITS[p])
                             for p in CHANNEL_PARAMS for k in CHANNEL_NAMES)
EQ_AUX_VAR = "I_ref: amp\nlast_spike: second"
EQ_MG_FACTOR = _per_channel("Mg_{0} = (1/(1 +  Mg_fac_{0} * exp(Mg_slope_{0} * (Mg_half_{0}*mV - V)/mV))):1")
EQ_CHANNEL_CURRENT = _per_channel("I_{0} = g_{0} * (E_{0} - V) * Mg_{0}: amp")
EQ_CHANNEL_CONDUCTANCE = _per_channel("g_{0} = g_{0}_off - g_{0}_on: siemens")
EQ_CHANNEL_DGDT = _per_channel("dg_{0}_off/dt = - (1/tau_off_{0}) * g_{0}_off: siemens\ndg_{0}_on/dt = - (1/tau_on_{0}) * g_{0}_on: siemens")
EQ_CURRENT = ("I_DC: amp\nI_AC = {}*pA: amp\n" + "I_syn = " + " + ".join("I_" + n for n in CHANNEL_NAMES) + ": amp\n"
              + "I_inj = I_DC + I_AC: amp\nI_tot =  I_syn + I_inj: amp")
_BLOCK = "int(I_tot >= I_ref) * int(t - last_spike < 5 * ms)"
_W_BAND = "int(w > w_V - D0/tau_w) * int(w < w_V + D0/tau_w) * int(V <= V_T)"
EQ_MEMB = "\n".join([
    "I_exp = g_L * delta_T * exp((V - V_T)/delta_T): amp",
    "w_V = I_tot + I_exp -g_L * (V - E_L): amp",
    "dV = " + _BLOCK + " * (-g_L/C)*(V - V_r) + (1 - " + _BLOCK + ") * (I_tot + I_exp - g_L * (V - E_L) - w)/C: volt/second",
    "dV/dt = dV: volt",
    "",
    "D0 = (C/g_L) * w_V:  coulomb",
    "dD0 = C *(exp((V - V_T)/delta_T)-1): farad",
    "dw/dt = " + _W_BAND + " * int(I_tot < I_ref) * -(g_L * (1 - exp((V - V_T)/delta_T)) + dD0/tau_w)*dV: amp"])
EQ_MEMBR_MODEL = "\n\n".join([EQ_MEMB_PARAM, EQ_CHANNEL_PARAM, EQ_MG_FACTOR, EQ_CHANNEL_CURRENT, EQ_CHANNEL_CONDUCTANCE,
                              EQ_CHANNEL_DGDT, EQ_AUX_VAR, EQ_CURRENT, EQ_MEMB])
EQ_MEMBR_THRESHOLD = "V > V_up"
EQ_MEMBR_RESET = "V = V_r; w += b"
EQ_MEMBR_EVENT = {"w_crossing": {"condition": "w > w_V - D0/tau_w and w < w_V + D0/tau_w and V <= V_T", "vars": ["V", "w"],
                                 "reset": "w=w_V - D0/tau_w"}}


def IRHEOBASE(g_L, V_T, E_L, delta_T):
    return g_L * (V_T - E_L - delta_T)


def _decl(table):
    return "\n".join("%s: %s" % (v, UNIT_MAIN_DICT[table[v]["unit"]]) for v in table)


EQ_SYN_CHANNELS = "\n".join("%s: 1" % n for n in CHANNEL_NAMES)
EQ_SYN_MODEL = "\n\n".join([EQ_SYN_CHANNELS, _decl(STSP_DECL_VARS), _decl(SYN_PARAMS), _decl(SYN_AUX)])
EQ_EXT_SYN_MODEL = "\n\n".join([EQ_SYN_CHANNELS, _decl(SYN_PARAMS), _decl(SYN_AUX)])
EQ_SYN_PATHWAY = [dict(eq=e, order=i, delay=False) for i, e in enumerate([
    "failure = int(rand()<pfail)",
    "u_temp = U + u * (1 - U) * exp(-(t - last_spike_pre)/tau_fac)",
    "R_temp = 1 + (R - u * R - 1) * exp(- (t - last_spike_pre)/tau_rec)",
    "u = u_temp", "R = R_temp", "last_spike_pre = t", "a_syn = u * R"])]
EQ_EXT_SYN_PATHWAY = [dict(eq="failure = int(rand()<pfail)", order=0, delay=False)]
for _n in CHANNEL_NAMES:
    for _side in ("on", "off"):
        EQ_SYN_PATHWAY.append(dict(eq="g_{0}_{1}_post += {0} * gmax * a_syn * (1 - failure) * gsyn_amp".format(_n, _side), order=7, delay=True))
        EQ_EXT_SYN_PATHWAY.append(dict(eq="g_{0}_{1}_post += {0} * gmax {2}* (1 - failure) * gsyn_amp".format(_n, _side, " " if _side == "on" else ""), order=0, delay=False))
EQ_VAR_UNITS = dict(V="mV", w="pA", t="ms", I_tot="pA", I_syn="pA", I_inj="pA", I_DC="pA", I_AC="pA")
EQ_VAR_UNITS.update({"I_" + n: "pA" for n in CHANNEL_NAMES})
for _units in (MEMB_PARAM_UNITS, STSP_DECL_VARS, SYN_PARAMS, SYN_AUX):
    EQ_VAR_UNITS.update({v: _units[v]["unit"] for v in _units})


# ---- basics_setup (:803-853) -----------------------------------------------------------------------------------------
def basics_setup(Ncells_prompted, Nstripes, basics_scales=None, disp=True):
    """-> BasicsSetup(struct, membr, syn, equations, scalables, scales).  `basics_scales` = {'pCon' | 'membr_param_std' |
    'gmax_mean': [(selector dict, factor), ...]} multiplies the selected block of a COPY of the table (:819-823)."""
    pcon, cell_std, gmax = _T["PCON"].copy(), _T["CELL_STD"].copy(), _T["SYN_GMAX"].copy()
    scalables = {"pCon": pcon, "membr_param_std": cell_std, "gmax_mean": gmax}
    if basics_scales is not None:
        for param in basics_scales:
            for selector, scale in basics_scales[param]:
                scalables[param].loc[selector] = scalables[param].loc[selector].values * scale

    groups = GroupSetup(kinds=GROUP_KINDS, sets=GROUPS_SETS, names=list(GROUP_NAMES), N=len(GROUP_NAMES),
                        idcs={g: i for i, g in enumerate(GROUP_NAMES)})
    stripes = StripeSetup(N=Nstripes, inter=INTERSTRIPE_SETS)
    conn = ConnectionSetup(pCon=pcon, cluster=_T["CLUSTER_FLAG"])
    per_group = Table(np.ceil((Ncells_prompted * _T["PCELLS_PER_GROUP"]) / 100).astype(int), "group", [GROUP_NAMES])
    ncells = int(per_group.values.sum())
    struct = StructureSetup(Ncells_prompt=Ncells_prompted, Pcells_per_group=_T["PCELLS_PER_GROUP"], groups=groups, stripes=stripes,
                            conn=conn, Ncells=ncells, Ncells_total=Nstripes * ncells, Ncells_per_group=per_group)
    membr = MembraneSetup(names=list(MEMB_PARAM_NAMES), name_units=MEMB_PARAM_UNITS, unit_dict=UNITS_DICT, unit_main_dict=UNIT_MAIN_DICT,
                          mean=_T["CELL_MEAN"], covariance=_T["CELL_COVARIANCE"], k=_T["K_TRANSF"], std=cell_std,
                          min=_T["CELL_MINIMUM"], max=_T["CELL_MAXIMUM"], tau_m_min=_T["TAU_MIN"], tau_m_max=_T["TAU_MAX"])
    stsp = STSPSetup(decl_vars=STSP_DECL_VARS, kinds=_T["STSP_KINDS"], sets=_T["STSP_SETS"], distr=_T["STSP_SET_DISTR"])
    kinds_to_names = {"exc": [n for n in CHANNEL_NAMES if CHANNEL_KINDS[n] == "exc"], "inh": [n for n in CHANNEL_NAMES if CHANNEL_KINDS[n] == "inh"]}
    channels = ChannelSetup(kinds=CHANNEL_KINDS, params=_T["PARAM_CHANNELS"], gsyn_factor=_T["GSYN_FACTOR_CHANNELS"],
                            names=list(CHANNEL_NAMES), kinds_to_names=kinds_to_names)
    spike_params = ParamSetup(gmax=dict(mean=gmax, sigma=_T["GMAX_SIGMA"], min=_T["GMAX_MIN"], max=_T["GMAX_MAX"]), pfail=_T["SYN_PFAIL"])
    spiking = SpikeParamSetup(names=SYN_PARAMS, params=spike_params)
    delay = DelaySetup(delay_mean=_T["SYN_DELAY"], delay_sigma=_T["DELAY_SIGMA"], delay_min=_T["DELAY_MIN"], delay_max=_T["DELAY_MAX"])
    syn = SynapseSetup(kinds=_T["SYN_KINDS"], spiking=spiking, delay=delay, STSP=stsp, channels=channels)
    equations = EquationsSetup(membr_model=EQ_MEMBR_MODEL, membr_threshold=EQ_MEMBR_THRESHOLD, membr_reset=EQ_MEMBR_RESET,
                               membr_events=EQ_MEMBR_EVENT, rheobase=IRHEOBASE, syn_model=EQ_SYN_MODEL, syn_pathway=EQ_SYN_PATHWAY,
                               ext_syn_model=EQ_EXT_SYN_MODEL, ext_syn_pathway=EQ_EXT_SYN_PATHWAY, var_units=EQ_VAR_UNITS)
    if ncells != Ncells_prompted and disp:
        print("REPORT: The number of neurons was adjusted from {} to {} to preserve group distribution.\n".format(Ncells_prompted, ncells))
    return BasicsSetup(struct=struct, membr=membr, syn=syn, equations=equations, scalables=scalables, scales=basics_scales)


def engine_model(basics):
    """What the revision's equation set means for the engine (csrc/hhd_model.cuh evaluates the model natively; the strings
    above are carried as data): channel constants with the NMDA Mg factor of the revision, Gsyn amplitudes
    (CortexSetup.py:265-272), last_spike0 (:254) and whether tau_fac is loaded from tau_rec (BasicsSetup.py:447)."""
    P = basics.syn.channels.params
    col = lambda p: np.asarray(P.loc[dict(param=p)].values, dtype=np.float64)
    tau_on, tau_off = col("tau_on"), col("tau_off")
    nmda = CHANNEL_NAMES.index("NMDA")
    assert all(col("Mg_fac")[i] == 0.0 for i in range(3) if i != nmda), "only the NMDA channel carries a Mg block in the engine"
    factor = np.asarray(basics.syn.channels.gsyn_factor.values, dtype=np.float64)[:, 0]
    return dict(tau_on=tau_on, tau_off=tau_off, E_rev=col("E"), gsyn=factor * tau_off * tau_on / (tau_off - tau_on),
                mg_fac=float(col("Mg_fac")[nmda]), mg_slope=float(col("Mg_slope")[nmda]), mg_half=float(col("Mg_half")[nmda]),
                last_spike0_ms=-1000.0, tau_fac_from=basics.syn.STSP.decl_vars["tau_fac"]["value"])
