"""`Network` / `network_setup` of codes_revision/NetworkSetup.py (:15-170, 867-1060) without xarray.

`network_setup(Ncells, Nstripes, basics_scales, seed)` follows the revision's OWN generator (default, one stripe): the same
numpy draws in the same order as codes_revision/NetworkSetup.py:15-170 — multivariate-normal membrane parameters with
re-draws of the violators (:42-66, 306-333), I_ref by `fmin` on the closed-form onset rate (:84-87, 634-797), the
latency / adaptation re-sorting of the IN_L and IN_CL groups (:215-304), shuffled-index connectivity (:335-351), the
common-neighbour rule for PC->PC (:353-552) and per-synapse draws (:173-213, 560-614) — so the network is bit-identical to
the reference's for the same seed (tests/test_revision_generator.py against fixtures the reference's code produced, see
tests/golden/make_golden_revision.py).  That includes the revision's slips (SURVEY F8 and two more found while porting):
the PC_L5 <- IN_CC_L5 entry of pCon left at 0, and `delay_fac = 1` passed for intra-stripe projections, which ADDS 1 ms to
every mean delay and delay sigma (:100, :114, :186-187).  The revision's inter-stripe block calls SetCon_standard with
two integers and raises (:144 -> `len()` of an int), so the reference itself only runs with Nstripes = 1; for several
stripes pass generator="codes" (the codes/ generator, bit-exact port, sparse and parallel), which is also what
`Network.from_codes` wraps.

The container keeps the revision's field names and orientation: `syn_pairs[0]` targets, `syn_pairs[1]` sources,
`group_distr[stripe][group]`, `membr_params` with the 9 names of BasicsSetup.py:20-32, `syn_params` = {'channel', 'spiking'
(gmax, pfail), 'STSP_params' (U, tau_rec, tau_fac), 'delay'}, `basics` = the BasicsSetup tree.
"""
import os
from collections import namedtuple

import numpy as np
from scipy.integrate import quad
from scipy.optimize import fmin, fsolve

from .BasicsSetup import (BaseClass, CHANNEL_NAMES, GROUP_NAMES, GROUPS_SETS, MEMB_PARAM_NAMES, Table, basics_setup, engine_model)

MembraneTuple = namedtuple("MembraneTuple", MEMB_PARAM_NAMES)
SYN_FIELDS = {"channel": ["channel"], "spiking": ["gmax", "pfail"], "STSP_params": ["U", "tau_rec", "tau_fac"], "delay": ["delay"]}


def Labelled(values, names, second="cell_index"):
    """2-D table with named rows (membr_params: param x cell_index; syn_params[...]: param x syn_index)."""
    values = np.asarray(values)
    return Table(values, ["param", second], [list(names), np.arange(values.shape[1])])


class Network(BaseClass):
    """codes_revision/NetworkSetup.py:867-1060: the drawn network + index helpers."""

    def __init__(self, membr_params, refractory_current, syn_params, syn_pairs, group_distr, seed=None, basics=None, STypPar=None):
        self.membr_params = membr_params
        self.refractory_current = refractory_current
        self.syn_params = syn_params
        self.syn_pairs = np.asarray(syn_pairs, dtype=np.int64)
        self.group_distr = group_distr                       # [stripe][group] -> list of neuron ids
        self.seed = seed
        n = membr_params.values.shape[1]
        nstripes = max(len(group_distr), 1)
        if basics is None:                                   # arrays from elsewhere (codes/ generator, a saved file)
            basics = basics_setup(n // nstripes, nstripes, None, disp=False)
            basics.struct.Ncells, basics.struct.Ncells_total = n // nstripes, n
        basics.struct.Nstripes = nstripes                   # shorthand for basics.struct.stripes.N
        self.basics = basics
        # channel constants as the engine takes them: the codes/ table if the arrays came from codes/ (Mg factor 0.33),
        # else what the BasicsSetup tree says (revision: Mg factor 1)
        self.STypPar = None if STypPar is None else np.asarray(STypPar, dtype=np.float64)

    def channel_constants(self):
        """-> kwargs of Engine.set_channels (+ 'last_spike0_ms', 'tau_fac_from' for the revision's model variant)."""
        if self.STypPar is not None:
            from ..engine import channels_from_STypPar
            return dict(channels_from_STypPar(self.STypPar), last_spike0_ms=-1000.0, tau_fac_from="tau_fac")
        return engine_model(self.basics)

    # ---- construction ----------------------------------------------------------------------------------------------
    @classmethod
    def from_codes(cls, NeuPar, V0, STypPar, SynPar, target_arr, source_arr, group_distr, seed=None):
        """From the arrays codes/NetworkSetup returns (NeuPar rows C,g_L,E_L,delta_T,V_up,tau_w,a,b,V_r,V_T,I_ref,V_dep)."""
        NeuPar = np.asarray(NeuPar, dtype=np.float64)
        SynPar = np.asarray(SynPar, dtype=np.float64)
        membr = Labelled(NeuPar[[0, 1, 2, 3, 4, 5, 7, 8, 9]], MEMB_PARAM_NAMES)
        iref = Table(NeuPar[10], "cell_index", [np.arange(NeuPar.shape[1])])
        nsyn = SynPar.shape[1]
        syn = {"channel": Labelled((SynPar[0:1] - 1).astype(int) if nsyn else np.zeros((1, 0), dtype=int), ["channel"], "syn_index"),
               "spiking": Labelled(SynPar[[4, 6]] if nsyn else np.zeros((2, 0)), ["gmax", "pfail"], "syn_index"),
               "STSP_params": Labelled(SynPar[[1, 2, 3]] if nsyn else np.zeros((3, 0)), ["U", "tau_rec", "tau_fac"], "syn_index"),
               "delay": Labelled(SynPar[5:6] if nsyn else np.zeros((1, 0)), ["delay"], "syn_index")}
        nstripes = len(group_distr[0])
        gd = [[list(map(int, group_distr[g][s])) for g in range(len(group_distr))] for s in range(nstripes)]
        return cls(membr, iref, syn, np.stack([np.asarray(target_arr), np.asarray(source_arr)]), gd, seed, None, STypPar)

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
        """I_SN = g_L (V_T - E_L - delta_T), pA (content.tex:130; BasicsSetup.py IRHEOBASE)."""
        P = self.membr_params.values
        r = self.basics.equations.rheobase(P[1], P[8], P[2], P[3])
        return r if neuron_idcs is None else r[neuron_idcs]

    # ---- persistence: one .npz instead of netCDF + npy with Windows paths (codes_revision/NetworkSetup.py:886-976) ----
    def save(self, path):
        os.makedirs(path, exist_ok=True)
        flat = [x for st in self.group_distr for g in st for x in g]
        offs = np.cumsum([0] + [len(g) for st in self.group_distr for g in st])
        b = self.basics
        np.savez_compressed(os.path.join(path, "network.npz"), membr=self.membr_params.values, iref=self.refractory_current.values,
                            channel=self.syn_params["channel"].values, spiking=self.syn_params["spiking"].values,
                            stsp=self.syn_params["STSP_params"].values, delay=self.syn_params["delay"].values,
                            syn_pairs=self.syn_pairs, STypPar=np.zeros(0) if self.STypPar is None else self.STypPar,
                            group_flat=np.asarray(flat), group_offs=offs, nstripes=len(self.group_distr),
                            seed=-1 if self.seed is None else self.seed, Ncells_prompt=b.struct.Ncells_prompt,
                            scales=np.asarray(__import__("json").dumps(b.scales)))

    @classmethod
    def load(cls, path):
        d = np.load(os.path.join(path, "network.npz"))
        ns = int(d["nstripes"])
        offs, flat = d["group_offs"], d["group_flat"]
        ng = (len(offs) - 1) // ns
        gd = [[list(map(int, flat[offs[s * ng + g]:offs[s * ng + g + 1]])) for g in range(ng)] for s in range(ns)]
        syn = {k: Labelled(d[f], SYN_FIELDS[k], "syn_index") for k, f in (("channel", "channel"), ("spiking", "spiking"), ("STSP_params", "stsp"), ("delay", "delay"))}
        seed = int(d["seed"])
        scales = __import__("json").loads(str(d["scales"])) if "scales" in d.files else None
        if scales:
            scales = {k: [(sel, fac) for sel, fac in v] for k, v in scales.items()}
        basics = basics_setup(int(d["Ncells_prompt"]), ns, scales, disp=False) if "Ncells_prompt" in d.files else None
        iref = Table(d["iref"], "cell_index", [np.arange(len(d["iref"]))])
        return cls(Labelled(d["membr"], MEMB_PARAM_NAMES), iref, syn, d["syn_pairs"], gd, None if seed < 0 else seed, basics,
                   d["STypPar"] if d["STypPar"].size else None)


# ---- closed-form firing rate of the simplified AdEx (:634-797) -------------------------------------------------------
def get_FR(memb, I0, w_r, V_r):
    """Rate in Hz for constant current I0 from (w_r, V_r); None where the reference's function falls off its end."""
    C, g_L, E_L, delta_T, V_up, tau_w, b, _, V_T = memb

    def f13(V, I, w):                      # 1 / (dV/dt) while w stays put (parts 1 and 3 of the ISI)
        return 1 / ((1 / C) * (I - w + g_L * delta_T * np.exp((V - V_T) / delta_T) - g_L * (V - E_L)))

    def f2(V):                             # ... while w slides along the nullcline band (part 2)
        ratio = C / (tau_w * g_L)
        k0 = (ratio - 1) * g_L
        k1 = (1 - ratio) * I0 - k0 * E_L
        k2 = (1 - ratio) * g_L * delta_T
        e = np.exp((V - V_T) / delta_T)
        return 1 / ((1 / C) * (I0 - (k0 * V + k1 + k2 * e) + g_L * delta_T * np.exp((V - V_T) / delta_T) - g_L * (V - E_L)))

    def band_crossings(I, w0, w_ref, signal):
        if w0 > w_ref:
            ratio = C / (g_L * tau_w)
            bound = lambda V: (1 + signal * ratio) * (I - g_L * (V - E_L) + g_L * delta_T * np.exp((V - V_T) / delta_T)) - w0
            out = []
            for lo, hi in ((E_L + (I - (w0 / (1 + signal * ratio))) / g_L - 0.1, E_L + delta_T + (I - (w0 / (1 + signal * ratio))) / g_L),
                           (E_L + delta_T + (I - (w0 / (1 + signal * ratio))) / g_L, V_up)):
                for _ in range(1000):
                    if np.sign(bound(hi)) * np.sign(bound(lo)) == -1:
                        break
                    lo -= 1
                out.append(fsolve(bound, (lo + hi) / 2))
            return out
        return [V_T] if w0 == w_ref else []

    tau_ratio = (C / g_L) / tau_w
    Dist_VT = tau_ratio * (I0 + g_L * delta_T - g_L * (V_T - E_L))
    w_end = -g_L * (V_T - E_L) + g_L * delta_T + I0 - Dist_VT
    if Dist_VT <= 0 or tau_ratio >= 1 or C <= 0 or g_L <= 0 or tau_w <= 0 or delta_T <= 0:
        return None
    e_r = np.exp((V_r - V_T) / delta_T)
    dist_V_r = tau_ratio * (I0 + g_L * delta_T * e_r - g_L * (V_r - E_L))
    wV_V_r = -g_L * (V_r - E_L) + g_L * delta_T * e_r + I0
    w1, w2 = wV_V_r - dist_V_r, wV_V_r + dist_V_r
    Q = lambda f, lo, hi: quad(f, lo, hi)[0]
    t1 = t2 = 0
    if V_r >= V_T:
        if w_r >= wV_V_r:
            w_ref = -g_L * (V_T - E_L) + g_L * delta_T + I0 + Dist_VT
            X = band_crossings(I0, w_r, w_ref, 1)
            if len(X) < 2:
                t1 = Q(lambda x: f13(x, I0, w_r), V_r, V_T)
            else:
                Vc = np.min(X)
                t1 = Q(lambda x: f13(x, I0, w_r), V_r, Vc)
                t2 = Q(f2, Vc, V_T)
            w_stop, V1b = w_end, V_T
        else:
            V1b, w_stop = V_r, w_r
    else:
        if w1 < w_r < w2:
            t2 = Q(f2, V_r, V_T)
            w_stop = w_end
        else:
            if w_r <= w1:
                signal, w_ref = -1, -g_L * (V_T - E_L) + g_L * delta_T + I0 - Dist_VT
            else:
                signal, w_ref = 1, -g_L * (V_T - E_L) + g_L * delta_T + I0 + Dist_VT
            X = band_crossings(I0, w_r, w_ref, signal)
            if not X or len(X) == 1:
                t1 = Q(lambda x: f13(x, I0, w_r), V_r, V_T)
                w_stop = w_r
            else:
                Vb = np.min(X)
                t1 = Q(lambda x: f13(x, I0, w_r), V_r, Vb)
                t2 = Q(f2, Vb, V_T)
                w_stop = w_end
        V1b = V_T
    t3 = 0 if V1b >= V_up else Q(lambda x: f13(x, I0, w_stop), V1b, V_up)
    return 1000 / np.asarray(t1 + t2 + t3)


def get_statFR(memb, I0):
    tau_ratio = (memb.C / memb.g_L) / memb.tau_w
    Dist_VT = tau_ratio * (I0 + memb.g_L * memb.delta_T - memb.g_L * (memb.V_T - memb.E_L))
    w_end = -memb.g_L * (memb.V_T - memb.E_L) + memb.g_L * memb.delta_T + I0 - Dist_VT
    return get_FR(memb, I0, w_end + memb.b if memb.b != 0 else 0, memb.V_r)


def get_transFR(memb, I0, w0, V0):
    return get_FR(memb, I0, w0, V0)


def Define_I_ref(I, memb):
    re = get_transFR(memb, I, 0, memb.V_r)
    return ((0 if re is None else re) - 200) ** 2


def latency_AdEx(memb, I):
    return quad(lambda V: memb.C / (I - memb.g_L * (V - memb.E_L) + memb.g_L * memb.delta_T * np.exp((V - memb.V_T) / memb.delta_T)), memb.E_L, memb.V_T / 2)[0]


def latency_LIF(memb, I):
    return memb.C * np.log(I / (I + memb.g_L * (memb.E_L - memb.V_T / 2))) / memb.g_L


# ---- random draws (same numpy calls in the same order as the reference) ----------------------------------------------
def inv_transform(X_trans, k, mean_X, std_X, min_X):
    base = mean_X + 0.8 * std_X * X_trans
    X = base ** (1 / k) if k > 0 else np.exp(base)
    return X + 1.1 * min_X if min_X < 0 else X


def random_param(N, par_mean, par_std, par_min, par_max, distr_flag):
    if par_std == 0:
        return par_mean * np.ones(N) if par_max == 0 else par_min + (par_max - par_min) * np.random.random(size=N)
    if distr_flag == "uniform":
        return par_min + (par_max - par_min) * np.random.random(size=N)
    if distr_flag == "normal":
        par = par_mean + par_std * np.random.normal(size=N)
    else:                                                   # 'log_normal'
        par = np.exp(np.random.normal(size=N) * par_std + par_mean)
    out = np.where((par < par_min) | (par > par_max))[0]
    par[out] = par_min + (par_max - par_min) * np.random.random(size=len(out))
    return par


def get_TSpairs(conn_idc, Nsource):
    return np.asarray([conn_idc // Nsource, conn_idc % Nsource]).astype(int)


def SetCon(Nsyn, target_arr, source_arr, pCon):
    """A random subset of round(pCon Nt Ns) of the Nt x Ns pairs (positions within the two arrays)."""
    Nt, Ns = len(target_arr), len(source_arr)
    conn = np.arange(Nt * Ns)
    np.random.shuffle(conn)
    n = int(np.round(pCon * Nt * Ns, 0))
    return get_TSpairs(conn[:n], Ns), (np.arange(n) + Nsyn).astype(int)


def SetCon_standard(Nsyn, target_arr, source_arr, pCon):
    target_arr, source_arr = np.asarray(target_arr), np.asarray(source_arr)
    TS, idc = SetCon(Nsyn, target_arr, source_arr, pCon)
    return np.array([target_arr[TS[0]], source_arr[TS[1]]]), idc


def get_commonneigh_recur(TS, Ncell):
    """Common neighbours of every pair (connections in either direction count; :832-847), as one integer matrix product."""
    conn = np.zeros((Ncell, Ncell), dtype=np.int64)
    conn[TS[0], TS[1]] = 1
    conn = ((conn + conn.T) > 0).astype(np.int64)
    d = np.diag(conn)
    M = conn @ conn - conn * d[:, None] - conn * d[None, :]
    np.fill_diagonal(M, 0)
    return M


def _pairs(mask):
    return np.array(np.nonzero(mask))


def _both_ways(pairs, n):
    m = np.zeros((n, n), dtype=bool)
    m[pairs[0], pairs[1]] = True
    return _pairs(m | m.T)


def _neighbour_quota(Neigh_mat, TS, pCon, pSelfCon, slope):
    """How many connections pairs with n common neighbours get (p_calc_recur, :421-461)."""
    Ncell = Neigh_mat.shape[0]
    N_neigh = np.unique(Neigh_mat)
    count = np.array([np.sum(Neigh_mat == n) for n in N_neigh], dtype=np.float64)

    def p_min(N0):
        N1 = np.round(min(max(N_neigh), N0 + 1 / slope / pCon))
        mid = (N_neigh > N0) & (N_neigh <= N1)
        return abs(np.sum(count[mid] / np.sum(count) * slope * (N_neigh[mid] - N0)) + np.sum(count[N_neigh > N1] / np.sum(count)) - 1)

    offset = fmin(p_min, np.max(N_neigh), disp=False)
    N1 = np.floor(min(np.max(N_neigh), offset + 1 / slope / pCon))
    N0 = max(np.min(N_neigh), np.ceil(offset))
    ind = np.isin(N_neigh, np.arange(float(np.asarray(N0).reshape(-1)[0]), float(np.asarray(N1 + 1).reshape(-1)[0])))
    factor = np.zeros(len(N_neigh))
    factor[ind] = pCon * slope * (N_neigh[ind] - offset)
    if ind.any():
        factor[np.where(ind)[0][-1] + 1:] = 1
    else:
        factor[:] = 1
    N_original = Ncell ** 2 * pCon - Ncell * pSelfCon
    return N_neigh.astype(int), np.round(factor * N_original * count / (count @ factor)).astype(int)


def SetCon_CommonNeighbour(Nsyn, cells_arr, pCon, pSelfCon):
    """PC -> PC connectivity with the common-neighbour rule and a share pSelfCon of reciprocal pairs (:353-552).  Pair sets
    are boolean N x N masks; their enumeration order (row-major, what np.where / np.unique give the reference) fixes which
    pairs a shuffled index picks, so the same shuffles select the same pairs."""
    cells_arr = np.array(cells_arr)
    n = len(cells_arr)
    TS, _ = SetCon(Nsyn, cells_arr, cells_arr, pCon)
    Neigh = get_commonneigh_recur(TS, n)
    N_neigh, quota = _neighbour_quota(Neigh, TS, pCon, pSelfCon, 20 * 3.9991 / n)

    X = np.zeros((n, n), dtype=bool)
    X[TS[0], TS[1]] = True
    eye = np.eye(n, dtype=bool)
    off = X & ~eye
    recur = off & off.T                                   # connected both ways
    uni = off & ~off.T                                    # one way only, [target, source]
    free = ~(X | X.T | eye)                               # not connected in either direction
    lower = np.tril(np.ones((n, n), dtype=bool), -1)
    chosen = [np.zeros((2, 0), dtype=np.int64)]

    def pick(pairs, k):
        idc = np.arange(pairs.shape[1])
        np.random.shuffle(idc)
        return pairs[:, idc[:k]]

    for q in np.nonzero(quota > 0)[0]:
        here = Neigh == N_neigh[q]
        free_q, recur_q, uni_q = _pairs(free & here), _pairs(recur & here), _pairs(uni & here)
        free_low, recur_low = _pairs(free & here & lower), _pairs(recur & here & lower)
        n_recur = int(np.floor(quota[q] * pSelfCon / 2))
        n_uni = quota[q] - 2 * n_recur
        if recur_low.shape[1] >= n_recur:                 # enough reciprocal pairs: keep a random n_recur of them
            chosen.append(_both_ways(pick(recur_low, n_recur), n))
        else:                                             # keep all, make more out of unconnected pairs
            chosen.append(recur_q)
            new = _both_ways(pick(free_low, n_recur - recur_low.shape[1]), n)
            chosen.append(new)
            if free_q.shape[1] and new.shape[1]:
                m = free & here
                m[new[0], new[1]] = False
                free_q = _pairs(m)
        if uni_q.shape[1] > n_uni:
            chosen.append(pick(uni_q, n_uni))
        else:
            chosen.append(uni_q)
            chosen.append(pick(free_q, n_uni - uni_q.shape[1]))
    n_diag = int(round(pSelfCon * n))
    diag_on, diag_off = _pairs(X & eye), _pairs(~X & eye)
    if diag_on.shape[1] > n_diag:
        chosen.append(pick(diag_on, n_diag))
    else:
        chosen.append(diag_on)
        chosen.append(pick(diag_off, n_diag - diag_on.shape[1]))
    sel = np.concatenate(chosen, axis=1).astype(int)
    return np.array([cells_arr[sel[0]], cells_arr[sel[1]]]), (Nsyn + np.arange(sel.shape[1])).astype(int)


# ---- the generator (:15-170) -----------------------------------------------------------------------------------------
def _draw_membrane(P, cells, group, basics):
    """Multivariate-normal draw in the transformed space, inverse transform, re-draw of every cell that violates a bound
    (:42-66, 306-333).  P: [9, N] array filled in place."""
    m = basics.membr
    g = GROUP_NAMES.index(group)
    cov, k, mean, std, lo, hi = (np.asarray(t.values) for t in (m.covariance, m.k, m.mean, m.std, m.min, m.max))
    tau_lo, tau_hi = m.tau_m_min.values[g], m.tau_m_max.values[g]
    iC, igL, itw, iVr, iVT = (MEMB_PARAM_NAMES.index(x) for x in ("C", "g_L", "tau_w", "V_r", "V_T"))
    redo = cells
    for _ in range(1000):
        X = np.random.multivariate_normal(np.zeros(len(MEMB_PARAM_NAMES)), cov[g], len(redo)).transpose()
        for r in range(len(MEMB_PARAM_NAMES)):
            P[r, redo] = inv_transform(X[r], k[r, g], mean[r, g], std[r, g], lo[r, g])
        P[iC, redo] = P[iC, redo] * P[igL, redo]                       # the 'C' row is drawn as tau_m
        cur = P[:, cells]
        tau_m = cur[iC] / cur[igL]
        bad = (cur < lo[:, g][:, None]) | (cur > hi[:, g][:, None])
        bad = bad.any(axis=0) | (cur[iVr] >= cur[iVT]) | (tau_m < tau_lo) | (tau_m > tau_hi) | (cur[itw] <= tau_m) | np.isnan(cur).any(axis=0)
        redo = cells[bad]
        if len(redo) == 0:
            return
    raise RuntimeError("membrane parameters of group %s: 1000 re-draws exceeded (the reference's uniform fallback at "
                       "codes_revision/NetworkSetup.py:68-80 passes the wrong arguments and cannot run)" % group)


def _resort(group_distr, stripe, P):
    """IN_L / IN_L_d by onset latency against the LIF's, IN_CL / IN_CL_AC by the median onset/steady rate ratio (:215-304)."""
    gi = GROUP_NAMES.index
    memb = lambda c: MembraneTuple(*P[:, c])
    for layer in ("L23", "L5"):
        a, b = gi("IN_L_" + layer), gi("IN_L_d_" + layer)
        pool = list(group_distr[stripe][a]) + list(group_distr[stripe][b])
        late = [latency_AdEx(memb(c), 500) - latency_LIF(memb(c), 500) > 0 for c in pool]
        group_distr[stripe][a] = [c for c, l in zip(pool, late) if not l]
        group_distr[stripe][b] = [c for c, l in zip(pool, late) if l]
    I_range = np.arange(25, 301, 25)
    for layer in ("L23", "L5"):
        a, b = gi("IN_CL_" + layer), gi("IN_CL_AC_" + layer)
        pool = list(group_distr[stripe][a]) + list(group_distr[stripe][b])
        adapting = []
        for c in pool:
            mc = memb(c)
            f_trans = np.array([get_transFR(mc, I, 0, mc.E_L) for I in I_range], dtype=np.float64)     # None -> nan
            f_stat = np.array([get_statFR(mc, I) for I in I_range], dtype=np.float64)
            ok = np.where(f_stat > 0)[0]
            adapting.append(bool(np.median(f_trans[ok] / f_stat[ok]) > 1.5834))
        group_distr[stripe][a] = [c for c, ad in zip(pool, adapting) if not ad]
        group_distr[stripe][b] = [c for c, ad in zip(pool, adapting) if ad]
    return group_distr


def _stsp_counts(n, shares):
    cnt = np.round(n * shares, 0).astype(int)
    frac = n * shares - np.floor(n * shares)
    while sum(cnt) != n:
        if sum(cnt) > n:
            cnt[np.argmin(frac)] -= 1
        else:
            cnt[np.argmax(frac)] += 1
    return cnt


def _synapse_block(basics, gmax_fac, delay_fac, group_target, group_source, n):
    """Per-synapse parameters of one projection, one block per channel of the source's kind (:173-213).
    -> list of (channel index, gmax[n], pfail, U[n], tau_rec[n], tau_fac[n], delay[n]); the delays are shared by the blocks."""
    t, s = GROUP_NAMES.index(group_target), GROUP_NAMES.index(group_source)
    syn = basics.syn
    kind = str(syn.kinds.values[t, s])
    G = syn.spiking.params.gmax
    gmax, gsig = float(G["mean"].values[t, s]) * gmax_fac, float(G["sigma"].values[t, s]) * gmax_fac
    gmin, gmaxx = float(G["min"].values[t, s]), float(G["max"].values[t, s])
    D = syn.delay
    delay, dsig = float(D.delay_mean.values[t, s]) + delay_fac, float(D.delay_sigma.values[t, s]) + delay_fac
    dmin, dmax = float(D.delay_min.values[t, s]), float(D.delay_max.values[t, s])
    sets = syn.STSP.sets
    shares_all = np.asarray(sets.loc[dict(set=str(syn.STSP.distr.values[t, s]))].values)
    present = np.isfinite(shares_all)
    kinds, shares = np.asarray(sets.coords["kind"])[present], shares_all[present]
    K = syn.STSP.kinds
    delays = random_param(n, delay, dsig, dmin * delay, dmax * delay, "normal")
    blocks = []
    for channel in syn.channels.kinds_to_names[kind]:
        U, rec, fac = np.zeros(n), np.zeros(n), np.zeros(n)
        lim = np.concatenate([[0], np.cumsum(_stsp_counts(n, shares))]).astype(int)
        for q, kd in enumerate(kinds):
            row = {p: float(K.loc[kd, p].values) for p in K.coords["param"]}
            m = lim[q + 1] - lim[q]
            U[lim[q]:lim[q + 1]] = random_param(m, row["U_mean"], row["U_std"], 0, 1, "normal")
            rec[lim[q]:lim[q + 1]] = random_param(m, row["tau_rec_mean"], row["tau_rec_std"], 0, 1500, "normal")
            fac[lim[q]:lim[q + 1]] = random_param(m, row["tau_fac_mean"], row["tau_fac_std"], 0, 1500, "normal")
        mu = np.log((gmax ** 2) / np.sqrt(gsig ** 2 + gmax ** 2))
        sd = np.sqrt(np.log(gsig ** 2 / gmax ** 2 + 1))
        g = random_param(n, mu, sd, gmin * gmax, gmaxx * gmax, "log_normal")
        blocks.append((syn.channels.names.index(channel), g, float(syn.spiking.params.pfail.values[t, s]), U, rec, fac, delays))
    return blocks


def _revision_network(Ncells_prompted, Nstripes, basics_scales, seed):
    np.random.seed(seed)
    basics = basics_setup(Ncells_prompted, Nstripes, basics_scales, disp=False)
    if Nstripes > 1:
        raise NotImplementedError("the revision's generator cannot draw inter-stripe connections (codes_revision/NetworkSetup.py:144 "
                                  "passes two integers to SetCon_standard); use network_setup(..., generator='codes')")
    N = basics.struct.Ncells_total
    P = np.zeros((len(MEMB_PARAM_NAMES), N))
    I_ref = np.zeros(N)
    group_distr = [[[] for _ in GROUP_NAMES] for _ in range(Nstripes)]
    pairs, chan, gmax, pfail, U, rec, fac, delay = ([] for _ in range(8))
    pCon, cluster = basics.struct.conn.pCon.values, basics.struct.conn.cluster.values
    nsyn = 0

    def add(group_target, group_source, TS, n):
        nonlocal nsyn
        for c, g, pf, u, r, f, d in _synapse_block(basics, 1, 1, group_target, group_source, n):
            pairs.append(TS); chan.append(np.full(n, c)); gmax.append(g); pfail.append(np.full(n, pf))
            U.append(u); rec.append(r); fac.append(f); delay.append(d)
            nsyn += n

    first = 0
    for stripe in range(Nstripes):
        for gi, group in enumerate(GROUP_NAMES):
            cells = np.arange(int(basics.struct.Ncells_per_group.values[gi])) + first
            first += len(cells)
            group_distr[stripe][gi].extend(cells)
            _draw_membrane(P, cells, group, basics)
            for c in cells:
                mc = MembraneTuple(*P[:, c])
                I_ref[c] = fmin(Define_I_ref, mc.g_L * (mc.V_T - mc.E_L) - mc.g_L * mc.delta_T + 100, disp=False, args=(mc,))[0]
        group_distr = _resort(group_distr, stripe, P)
        for gi, group in enumerate(GROUP_NAMES):                      # within a group
            cells = group_distr[stripe][gi]
            if len(cells) > 0:
                if cluster[gi, gi] == 1:
                    TS, idc = SetCon_CommonNeighbour(nsyn, cells, float(pCon[gi, gi]), 0.47)
                else:
                    TS, idc = SetCon_standard(nsyn, cells, cells, float(pCon[gi, gi]))
                if len(idc) > 0:
                    add(group, group, TS, len(idc))
        for ti, gt in enumerate(GROUP_NAMES):                         # between groups
            for si, gs in enumerate(GROUP_NAMES):
                tc, sc = group_distr[stripe][ti], group_distr[stripe][si]
                if si != ti and len(tc) * len(sc) > 0:
                    TS, idc = SetCon_standard(nsyn, tc, sc, float(pCon[ti, si]))
                    if TS.shape[1] > 0:
                        add(gt, gs, TS, len(idc))
    cat = lambda parts, dt=np.float64: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dtype=dt)
    syn_pairs = np.concatenate(pairs, axis=1).astype(np.int64) if pairs else np.zeros((2, 0), dtype=np.int64)
    syn = {"channel": Labelled(cat(chan)[None, :], ["channel"], "syn_index"),
           "spiking": Labelled(np.stack([cat(gmax), cat(pfail)]), ["gmax", "pfail"], "syn_index"),
           "STSP_params": Labelled(np.stack([cat(U), cat(rec), cat(fac)]), ["U", "tau_rec", "tau_fac"], "syn_index"),
           "delay": Labelled(cat(delay)[None, :], ["delay"], "syn_index")}
    gd = [[list(map(int, g)) for g in st] for st in group_distr]
    return Network(Labelled(P, MEMB_PARAM_NAMES), Table(I_ref, "cell_index", [np.arange(N)]), syn, syn_pairs, gd, seed, basics)


def network_setup(Ncells_prompted, Nstripes, basics_scales=None, seed=None, generator="revision", workers=None):
    """codes_revision/NetworkSetup.py:15-170.  generator='revision' (default): the revision's own draws, one stripe;
    'codes': the codes/ generator (any number of stripes; `basics_scales` are not applied by it)."""
    import contextlib
    import io
    import warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        if generator == "revision":
            return _revision_network(Ncells_prompted, Nstripes, basics_scales, seed)
        if generator != "codes":
            raise ValueError("generator must be 'revision' or 'codes'")
        if basics_scales:
            raise NotImplementedError("basics_scales select tables of the revision's BasicsSetup; the codes/ generator takes its own `scales`")
        from ..codes.NetworkSetup import sparse_network
        np.random.seed(seed)
        scales = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))
        NeuPar, V0, STypPar, SynPar, (t, s), gd = sparse_network(Ncells_prompted, Nstripes, scales, True, workers=workers)
    return Network.from_codes(NeuPar, V0, STypPar, SynPar, t, s, gd, seed)
