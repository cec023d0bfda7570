"""`NetworkSetup(N1, Nstripes, scales, clustering)` — network-instance construction with the reference's signature and
return values (codes/NetworkSetup.py:9, 323), bit-identical for the same numpy seed: every `np.random` call happens in
the reference's order with the reference's arguments (tests/test_network_setup.py checks the golden fixtures).

Differences in form, not in result:
  * synapses are kept sparse while building (pair lists per block instead of the dense N x N x 2 `SPMtx`); the dense
    matrix is materialised at the end only for networks where it is affordable (`dense_limit` neurons), otherwise the
    fifth return value is the pair `(target_arr, source_arr)` that our CortexNetwork accepts as well;
  * the per-neuron work that consumes no random numbers (I_ref by `fmin` over the closed-form f-I curve, latency
    integrals, adaptation ratio: 90 ms per neuron, :112-149) can run in worker processes (`workers`);
  * `sparse_network(...)` returns (target_arr, source_arr) directly = `sourcetarget(SPMtx)` of the reference.
"""
import os

import numpy as np
from scipy.integrate import quad
from scipy.optimize import fmin, fsolve

from .ParamSetup import ParamSetup

_ROWS = [0, 1, 2, 3, 4, 5, 7, 8, 9]          # NeuPar rows filled from the 9 transformed normals (a = row 6 stays 0)
DENSE_LIMIT = 6000


# ------------------------------------------------------------------------------------------------------------------
# closed-form simplified-AdEx firing rate (codes/NetworkSetup.py:444-617) — deterministic, no random numbers
# ------------------------------------------------------------------------------------------------------------------
def _isi_integrand(V, I, par, i, j, k):
    Cm, gL, EL, sf, Vup, tcw, a, b, Vr, Vth = par[:10]
    e = np.exp((V - Vth) / sf)
    F = (1 / Cm) * (I - (i * V + j + k * e) + gL * sf * e - gL * (V - EL))
    return 1 / F


def IntPoints_simpAdEx(values, I, w0, w_ref, i):
    """Intersections of the trajectory w = w0 with the (scaled) V-nullcline (codes/NetworkSetup.py:559-606)."""
    Cm, gL, EL, sf, Vup, tcw, a, b, Vr, Vth = values[:10]
    Vfix = [None, None]
    if w0 > w_ref:
        f = Cm / (gL * tcw)

        def G(V):
            return (1 + i * f) * (I - gL * (V - EL) + gL * sf * np.exp((V - Vth) / sf)) - w0

        base = (I - (w0 / (1 + i * f))) / gL
        lb, ub, numcor = EL + base - 0.1, EL + sf + base, 0
        while np.sign(G(ub)) / np.sign(G(lb)) == 1:
            lb, ub = EL + base - numcor, EL + sf + base
            numcor += 1
            if numcor > 1000:
                print("Error in numcor")
        Vfix[0] = fsolve(G, (lb + ub) / 2)
        lb, ub, numcor = EL + sf + base, Vup, 0
        while np.sign(G(ub)) / np.sign(G(lb)) == 1:
            lb, ub = EL + sf + base - numcor, Vup
            numcor += 1
            if numcor > 1000:
                print("Error")
        Vfix[1] = fsolve(G, (lb + ub) / 2)
    elif w0 == w_ref:
        Vfix = Vth
    else:
        Vfix = []
    return Vfix


def FRsimpAdEx(values, I, w0=[], V0=[], bmin=[]):
    """Steady (w0 = []) or onset (w0 = 0) firing rate in Hz for constant current I; None where the reference's function
    falls off its end (no spiking), exactly as codes/NetworkSetup.py:444-551 behaves."""
    Cm, gL, EL, sf, Vup, tcw, a, b, Vr, Vth = values[:10]
    tau = Cm / gL
    f = tau / tcw
    X_Vth = f * (I + gL * sf - gL * (Vth - EL))
    w_end = -gL * (Vth - EL) + gL * sf + I - X_Vth
    if not np.size(w0):
        w_r = w_end + b if b != 0 else 0
    else:
        w_r = w0
    if not np.size(bmin) or bmin < 0:
        bmin = 0
    V_r = Vr if not np.size(V0) else V0
    if X_Vth <= 0 or f >= 1 or b < bmin or Cm <= 0 or gL <= 0 or tcw <= 0 or sf <= 0:
        return None                                               # the reference assigns fr = 0 and then returns nothing
    e_r = np.exp((V_r - Vth) / sf)
    X_Vr = f * (I + gL * sf * e_r - gL * (V_r - EL))
    wV_Vr = -gL * (V_r - EL) + gL * sf * e_r + I
    w1, w2 = wV_Vr - X_Vr, wV_Vr + X_Vr
    m = f * gL - gL
    slow = (m, (1 - f) * I - m * EL, (1 - f) * gL * sf)            # (i, j, k) while w follows the nullcline band

    def T(lo, hi, ijk):
        return quad(lambda x: _isi_integrand(x, I, values, *ijk), lo, hi)[0]

    t1 = t2 = 0
    if V_r >= Vth:
        if w_r >= wV_Vr:
            w_ref = -gL * (Vth - EL) + gL * sf + I + X_Vth
            Vfix = IntPoints_simpAdEx(values, I, w_r, w_ref, 1)
            if not Vfix or len(Vfix) == 1:
                t1 = T(V_r, Vth, (0, w_r, 0))
            else:
                Vlb = np.min(Vfix)
                t1 = T(V_r, Vlb, (0, w_r, 0))
                t2 = T(Vlb, Vth, slow)
            w_stop, Vlb = w_end, Vth
        else:
            Vlb, w_stop = V_r, w_r
    else:
        if w1 < w_r < w2:
            t2 = T(V_r, Vth, slow)
            w_stop = w_end
        else:
            if w_r <= w1:
                ns, w_ref = -1, -gL * (Vth - EL) + gL * sf + I - X_Vth
            else:
                ns, w_ref = 1, -gL * (Vth - EL) + gL * sf + I + X_Vth
            Vfix = IntPoints_simpAdEx(values, I, w_r, w_ref, ns)
            if not Vfix or len(Vfix) == 1:
                t1 = T(V_r, Vth, (0, w_r, 0))
                w_stop = w_r
            else:
                Vlb = np.min(Vfix)
                t1 = T(V_r, Vlb, (0, w_r, 0))
                t2 = T(Vlb, Vth, slow)
                w_stop = w_end
        Vlb = Vth
    t3 = 0 if Vlb >= Vup else T(Vlb, Vup, (0, w_stop, 0))
    return 1000 / np.asarray(t1 + t2 + t3)


def Define_I_ref(I0, Par):
    re = FRsimpAdEx(Par, I0, 0, [], [])
    if re is None:
        re = 0
    return (re - 200) ** 2


def _per_neuron(col):
    """I_ref, latency integrals and adaptation ratio of one neuron (codes/NetworkSetup.py:112-147)."""
    import warnings
    warnings.filterwarnings("ignore")
    C, gL, EL, sf, Vup, tcw, _, b, Vr, Vth = col[:10]
    I = 500
    Irheo = gL * (Vth - EL) - gL * sf
    Iref = fmin(Define_I_ref, Irheo + 100, disp=False, args=(col,))
    t_lat = quad(lambda V: C / (I - gL * (V - EL) + gL * sf * np.exp((V - Vth) / sf)), EL, Vth / 2)[0]
    t_lif = C * np.log(I / (I + gL * (EL - Vth / 2))) / gL
    cur = list(range(25, 301, 25))
    f1 = np.zeros(len(cur))
    f = np.zeros(len(cur))
    for q, Iq in enumerate(cur):
        f1[q] = FRsimpAdEx(col, Iq, 0, col[2], [])
        f[q] = FRsimpAdEx(col, Iq, [], [], [])
    pos = np.where(f > 0)[0]
    med = np.median(f1[pos] / f[pos])
    return float(np.asarray(Iref).reshape(-1)[0]), t_lat, t_lif, 0 if np.isnan(med) else med


# ------------------------------------------------------------------------------------------------------------------
# random draws (same calls, same order as the reference)
# ------------------------------------------------------------------------------------------------------------------
def inv_transform_distribution2(X_trans, k, mean_X, std_X, min_X):
    """Inverse Box-Cox-like transform of a standard-normal sample (codes/NetworkSetup.py:424-441)."""
    z = mean_X + 0.8 * std_X * X_trans
    X = z ** (1 / k) if k > 0 else np.exp(z)
    return X + 1.1 * min_X if min_X < 0 else X


def rand_par(N, par_mean, par_std, par_min, par_max, distr_flag):
    """Truncated normal (0) / uniform (1) / log-normal (2) with out-of-range draws replaced by uniform ones
    (codes/NetworkSetup.py:621-643)."""
    if par_std == 0:
        if par_max == 0:
            return par_mean * np.ones(N)
        return par_min + (par_max - par_min) * np.random.random(size=N)
    if distr_flag == 1:
        return par_min + (par_max - par_min) * np.random.random(size=N)
    if distr_flag == 0:
        par = par_mean + par_std * np.random.normal(size=N)
    else:
        par = np.exp(np.random.normal(size=N) * par_std + par_mean)
    out = np.where((par < par_min) | (par > par_max))[0]
    par[out] = par_min + (par_max - par_min) * np.random.random(size=len(out))
    return par


def SetCon(Nsyn, Nin, Nout, pCon):
    """Random block: round(pCon*Nin*Nout) of the Nin x Nout (target x source) pairs, numbered Nsyn+1.. in the order of
    the shuffled flat indices q = source*Nin + target (codes/NetworkSetup.py:326-339)."""
    k1 = np.arange(Nin * Nout)
    n = int(round(pCon * Nin * Nout, 0))
    np.random.shuffle(k1)
    k = k1[:n]
    idc = Nsyn + np.arange(n) + 1
    X = np.zeros(Nin * Nout)
    X[k] = idc
    return X.reshape(Nout, Nin).T, idc


def find_neigh_recur(X):
    """Number of common neighbours of every pair in the symmetrised block (codes/NetworkSetup.py:793-824)."""
    n = X.shape[0]
    Xl = X.copy()
    Xl[np.arange(n), np.arange(n)] = 0
    S = ((Xl + Xl.T) > 0).astype(np.float64)
    neigh = S @ S                                  # pairs (a, b) sharing a neighbour i, counted over i (small integers: exact in fp64, and BLAS)
    neigh[np.arange(n), np.arange(n)] = -1
    return neigh


def p_min(N0, p, pCon, slope):
    N1 = np.round(min(max(p[:, 0]), N0 + 1 / slope / pCon))
    mid = np.where((p[:, 0] > N0) & (p[:, 0] <= N1))
    N_2 = p[mid, 0]
    pN_2 = p[mid, 1] / np.sum(p[:, 1])
    pN_3 = p[np.where(p[:, 0] > N1), 1] / np.sum(p[:, 1])
    return abs(np.sum(pN_2 * slope * (N_2 - N0)) + np.sum(pN_3) - 1)


def p_calc_recur(X, N_neigh, pCon, pSelfCon, slope):
    """Connection probability as a function of the number of common neighbours (codes/NetworkSetup.py:829-887)."""
    n = X.shape[0]
    Xl = X.copy()
    Xl[np.arange(n), np.arange(n)] = 0
    Xl = np.tril(Xl)
    sel_t, sel_s = np.nonzero(Xl > 0)                  # connected pairs below the diagonal
    N_sel = N_neigh[sel_t, sel_s]
    p1 = np.unique(N_neigh[N_neigh >= 0])
    p = np.zeros((len(p1), 6))
    p[:, 0] = p1
    for q in range(len(p1)):
        p[q, 1] = np.count_nonzero(N_neigh == p1[q])
        p[q, 2] = np.sum((N_sel == p1[q]).astype(int))
        p[q, 3] = p[q, 2] / p[q, 1]
    off = fmin(lambda N0: p_min(N0, p, pCon, slope), np.max(p[:, 0]), disp=False)
    N1 = np.floor(min(np.max(p[:, 0]), off + 1 / slope / pCon))
    N0 = max(np.min(p[:, 0]), np.ceil(off))
    ind = np.isin(p[:, 0], np.arange(N0, N1 + 1))
    p[ind, 4] = pCon * slope * (p[ind, 0] - off)
    hit = np.where(ind.astype(int))[0]
    if len(ind) and len(hit):
        p[hit[-1] + 1:, 4] = 1
    else:
        p[:, 4] = 1
    N_original = n ** 2 * pCon - n * pSelfCon
    p[:, 5] = p[:, 4] * N_original / (p[:, 1] @ p[:, 4])
    return p


def _pairset(mask):
    """Set of (target, source) tuples of the mask's True entries, inserted in row-major order.  Tuples of Python ints hash and
    compare like the tuples of numpy integers the reference builds, so the set iterates in the same order — at a fifth of the cost."""
    rows, cols = np.where(mask)
    return set(zip(rows.tolist(), cols.tolist()))


def SetCon_CommonNeighbour_Recur(Nsyn, Npop, pCon, pSelfCon):
    """Recurrent block with the common-neighbour rule and a prescribed fraction of reciprocal pairs and autapses
    (codes/NetworkSetup.py:646-787).  The selection walks Python sets of (target, source) tuples, whose iteration order
    is part of the reference's result, so the same sets are built by the same operations in the same order."""
    slope = 20 * 3.9991 / Npop
    X, _ = SetCon(Nsyn, Npop, Npop, pCon)
    N_neigh = find_neigh_recur(X)
    p0 = p_calc_recur(X, N_neigh, pCon, pSelfCon, slope)
    p = np.stack([p0[:, 0], p0[:, 5]], axis=1)

    eye = np.eye(Npop, dtype=bool)
    X_off = np.where(eye, 0, X)
    X_one = np.where(eye, 1, X)
    k = _pairset(X_off > 0)
    X_int = (X_off > 0).astype(int)
    X_recur = X_int * X_int.T
    k_recur_tril = _pairset(np.tril(X_recur) > 0)
    k_recur = _pairset(X_recur > 0)
    k_missing = _pairset((X_one + X_one.T == 0).astype(int) > 0)

    chosen = []
    for q in range(p.shape[0]):
        here = _pairset(N_neigh == p[q, 0])
        missing = k_missing.intersection(here)
        old = k.intersection(here)
        rec_old = k_recur_tril.intersection(here)
        rec_old2 = k_recur.intersection(here)
        uni_old = old.difference(rec_old2)
        N_rec = (np.floor(p[q, 1] * pSelfCon / 2 * len(here))).astype(int)
        N_uni = (np.ceil(p[q, 1] * len(here)) - 2 * N_rec).astype(int)

        if len(rec_old) >= N_rec:
            order = np.arange(len(rec_old))
            np.random.shuffle(order)
            for c1, c2 in np.asarray(list(rec_old))[np.asarray(order[:N_rec])]:
                chosen.extend([(c1, c2), (c2, c1)])
        else:
            for c1, c2 in list(rec_old):
                chosen.extend([(c1, c2), (c2, c1)])
            N_new = N_rec - len(rec_old)
            cand = list(missing)
            order = np.arange(len(cand))
            np.random.shuffle(order)
            taken = []
            got = pos = 0
            if len(order) > 0 and len(cand) > 0:
                while got < N_new:
                    c1, c2 = cand[order[pos]]
                    pos += 1
                    if c1 > c2:
                        chosen.extend([(c1, c2), (c2, c1)])
                        taken.extend([(c1, c2), (c2, c1)])
                        got += 1
                missing = missing.difference(set(taken))

        if len(uni_old) > N_uni:
            order = np.arange(len(uni_old))
            np.random.shuffle(order)
            chosen.extend(np.asarray(list(uni_old))[np.asarray(order[:N_uni])])
        else:
            chosen.extend(list(uni_old))
            N_new = N_uni - len(uni_old)
            cand = np.asarray(list(missing))
            order = np.arange(len(cand))
            np.random.shuffle(order)
            chosen.extend(cand[np.asarray(order[:N_new])])

    diag_old = list(map(tuple, zip(*np.where((X * np.identity(Npop)) > 0))))
    n_diag = int(round(pSelfCon * Npop))
    if len(diag_old) > round(pSelfCon * Npop):
        order = np.arange(len(diag_old))
        np.random.shuffle(order)
        chosen.extend(np.asarray(diag_old)[order[:n_diag]])
    else:
        chosen.extend(diag_old)
        free = set((q, q) for q in range(Npop)).difference(set(diag_old))
        order = np.arange(n_diag - len(diag_old))
        np.random.shuffle(order)
        chosen.extend(np.asarray(np.asarray(list(free))[order]))

    idc = Nsyn + np.arange(len(chosen)) + 1
    pairs = np.vstack(chosen)
    X_out = np.zeros(X.shape)
    X_out[pairs[:, 0], pairs[:, 1]] = idc
    return X_out, idc


def SetSyn(SynPar, Nsyn, idc, ST, gmax, delay, p_fail, STSP_types, STSP_prob, get_STSP_prob, get_STSP_types, STSP_setup,
           S_sig, S_max, S_min):
    """Parameters of the synapses idc of one (target group, source group, type) block (codes/NetworkSetup.py:342-401):
    rows 0 type, 1 U, 2 tau_rec, 3 tau_fac, 4 g_max, 5 delay, 6 p_fail."""
    cols = idc - 1
    if not np.size(SynPar):
        SynPar = np.zeros((7, max(cols) + 1))
    if np.max(idc) > SynPar.shape[1]:
        grown = np.zeros((7, max(cols) + 1))
        grown[:, :SynPar.shape[1]] = SynPar
        SynPar = grown
    n = len(cols)
    SynPar[0, cols] = ST
    if gmax == 0 or S_sig[0] == 0:
        SynPar[4, cols] = rand_par(n, gmax, S_sig[0], S_min[0] * gmax, S_max[0] * gmax, 0)
    else:
        mu = np.log((gmax ** 2) / np.sqrt(S_sig[0] ** 2 + gmax ** 2))
        sd = np.sqrt(np.log(S_sig[0] ** 2 / gmax ** 2 + 1))
        SynPar[4, cols] = rand_par(n, mu, sd, S_min[0] * gmax, S_max[0] * gmax, 2)
    SynPar[5, cols] = rand_par(n, delay, S_sig[1], S_min[1] * delay, S_max[1] * delay, 0)
    SynPar[6, cols] = p_fail

    probs = np.asarray(get_STSP_prob[STSP_prob])
    if n:
        share = (np.round(probs * n)).astype(int)
        if sum(share) > n:
            share[np.argmin(probs * n - np.floor(probs * n))] -= 1
        elif sum(share) < n:
            share[np.argmax(share - np.floor(share))] += 1
    classes = get_STSP_types[STSP_types]
    start = Nsyn
    for q in range(len(probs)):
        act = (start + np.arange(share[q])).astype(int)
        start += share[q]
        tab = STSP_setup[classes[q]]
        SynPar[1, act] = rand_par(len(act), tab[0, 0], tab[1, 0], 0, 1, 0)
        SynPar[2, act] = rand_par(len(act), tab[0, 1], tab[1, 1], 0, 1500, 0)
        SynPar[3, act] = rand_par(len(act), tab[0, 2], tab[1, 2], 0, 1500, 0)
    Nsyn += len(idc)
    return SynPar, Nsyn, Nsyn + np.arange(len(idc)) + 1


def sourcetarget(SPMtx):
    from .CortexNetwork import sourcetarget as st
    return st(SPMtx)


# ------------------------------------------------------------------------------------------------------------------
class _Blocks:
    """Sparse record of the `SPMtx[Idc1, Idc2, k] = X + k*len(idc)` assignments, in order."""

    def __init__(self):
        self.items = []

    def add(self, targets, sources, X, n_types, n_idc):
        self.items.append((np.asarray(targets, dtype=np.int64), np.asarray(sources, dtype=np.int64), X, n_types, int(n_idc)))

    def dense(self, n):
        S = np.zeros((n, n, 2))
        for tg, sr, X, n_types, n_idc in self.items:
            for k in range(n_types):
                S[np.ix_(tg, sr, [k])] = (X + k * n_idc)[:, :, None]
        return S

    def sparse(self, nsyn):
        """= sourcetarget(dense SPMtx) wherever no block is overwritten by a later one (single columns, and column
        counts other than 2, 3, 4, 6, 8 — SURVEY App. B.10)."""
        target_arr = np.zeros(nsyn, dtype=np.int64)
        source_arr = np.zeros(nsyn, dtype=np.int64)
        for tg, sr, X, n_types, n_idc in self.items:
            a, b = np.nonzero(X > 0)
            first = X[a, b].astype(np.int64) - 1
            for k in range(n_types):
                target_arr[first + k * n_idc] = tg[a]
                source_arr[first + k * n_idc] = sr[b]
        return target_arr, source_arr


def _build(N1, Nstripes, scales, clustering, workers, fast=False):
    (NTypes, CellPar, V0Par, k_trans, ParCov, N_sig, N_max, N_min, SynTypes, Syngmax, Syndelay, Synpfail, STSP_prob,
     STSP_types, pCon, cluster_flag, S_sig, S_max, S_min, p_fail, _E1, _E2, _E3, _I1, _I2, _I3, STSP_setup, get_SynType,
     get_STSP_prob, get_STSP_types, STypPar, ConParStripes) = ParamSetup(N1, Nstripes, scales)
    NTypes = np.array(NTypes, dtype=float)
    N = int(round(np.sum(NTypes), 0))
    G = len(NTypes)
    NeuPar = np.zeros((len(CellPar[:, 0]) + 2, Nstripes * N))
    V0 = np.zeros((2, Nstripes * N))
    blocks = _Blocks()
    SynPar, Nsyn, NN = [], 0, 0
    group_distr = [[[] for _ in range(Nstripes)] for _ in range(G)]
    t_lat = np.zeros(N * Nstripes)
    t_lat_LIF = np.zeros(N * Nstripes)
    adapt = np.zeros(N * Nstripes)
    pool = None
    if workers and workers > 1:
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(workers)

    def violations(iset, i):
        bad = [np.where((NeuPar[r, iset] < N_min[r + 1, i]) | (NeuPar[r, iset] > N_max[r + 1, i]))[0] for r in _ROWS]
        tau = NeuPar[0, iset] / NeuPar[1, iset]
        bad.append(np.where(NeuPar[8, iset] >= NeuPar[9, iset])[0])                      # V_r must stay below V_T
        bad.append(np.where((tau < N_min[11, i]) | (tau > N_max[11, i]))[0])
        bad.append(np.where(NeuPar[5, iset] <= tau)[0])                                   # tau_w must exceed tau_m
        bad.append(np.argwhere(np.isnan(NeuPar[:, iset]))[:, 1])
        return np.unique(np.concatenate(bad)).astype(int)

    def synapse_block(tg_ids, sr_ids, X, idc, i, j, gmax, delay, pf, sig):
        nonlocal SynPar, Nsyn
        types = get_SynType[SynTypes[i, j]]
        blocks.add(tg_ids, sr_ids, X, len(types), len(idc))   # the twin's index offset is len(idc), duplicates included
        back = None
        for k, ST in enumerate(types):
            cols = idc - 1
            SynPar, Nsyn, idc = SetSyn(SynPar, Nsyn, idc, ST, gmax, delay, pf, STSP_types[i, j], STSP_prob[i, j], get_STSP_prob,
                                       get_STSP_types, STSP_setup, sig, S_max[i, j, :], S_min[i, j, :])
            if k == 0:
                back = SynPar[5, cols]
            else:
                SynPar[5, cols] = back                     # the NMDA twin shares the AMPA synapse's delay (:230-234)

    for ii in range(Nstripes):
        for i in range(G):
            iset = (np.arange(int(round(NTypes[i], 0))) + NN).astype(int)
            group_distr[i][ii].extend(iset)
            out = iset
            tries = 0
            while out.size and tries < 999:
                z = np.random.multivariate_normal(np.zeros(len(ParCov[i][:, 0])), ParCov[i], len(out)).transpose()
                for j in range(9):
                    z[j, :] = inv_transform_distribution2(z[j, :], k_trans[j, i], CellPar[j + 1, i], N_sig[j + 1, i], N_min[_ROWS[j] + 1, i])
                for j in range(9):
                    NeuPar[_ROWS[j], out] = z[j, :]
                NeuPar[0, out] = NeuPar[0, out] * NeuPar[1, out]              # C from tau_m and g_L
                out = iset[violations(iset, i)]
                tries += 1
            tries = 0
            while out.size and tries < 1000:      # uniform fallback; the reference never refreshes `out` here (:81-104)
                if tries == 0:
                    print("WARNING: Number of trials for the full multivariate distribution has been exceeded. Use multivariate uniform distribution for the rest.\n")
                z = np.random.uniform(N_min[np.asarray(_ROWS) + 1, i], N_max[np.asarray(_ROWS) + 1, i], len(out)).transpose()
                for j in range(9):
                    NeuPar[_ROWS[j], out] = z[j, :]
                NeuPar[0, out] = NeuPar[0, out] * NeuPar[1, out]
                NeuPar[6, out] = 0
                tries += 1
            if out.size:
                print("ERROR: Number of trials for univariate uniform distribution has been exceeded.\n")

            cols = [NeuPar[:, n].copy() for n in iset]
            if fast:                            # all neurons of the group at once (fast_neuron.py: not bit-identical, see there)
                from .fast_neuron import per_neuron as _per_neuron_vec
                res = list(zip(*_per_neuron_vec(NeuPar[:, iset], latency=i in (1, 2, 8, 9), adaptation=i in (3, 4, 10, 11)))) if len(iset) else []
            else:
                res = pool.map(_per_neuron, cols, chunksize=4) if pool is not None else [_per_neuron(c) for c in cols]
            for n, (Iref, tl, tlif, ad) in zip(iset, res):
                NeuPar[10, n] = Iref
                NeuPar[11, n] = NeuPar[8, n]                                    # V_dep = V_r (:127)
                t_lat[n], t_lat_LIF[n], adapt[n] = tl, tlif, ad
            V0[1, iset] = 0
            V0[0, iset] = NeuPar[2, iset]
            NN = NN + NTypes[i]

        # re-typing: IN-L -> IN-L-d by latency, IN-CL -> IN-CL-AC by adaptation (:158-204)
        for a, b, crit in ((1, 2, "lat"), (8, 9, "lat"), (3, 4, "ad"), (10, 11, "ad")):
            ids = np.asarray(group_distr[a][ii] + group_distr[b][ii])
            if len(ids) > 0:
                sel = np.where((t_lat[ids] - t_lat_LIF[ids] > 0) if crit == "lat" else (adapt[ids] > 1.5834))[0]
                moved = ids[sel]
                kept = np.setdiff1d(ids, moved)
                NTypes[a], NTypes[b] = len(kept), len(moved)
                group_distr[a][ii], group_distr[b][ii] = list(kept), list(moved)

        for i in range(G):                                  # within a cell type
            ids = group_distr[i][ii]
            if len(ids):
                if cluster_flag[i, i] == 1 and clustering:
                    X, idc = SetCon_CommonNeighbour_Recur(Nsyn, len(ids), pCon[i, i], 0.47)
                else:
                    X, idc = SetCon(Nsyn, len(ids), len(ids), pCon[i, i])
                if len(idc) > 0:
                    synapse_block(ids, ids, X, idc, i, i, Syngmax[i, i], Syndelay[i, i], Synpfail[i, i], S_sig[i, i, :])
        for i in range(G):                                  # between cell types: target i, source j
            if len(group_distr[i][ii]):
                for j in np.setdiff1d(np.arange(G), [i]):
                    if len(group_distr[j][ii]):
                        X, idc = SetCon(Nsyn, len(group_distr[i][ii]), len(group_distr[j][ii]), pCon[i, j])
                        if np.size(idc):
                            synapse_block(group_distr[i][ii], group_distr[j][ii], X, idc, i, j, Syngmax[i, j], Syndelay[i, j],
                                          Synpfail[i, j], S_sig[i, j, :])
    if pool is not None:
        pool.close()
    print("REPORT: Synaptic connectivity and parameters generated\n")

    if Nstripes > 1:                                        # between stripes (:267-321)
        pConStripes, coefStripes, interStripes = ConParStripes
        for q, (ia, ja) in enumerate(pConStripes):
            for j in range(Nstripes):
                for off in interStripes[q]:
                    tgt = (j + off) % Nstripes
                    dist = abs(off)
                    fall = np.exp(-dist / coefStripes[q][0])
                    X, idc = SetCon(Nsyn, len(group_distr[ia][tgt]), len(group_distr[ja][j]), pCon[ia, ja] * fall)
                    sig = [S_sig[ia, ja, 0] * fall, S_sig[ia, ja, 1] + coefStripes[q][1] * dist]
                    synapse_block(group_distr[ia][tgt], group_distr[ja][j], X, idc, ia, ja, Syngmax[ia, ja] * fall,
                                  Syndelay[ia, ja] + coefStripes[q][1] * dist, 0.3, sig)
    return NeuPar, V0, STypPar, SynPar, blocks, group_distr, Nsyn


def NetworkSetup(N1, Nstripes, scales, clustering, workers=None, dense_limit=DENSE_LIMIT):
    """-> NeuPar, V0, STypPar, SynPar, SPMtx, group_distr   (codes/Main.py:585).  For more than `dense_limit` neurons
    SPMtx is the pair (target_arr, source_arr)."""
    workers = workers if workers is not None else int(os.environ.get("HHD_SETUP_WORKERS", "1"))
    NeuPar, V0, STypPar, SynPar, blocks, group_distr, nsyn = _build(N1, Nstripes, scales, clustering, workers)
    n = NeuPar.shape[1]
    SPMtx = blocks.dense(n) if n <= dense_limit else blocks.sparse(nsyn)
    return NeuPar, V0, STypPar, SynPar, SPMtx, group_distr


def sparse_network(N1, Nstripes, scales, clustering, workers=None, fast=False):
    """Same draws, sparse result: NeuPar, V0, STypPar, SynPar, (target_arr, source_arr), group_distr.
    fast=True: I_ref, latencies and adaptation ratios of a whole cell group as array expressions (fast_neuron.py, 60x faster,
    I_ref within 1e-4 pA of the reference's value instead of bit-identical) — the path for many-column networks."""
    workers = workers if workers is not None else int(os.environ.get("HHD_SETUP_WORKERS", "1"))
    NeuPar, V0, STypPar, SynPar, blocks, group_distr, nsyn = _build(N1, Nstripes, scales, clustering, 0 if fast else workers, fast)
    return NeuPar, V0, STypPar, SynPar, blocks.sparse(nsyn), group_distr
