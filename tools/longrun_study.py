"""Long-run statistics of the reference's default column over several NETWORK REALISATIONS, against the numbers the
paper publishes (content.tex:239, 318, 334-335, 344).

For every numpy seed s: `np.random.seed(s); NetworkSetup(1000, 1, scales, True)` (our bit-identical restatement of
codes/NetworkSetup.py:9), then one run analysed the way the reference's scripts do:
  * ISI mean / CV / spiking fractions: 61 s (codes/Main_ISI.py:29), neurons with >= 21 spikes in [1 s, 61 s)
    (codes/CortexNetwork.py:2748-2765 — the ISIs themselves are taken over ALL spikes of those neurons, as the reference does);
  * zero-lag Pearson correlation of 2 ms-binned trains over pairs of those neurons (codes/AuxiliarEquations.py:537-656);
  * LFP = sum of I_tot over all neurons, [1 s, 31 s), periodogram at fs = 1/dt, log-log slope below / above 60 Hz
    (codes/CortexNetwork.py:3986-4017, content.tex:334-335);
  * sub-threshold V standard deviation of neurons with >= 4 spikes, [1 s, 13 s), spikes excised
    (codes/Main_V.py:28, 392-397; codes/CortexNetwork.py:3480-3512; codes/AuxiliarEquations.py:263-306).

    python tools/longrun_study.py --engine gpu|oracle --seeds 0,1,2,3,4 --method rk4 --out profiles/r2_longrun.json
The oracle engine is test infrastructure (CPU); `--engine gpu` is the product path.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCALES = ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14)))
CACHE = os.path.join(ROOT, "_cache")


def build_network(seed, n1=1000, workers=8):
    """-> dict like the golden fixtures (NeuPar, V0, STypPar, SynPar, source_arr, target_arr, groups)."""
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, "net_n%d_s%d.npz" % (n1, seed))
    if os.path.exists(path):
        d = np.load(path, allow_pickle=True)
        return {k: d[k] for k in d.files}
    from rescience_hass_hertag_durstewitz_2016_b200.codes.NetworkSetup import sparse_network
    import warnings
    warnings.filterwarnings("ignore")
    np.random.seed(seed)
    NeuPar, V0, STypPar, SynPar, pairs, group_distr = sparse_network(n1, 1, SCALES, True, workers=workers)
    target_arr, source_arr = pairs
    flat, offs = [], [0]
    for g in range(len(group_distr)):
        flat.extend(int(x) for x in group_distr[g][0])
        offs.append(len(flat))
    out = dict(NeuPar=NeuPar, V0=V0, STypPar=STypPar, SynPar=SynPar, source_arr=np.asarray(source_arr, dtype=np.int32),
               target_arr=np.asarray(target_arr, dtype=np.int32), group_flat=np.asarray(flat, dtype=np.int32),
               group_offs=np.asarray(offs, dtype=np.int64), nstripes=1)
    np.savez_compressed(path, **out)
    return out


def groups_of(g):
    offs, flat = g["group_offs"], g["group_flat"]
    return [flat[offs[k]:offs[k + 1]] for k in range(len(offs) - 1)]


def idc_of(g, i_exc=250.0, i_inh=200.0):
    idc = np.zeros(g["NeuPar"].shape[1])
    for gi, ids in enumerate(groups_of(g)):
        idc[ids] = i_exc if gi in (0, 7) else i_inh
    return idc


# ---- the reference's analyses, restated on plain arrays ---------------------------------------------------------------
def isi_stats(i, t_ms, n, start, stop, min_sp=21, pc=None, include_transient=True):
    """codes/CortexNetwork.py:2748-2765 (ISI_analysis) + the spiking fractions of content.tex:239 (> 10 spikes per 30 s)."""
    sel = (t_ms >= start) & (t_ms < stop)
    counts = np.bincount(i[sel], minlength=n)
    chosen = np.nonzero(counts >= min_sp)[0]
    order = np.argsort(i, kind="stable")
    si, st = i[order], t_ms[order]
    b = np.searchsorted(si, np.arange(n + 1))
    means, cvs = [], []
    for q in chosen:
        tr = st[b[q]:b[q + 1]]
        if not include_transient:
            tr = tr[(tr >= start) & (tr < stop)]
        isi = np.diff(tr)
        means.append(isi.mean())
        cvs.append(isi.std() / isi.mean())
    dur_s = (stop - start) / 1000.0
    thr = 10.0 * dur_s / 30.0
    out = dict(n_trains=int(len(chosen)), mean_isi=float(np.mean(means)), sd_isi=float(np.std(means)),
               cv=float(np.mean(cvs)), sd_cv=float(np.std(cvs)), frac_spiking=float((counts > thr).mean()),
               rate_hz=float(sel.sum() / n / dur_s))
    if pc is not None:
        out["frac_spiking_pc"] = float((counts[pc] > thr).mean())
    return out, chosen


def zero_lag_pearson(i, t_ms, chosen, start, stop, t_bin=2.0, max_neurons=400, rng=None):
    """Zero-lag Pearson coefficient of binary 2 ms trains over all unordered pairs of `chosen`
    (codes/AuxiliarEquations.py:537-656: bins = ((t - start) / t_bin).astype(int), binary)."""
    nb = int((stop - start) / t_bin)
    if len(chosen) > max_neurons:
        chosen = (rng or np.random.default_rng(0)).choice(chosen, max_neurons, replace=False)
    M = np.zeros((len(chosen), nb), dtype=np.float32)
    pos = {int(q): k for k, q in enumerate(chosen)}
    sel = (t_ms >= start) & (t_ms < stop)
    for q, t in zip(i[sel], t_ms[sel]):
        k = pos.get(int(q))
        if k is not None:
            M[k, min(int((t - start) / t_bin), nb - 1)] = 1.0
    C = np.corrcoef(M.astype(np.float64))
    iu = np.triu_indices(len(chosen), 1)
    v = C[iu]
    v = v[np.isfinite(v)]
    return dict(cc0_mean=float(v.mean()), cc0_sd=float(v.std()), cc0_pairs=int(len(v)))


def lfp_slopes(lfp, dt_ms, f_split=60.0, f_lo=np.exp(2.0), f_hi=1000.0):
    """Periodogram of the summed I_tot (codes/CortexNetwork.py:3995-3998) and least-squares slopes of log P over log f."""
    from scipy.signal import periodogram
    lfp = np.asarray(lfp, dtype=np.float64)
    bad = ~np.isfinite(lfp)
    if bad.any():           # a neuron whose V ran away while refractory poisons the sum: bridge those samples
        good = np.nonzero(~bad)[0]
        lfp = np.interp(np.arange(len(lfp)), good, lfp[good])
    f, p = periodogram(lfp, 1000.0 / dt_ms)
    keep = (f > 0) & (p > 0)
    lf, lp = np.log(f[keep]), np.log(p[keep])

    def slope(a, b):
        m = (lf >= np.log(a)) & (lf < np.log(b))
        # equal weight per log-frequency interval (the figure is read on a log axis): average within 40 log bins first
        edges = np.linspace(np.log(a), np.log(b), 41)
        k = np.digitize(lf[m], edges) - 1
        xs = np.array([lf[m][k == j].mean() for j in range(40) if np.any(k == j)])
        ys = np.array([lp[m][k == j].mean() for j in range(40) if np.any(k == j)])
        return float(np.polyfit(xs, ys, 1)[0])
    return dict(lfp_slope_low=slope(f_lo, f_split), lfp_slope_high=slope(f_split, f_hi), lfp_nonfinite_samples=int(bad.sum()))


def extract_spikes(arr, V_r, V_T, ampl=1.0):
    """Spike excision of codes/AuxiliarEquations.py:263-306 restated: around every jump of |V| by >= 1 mV in one step
    (the reset), drop the samples after it that are still below the trace's resting minimum and the samples before it
    that are above V_T."""
    arr = np.asarray(arr, dtype=np.float64)[1:]
    d = np.diff(np.abs(arr))
    sppos = np.nonzero(d >= ampl)[0][::-1]
    # get_min (codes/AuxiliarEquations.py:248-260): local minima of the trace relative to V_r
    from scipy.signal import argrelextrema
    mins = arr[argrelextrema(arr, np.less_equal)]
    if len(mins) == 0:
        arrmin = V_r
    elif mins.min() > V_r:
        arrmin = mins[mins > V_r].min()
    else:
        arrmin = mins.max()
    drop = np.zeros(len(arr), dtype=bool)
    for i in sppos:
        j, last = i + 1, -1000.0
        while arr[j] < arrmin:
            drop[j] = True
            if arr[j] - last < 0 or j == len(arr) - 1:
                break
            last = arr[j]
            j += 1
        j = i
        while arr[j] > V_T:
            drop[j] = True
            if arr[j] - last > 0 or j == 0:
                break
            last = arr[j]
            j -= 1
    return arr[~drop]


def v_std(V, ids, NeuPar, pc):
    """V: [n_samples, len(ids)] in mV -> mean +- sd over neurons of the spike-excised standard deviation."""
    sd = np.array([np.std(extract_spikes(V[:, k], NeuPar[8, q], NeuPar[9, q])) for k, q in enumerate(ids)])
    is_pc = np.isin(ids, pc)
    return dict(v_sd_mean=float(sd.mean()), v_sd_sd=float(sd.std()), v_sd_pc_mean=float(sd[is_pc].mean()) if is_pc.any() else None,
                v_sd_pc_sd=float(sd[is_pc].std()) if is_pc.any() else None, v_neurons=int(len(ids)))


# ---- one realisation ----------------------------------------------------------------------------------------------
def run_one(seed, engine="gpu", method="rk4", dt=0.05, bio_ms=61000.0, lfp_ms=31000.0, v_ms=13000.0, engine_seed=None,
            engine_kw=None, with_v=True, save_raw=True):
    if engine == "gpu":
        from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine as E
    else:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        from oracle_binding import OracleEngine as E
    from rescience_hass_hertag_durstewitz_2016_b200.engine import load_codes_network
    g = build_network(seed)
    lfp_ms, v_ms = min(lfp_ms, bio_ms), min(v_ms, bio_ms)
    n = g["NeuPar"].shape[1]
    groups = groups_of(g)
    pc = np.concatenate([groups[0], groups[7]])
    e = E(n, dt, method={"gsl_rk2": "rk23"}.get(method, method), seed=seed if engine_seed is None else engine_seed, **(engine_kw or {}))
    load_codes_network(e, g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], g["source_arr"], g["target_arr"], idc_of(g))
    t0 = time.time()
    out = dict(net_seed=int(seed), engine=engine, method=method, dt=dt)
    lfp_mon = e.add_state_monitor("I_tot", None, "sum", capacity_steps=int(lfp_ms / dt))
    steps_done = 0
    if with_v:
        # pass 1 to v_ms with V of all neurons recorded (codes/Main_V.py records every neuron)
        v_mon = e.add_state_monitor("V", None, "none", capacity_steps=int(v_ms / dt))
        e.run(int(v_ms / dt))
        steps_done = int(v_ms / dt)
        V = e.read_monitor(v_mon)
        e.set_monitor_active(v_mon, False)
        e.clear_monitor(v_mon)
        i, s = e.spikes()
        counts = np.bincount(i, minlength=n)                       # spikemonitor.count: whole run (:3499)
        ids = np.nonzero(counts >= 4)[0]
        first = int(1000.0 / dt)
        out.update(v_std(V[first:, ids], ids, g["NeuPar"], pc))
        del V
    e.run(int(lfp_ms / dt) - steps_done)
    lfp = e.read_monitor(lfp_mon)[:, 0]
    e.set_monitor_active(lfp_mon, False)
    if save_raw:
        np.save(os.path.join(CACHE, "lfp_%s_%s_s%d.npy" % (engine, method, seed)), lfp.astype(np.float32))
    out.update(lfp_slopes(lfp[int(1000.0 / dt):], dt))
    e.run(int(bio_ms / dt) - int(lfp_ms / dt))
    i, s = e.spikes()
    t_ms = s * dt
    c = e.counters()
    e.close()
    if save_raw:
        np.savez_compressed(os.path.join(CACHE, "spikes_%s_%s_s%d.npz" % (engine, method, seed)), i=i, s=s)
    st, chosen = isi_stats(i, t_ms, n, 1000.0, bio_ms, 21, pc)
    out.update(st)
    st30, _ = isi_stats(i, t_ms, n, 1000.0, 31000.0, 11, pc)
    out.update(frac_spiking_30s=st30["frac_spiking"], frac_spiking_pc_30s=st30["frac_spiking_pc"])
    out.update(zero_lag_pearson(i, t_ms, chosen, 1000.0, bio_ms))
    out.update(nonfinite=int(c["nonfinite"]), spikes=int(c["spikes"]), syn_events=int(c["syn_events"]), wall_s=round(time.time() - t0, 2))
    return out


def summarise(rows):
    keys = ["mean_isi", "cv", "frac_spiking", "frac_spiking_pc", "rate_hz", "cc0_mean", "cc0_sd", "lfp_slope_low", "lfp_slope_high",
            "v_sd_mean", "v_sd_pc_mean"]
    return {k: dict(mean=float(np.mean([r[k] for r in rows if r.get(k) is not None])),
                    sd=float(np.std([r[k] for r in rows if r.get(k) is not None])),
                    min=float(np.min([r[k] for r in rows if r.get(k) is not None])),
                    max=float(np.max([r[k] for r in rows if r.get(k) is not None]))) for k in keys if any(r.get(k) is not None for r in rows)}


def _worker(a):
    return run_one(*a[0], **a[1])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="gpu")
    ap.add_argument("--seeds", default="0,1,2,3,4")
    ap.add_argument("--method", default="rk4")
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--bio-ms", type=float, default=61000.0)
    ap.add_argument("--no-v", action="store_true")
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    seeds = [int(x) for x in a.seeds.split(",")]
    for s in seeds:
        build_network(s)
    jobs = [((s,), dict(engine=a.engine, method=a.method, dt=a.dt, bio_ms=a.bio_ms, with_v=not a.no_v)) for s in seeds]
    if a.procs > 1:
        import multiprocessing as mp
        os.environ["OMP_NUM_THREADS"] = "1"
        with mp.get_context("spawn").Pool(a.procs) as pool:
            rows = pool.map(_worker, jobs)
    else:
        rows = []
        for j in jobs:
            rows.append(_worker(j))
            print(json.dumps(rows[-1]), flush=True)
    res = dict(paper=dict(mean_isi="523 +- 655 (rk4) / 511 +- 665 (gsl_rk2) / 516 +- 680 (original)", cv="1.13 +- 0.49 / 0.97 +- 0.45 / 1.10 +- 0.45",
                          frac_spiking="20.66 +- 2.55 % (16.95-27.62)", frac_spiking_pc="12.28 +- 2.81 % (8.12-20.12)",
                          cc0="9.7e-4 +- 6.7e-3", v_sd="2.82 +- 1.43 mV (PC 2.42 +- 0.94)", lfp="slope ~ -1 below 60 Hz, between -2 and -3 above"),
               runs=rows, summary=summarise(rows))
    print(json.dumps(res["summary"], indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
