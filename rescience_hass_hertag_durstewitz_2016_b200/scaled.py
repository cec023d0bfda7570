"""Scaled-up inputs for configs 3-5 of BASELINE.json: many columns / many independent instances in one engine.

`tile_columns` repeats a one-column template (the arrays `codes/NetworkSetup.NetworkSetup` returns for Nstripes=1)
K times with shifted neuron ids.  Every copy keeps the template's parameters and intra-column connectivity; what
differs between copies is the synapse index (hence every release-failure draw) and, optionally, the background
current.  With `inter_column=False` this is exactly BASELINE config 3 (independent instances batched on one GPU);
it is also the synthetic stand-in used by bench.py for the 100- and 1000-column workloads until the sparse
multi-column builder replaces it.
"""
import numpy as np


def usable_cpus():
    """Cores this process may run on (the affinity mask where the platform has one: a container often sees fewer than cpu_count)."""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def tile_columns(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, n_columns, jitter_idc=0.0, seed=0):
    """-> dict(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, n_per_column) for n_columns copies."""
    NeuPar = np.asarray(NeuPar, dtype=np.float64)
    n1 = NeuPar.shape[1]
    k = int(n_columns)
    offs = (np.arange(k, dtype=np.int64) * n1)
    src = (np.asarray(source_arr, dtype=np.int64)[None, :] + offs[:, None]).reshape(-1).astype(np.int32)
    tgt = (np.asarray(target_arr, dtype=np.int64)[None, :] + offs[:, None]).reshape(-1).astype(np.int32)
    idc = np.tile(np.asarray(I_DC, dtype=np.float64), k)
    if jitter_idc:
        rng = np.random.default_rng(seed)
        idc = idc * (1.0 + jitter_idc * (rng.random(idc.size) - 0.5))
    return dict(NeuPar=np.tile(NeuPar, (1, k)), V0=np.tile(np.asarray(V0, dtype=np.float64), (1, k)),
                SynPar=np.tile(np.asarray(SynPar, dtype=np.float64), (1, k)), source_arr=src, target_arr=tgt,
                I_DC=idc, n_per_column=n1)


# BASELINE config 3 (codes/Main_poisson.py:29-30, 145-150; codes/Main_regular.py:29-30, 165-172): dt = 0.01 ms, 4 s,
# Poisson = 100 sources x 30 Hz x 100 ms, 2 nS, pCon 0.1 onto PC_L23 at 1000 ms and onto PC_L5 at 2500 ms;
# regular = 1 source, 250 spikes in 5 ms at 1000 ms and 500 spikes in 5 ms at 2500 ms, 0.1 nS, pCon 0.1 onto PC_L23.
POISSON_STIMULI = [[100, 1, 30, 2, 0, 1000, 1100, [[0, 0, 0.1]]], [100, 1, 30, 2, 0, 2500, 2600, [[7, 0, 0.1]]]]
REGULAR_STIMULI = [[1, 1, 250, 0.1, 0, 1000, 1005, [[0, 0, 0.1]]], [1, 1, 500, 0.1, 0, 2500, 2505, [[0, 0, 0.1]]]]


def shift_stimuli(stimuli_list, factor):
    """The same stimulus sets with their times scaled (tests run a shortened protocol)."""
    return [[a, b, c, d, e, t0 * factor, t0 * factor + (t1 - t0), tg] for a, b, c, d, e, t0, t1, tg in stimuli_list]


def _one_instance_stimulus(job):
    q, seed, use_poisson, protocol, group_distr, dt = job
    from . import stimuli as st
    state = np.random.get_state()
    try:
        np.random.seed(int(seed) + q)
        s = st.draw_poisson_input(protocol, group_distr, dt) if use_poisson else st.draw_regular_input(protocol, group_distr, dt)
    finally:
        np.random.set_state(state)
    st.check_one_spike_per_step(s["spk_src"], s["spk_t"], dt)
    return s


def draw_instance_stimuli(group_distr, n_instances, n_per_instance, dt, poisson=None, regular=None, seed=0, workers=None):
    """The external spike sources of BASELINE config 3 for `n_instances` independent copies of one column: instance q draws
    with np.random.seed(seed + q) exactly as a separate run of codes/Main_poisson.py / Main_regular.py would (stimuli.draw_*).
    Even instances get the Poisson protocol, odd ones the regular protocol when both are given.  Instances are independent,
    so they are drawn by a pool of worker processes (256 instances: 30 s of numpy shuffles on one core).
    -> stimulus dict for `attach_stimuli` (global source / target ids)."""
    from . import stimuli as st
    jobs = []
    for q in range(int(n_instances)):
        use_poisson = poisson is not None and (regular is None or q % 2 == 0)
        jobs.append((q, seed, use_poisson, poisson if use_poisson else regular, group_distr, dt))
    import os
    workers = min(len(jobs), workers or min(32, usable_cpus()))
    if workers > 1 and len(jobs) >= 8:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            drawn = pool.map(_one_instance_stimulus, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
    else:
        drawn = [_one_instance_stimulus(j) for j in jobs]
    parts, n_src = [], 0
    for q, s in enumerate(drawn):
        parts.append(dict(spk_src=s["spk_src"] + n_src, spk_t=s["spk_t"], syn_src=s["syn_src"] + n_src,
                          syn_tgt=s["syn_tgt"] + q * n_per_instance, mask=st.channel_mask(s["syn_type"]), gmax=s["syn_gmax"],
                          pfail=np.full(len(s["syn_src"]), s["pfail"])))
        n_src += s["n_sources"]
    stim = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    stim["n_sources"] = n_src
    return stim


def instances_with_stimuli(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, group_distr, n_instances, dt,
                           poisson=None, regular=None, seed=0):
    """BASELINE config 3: `n_instances` independent copies of one column in ONE engine (no synapse between copies), each
    with its OWN draw of the Poisson / regular stimuli (draw_instance_stimuli).
    -> (network dict of tile_columns, stimulus dict for Engine.add_spike_source with global source / target ids)."""
    net = tile_columns(NeuPar, V0, SynPar, source_arr, target_arr, I_DC, n_instances)
    return net, draw_instance_stimuli(group_distr, n_instances, net["n_per_column"], dt, poisson, regular, seed)


def attach_stimuli(engine, stim):
    engine.add_spike_source(stim["n_sources"], stim["spk_src"], stim["spk_t"], stim["syn_src"], stim["syn_tgt"], stim["mask"],
                            stim["gmax"], stim["pfail"])


def tile_synapses_device(SynPar, source_arr, target_arr, n_columns, n_per_column, device, id_offset=0):
    """The synapse arrays of `tile_columns`, built as torch CUDA tensors directly in HBM (PyTorch is used for device
    buffers only): -> (src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms) for Engine.set_synapses_dev.
    1000 columns of the 1003-neuron column are 2.9e8 synapses = 17 GB, which should not be staged through host memory."""
    import torch
    S = torch.as_tensor(np.ascontiguousarray(SynPar, dtype=np.float64), device=device)
    k = int(n_columns)
    offs = (torch.arange(k, device=device, dtype=torch.int64) * int(n_per_column) + int(id_offset))[:, None]
    src1 = torch.as_tensor(np.asarray(source_arr, dtype=np.int64), device=device)[None, :]
    tgt1 = torch.as_tensor(np.asarray(target_arr, dtype=np.int64), device=device)[None, :]
    src = (src1 + offs).reshape(-1).to(torch.int32).contiguous()
    tgt = (tgt1 + offs).reshape(-1).to(torch.int32).contiguous()
    chan = (S[0].to(torch.int64) - 1).to(torch.uint8).repeat(k).contiguous()
    rows = [S[r].repeat(k).contiguous() for r in (4, 1, 2, 3, 6, 5)]   # g_max, U, tau_rec, tau_fac, p_fail, delay
    return (src, tgt, chan) + tuple(rows)


def _clipped(torch, draw, lo, hi, gen, device):
    """The reference's `rand_par` rule (codes/NetworkSetup.py:621-643): draws outside [lo, hi] are replaced by uniform ones."""
    bad = (draw < lo) | (draw > hi)
    return torch.where(bad, lo + (hi - lo) * torch.rand(draw.shape, generator=gen, device=device, dtype=torch.float64), draw)


def column_chunk(n_columns):
    """Columns per random-number chunk of `inter_column_synapses_device`: the draws of a chunk depend only on (seed, chunk
    index), so every partition whose boundaries are multiples of the chunk builds the SAME network (1000 or 4000 columns on
    1, 2, 4 or 8 GPUs: chunks of 25)."""
    return 25 if n_columns % 200 == 0 else (5 if n_columns % 40 == 0 else 1)


def inter_column_synapses_device(groups, n_per_column, n_columns, device, col_lo=0, col_hi=None, seed=0):
    """Inter-column synapses of a ring of `n_columns` identical columns, for the TARGET columns [col_lo, col_hi), drawn on
    `device` (a CUDA device, or "cpu" for the CPU arm of bench.py) with the reference's rules (codes/ParamSetup.py:575-587,
    codes/NetworkSetup.py:267-321): PC_L23 -> PC_L23 at
    +-2 and +-4 columns, IN_CC_L5 -> PC_L5 and IN_F_L5 -> PC_L5 at +-1, connection probability and peak conductance
    falling off as exp(-d/0.909), delay mean and spread growing linearly with d (up to ~15 ms, i.e. beyond the refractory
    period), AMPA synapses with an NMDA twin of the same delay, STP classes mixed as within a column.
    `groups[g]` = template-local neuron ids of cell group g.  Pairs are drawn WITH replacement (< 1 % duplicates), which
    the reference's own blocks also contain; this is the synthetic stand-in for SURVEY §8f-1, not a bit-exact port.
    The random stream is restarted per chunk of `column_chunk(n_columns)` target columns, so the network does not depend on
    how it is partitioned (round 1 seeded by col_lo: 1 GPU and 8 GPUs simulated different networks).
    -> (src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms) as torch tensors on `device`, ids global."""
    import contextlib
    import io
    import torch
    from .codes.ParamSetup import ParamSetup
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParamSetup(1000, 1, ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14))))
    col_hi = n_columns if col_hi is None else col_hi
    chunk = column_chunk(n_columns)
    if col_lo % chunk or col_hi % chunk:
        raise ValueError("partition boundaries must be multiples of %d columns for a %d-column ring" % (chunk, n_columns))
    parts = [_inter_column_chunk(torch, ps, groups, n_per_column, n_columns, device, c, c + chunk, seed) for c in range(col_lo, col_hi, chunk)]
    parts = [p for p in parts if p is not None]
    if not parts:
        return None
    return tuple(torch.cat([p[q] for p in parts]).contiguous() for q in range(9))


def inter_column_count(groups, n_columns):
    """How many synapses inter_column_synapses_device draws for a whole ring of `n_columns` columns (counts are fixed by the
    rules: round(pCon exp(-d/0.909) N_target N_source) pairs per target column, rule and offset, twice for AMPA + NMDA)."""
    import contextlib
    import io
    from .codes.ParamSetup import ParamSetup
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParamSetup(1000, 1, ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14))))
    SynTypes, pCon, get_SynType = ps[8], ps[14], ps[27]
    rules, coef, offsets = ps[31]
    per_column = 0
    for q, (ia, ja) in enumerate(rules):
        for off in offsets[q]:
            n = int(round(pCon[ia, ja] * float(np.exp(-abs(off) / coef[q][0])) * len(groups[ia]) * len(groups[ja]), 0))
            if n and n_columns > abs(off):
                per_column += n * len(get_SynType[int(SynTypes[ia, ja])])
    return per_column * int(n_columns)


def _inter_column_chunk(torch, ps, groups, n_per_column, n_columns, device, col_lo, col_hi, seed):
    SynTypes, Syngmax, Syndelay, STSP_prob, STSP_types, pCon, S_sig = ps[8], ps[9], ps[10], ps[12], ps[13], ps[14], ps[16]
    STSP_setup, get_SynType, get_STSP_prob, get_STSP_types = ps[26], ps[27], ps[28], ps[29]
    rules, coef, offsets = ps[31]
    ncol = col_hi - col_lo
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1000003 + 7919 * (col_lo // column_chunk(n_columns)) + 1)
    cols = torch.arange(col_lo, col_hi, device=device, dtype=torch.int64)[:, None]
    out = [[] for _ in range(9)]
    f64 = dict(generator=gen, device=device, dtype=torch.float64)
    for q, (ia, ja) in enumerate(rules):
        tg = torch.as_tensor(np.asarray(groups[ia], dtype=np.int64), device=device)
        sr = torch.as_tensor(np.asarray(groups[ja], dtype=np.int64), device=device)
        if len(tg) == 0 or len(sr) == 0:
            continue
        for off in offsets[q]:
            dist = abs(off)
            fall = float(np.exp(-dist / coef[q][0]))
            n = int(round(pCon[ia, ja] * fall * len(tg) * len(sr), 0))
            if n == 0 or n_columns <= dist:
                continue
            t_loc = tg[torch.randint(len(tg), (ncol, n), generator=gen, device=device)]
            s_loc = sr[torch.randint(len(sr), (ncol, n), generator=gen, device=device)]
            tgt = (cols * n_per_column + t_loc).reshape(-1)
            src = (((cols - off) % n_columns) * n_per_column + s_loc).reshape(-1)      # target column = source column + off
            m = tgt.numel()
            d_mean, d_sig = float(Syndelay[ia, ja] + coef[q][1] * dist), float(S_sig[ia, ja, 1] + coef[q][1] * dist)
            delay = _clipped(torch, d_mean + d_sig * torch.randn(m, **f64), 0.0, 2.0 * d_mean, gen, device)
            g_mean, g_sig = float(Syngmax[ia, ja]) * fall, float(S_sig[ia, ja, 0]) * fall
            mu, sd = np.log(g_mean ** 2 / np.sqrt(g_sig ** 2 + g_mean ** 2)), np.sqrt(np.log(g_sig ** 2 / g_mean ** 2 + 1))
            probs = np.asarray(get_STSP_prob[int(STSP_prob[ia, ja])])
            classes = get_STSP_types[int(STSP_types[ia, ja])]
            for stype in get_SynType[int(SynTypes[ia, ja])]:             # 1 AMPA (+ 3 NMDA twin) or 2 GABA
                gmax = _clipped(torch, torch.exp(mu + sd * torch.randn(m, **f64)), 0.0, 100.0 * g_mean, gen, device)
                cls = torch.multinomial(torch.as_tensor(probs, device=device, dtype=torch.float64), m, replacement=True, generator=gen)
                U = torch.empty(m, device=device, dtype=torch.float64)
                trec = torch.empty_like(U)
                tfac = torch.empty_like(U)
                for c, name in enumerate(classes):
                    tab = STSP_setup[name]
                    sel = cls == c
                    k = int(sel.sum())
                    U[sel] = _clipped(torch, tab[0, 0] + tab[1, 0] * torch.randn(k, **f64), 0.0, 1.0, gen, device)
                    trec[sel] = _clipped(torch, tab[0, 1] + tab[1, 1] * torch.randn(k, **f64), 0.0, 1500.0, gen, device)
                    tfac[sel] = _clipped(torch, tab[0, 2] + tab[1, 2] * torch.randn(k, **f64), 0.0, 1500.0, gen, device)
                trec.clamp_(min=1e-3)
                tfac.clamp_(min=1e-3)
                for lst, val in zip(out, (src.to(torch.int32), tgt.to(torch.int32), torch.full((m,), stype - 1, device=device, dtype=torch.uint8),
                                          gmax, U, trec, tfac, torch.full((m,), 0.3, **{k2: v for k2, v in f64.items() if k2 != "generator"}), delay)):
                    lst.append(val)
    if not out[0]:
        return None
    return tuple(torch.cat(x).contiguous() for x in out)


def multicolumn_synapses_device(SynPar, source_arr, target_arr, groups, n_columns, n_per_column, device, col_lo=0, col_hi=None, seed=0):
    """Tiled intra-column synapses + drawn inter-column synapses for the target columns [col_lo, col_hi) of a ring."""
    import torch
    col_hi = n_columns if col_hi is None else col_hi
    intra = tile_synapses_device(SynPar, source_arr, target_arr, col_hi - col_lo, n_per_column, device, id_offset=col_lo * n_per_column)
    inter = inter_column_synapses_device(groups, n_per_column, n_columns, device, col_lo, col_hi, seed)
    if inter is None:
        return intra
    return tuple(torch.cat([a, b]).contiguous() for a, b in zip(intra, inter))


# ---- many DISTINCT columns ---------------------------------------------------------------------------------
# Column c of a scaled network = what the reference's NetworkSetup(N1, 1 stripe) draws after np.random.seed(column_seed(seed, c)):
# its own membrane parameters, I_ref, re-typed interneurons, connectivity (common-neighbour rule for PC -> PC included) and
# synaptic parameters — the bit-exact port of codes/NetworkSetup.py with the vectorised per-neuron path (codes/fast_neuron.py:
# I_ref within 1e-4 pA of the reference's value, everything else bit-identical, tests/test_fast_neuron.py).  Columns are
# independent draws, so they are built by a pool of worker processes and do not depend on how the network is partitioned; the
# inter-column synapses come from inter_column_synapses_device (the cell groups that project across columns — PC_L23, PC_L5,
# IN_CC_L5, IN_F_L5 — are not re-typed, so their template-local ids are the same in every column).
def column_seed(seed, c):
    return (int(seed) * 1000003 + int(c)) % (2 ** 32)


def _build_column(job):
    seed, c, n1, cache_dir = job
    import contextlib
    import io
    import os
    import warnings
    path = os.path.join(cache_dir, "col_n%d_s%d_c%05d.npz" % (n1, seed, c)) if cache_dir else None
    if path and os.path.exists(path):
        try:
            with np.load(path) as d:
                return {k: d[k] for k in d.files}
        except Exception:
            pass
    from .codes.NetworkSetup import sparse_network
    state = np.random.get_state()
    try:
        np.random.seed(column_seed(seed, c))
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            NeuPar, V0, _STyp, SynPar, (tgt, src), gd = sparse_network(n1, 1, ([1, 1, 1, 1], [1, 1, 1, 1], np.ones((10, 14))), True, fast=True)
    finally:
        np.random.set_state(state)
    col = dict(NeuPar=NeuPar, V0=V0, src=np.asarray(src, dtype=np.int32), tgt=np.asarray(tgt, dtype=np.int32),
               chan=(SynPar[0].astype(np.int64) - 1).astype(np.uint8), gmax=SynPar[4], U=SynPar[1], tau_rec=SynPar[2], tau_fac=SynPar[3],
               pfail=SynPar[6], delay=SynPar[5],
               group_sizes=np.asarray([len(gd[g][0]) for g in range(len(gd))], dtype=np.int32),
               group_flat=np.asarray([x for g in range(len(gd)) for x in gd[g][0]], dtype=np.int32))
    if path:
        try:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = path + ".%d.tmp.npz" % os.getpid()
            np.savez(tmp, **col)
            os.replace(tmp, path)
        except Exception:
            pass
    return col


def build_distinct_columns(n1, col_lo, col_hi, seed=0, workers=None, cache_dir=None):
    """Columns [col_lo, col_hi) of the scaled network as host arrays (list of dicts, see _build_column).  Call BEFORE CUDA /
    NCCL are initialised in this process: the pool forks.  A 1003-neuron column takes 1.5 s of one core."""
    import os
    jobs = [(seed, c, n1, cache_dir) for c in range(col_lo, col_hi)]
    workers = min(len(jobs), workers or int(os.environ.get("HHD_SETUP_WORKERS", "0")) or usable_cpus())
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            return pool.map(_build_column, jobs, chunksize=1)
    return [_build_column(j) for j in jobs]


def distinct_columns_network(columns, i_exc=250.0, i_inh=200.0):
    """Neuron-side arrays of a list of built columns: dict(NeuPar[12, N], V0[2, N], I_DC[N], n_per_column, groups) with the
    background current of codes/Main.py:109-112 (PC groups 0 and 7: i_exc, interneurons: i_inh)."""
    n1 = columns[0]["NeuPar"].shape[1]
    idc = []
    for col in columns:
        sizes, flat = col["group_sizes"], col["group_flat"]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        cur = np.full(n1, i_inh)
        for g in (0, 7):
            cur[flat[offs[g]:offs[g + 1]]] = i_exc
        idc.append(cur)
    c0 = columns[0]
    offs = np.concatenate([[0], np.cumsum(c0["group_sizes"])])
    groups = [c0["group_flat"][offs[g]:offs[g + 1]].astype(np.int64) for g in range(len(c0["group_sizes"]))]
    return dict(NeuPar=np.concatenate([c["NeuPar"] for c in columns], axis=1), V0=np.concatenate([c["V0"] for c in columns], axis=1),
                I_DC=np.concatenate(idc), n_per_column=n1, groups=groups)


def distinct_synapses_device(columns, col_lo, n_columns, device, inter_column=True, seed=0):
    """The synapse arrays of the built columns [col_lo, col_lo + len(columns)) as torch tensors on `device` for
    Engine.set_synapses_dev: every column's own intra-column synapses (global ids) + the drawn inter-column synapses of the ring."""
    import torch
    n1 = columns[0]["NeuPar"].shape[1]
    keys = ("src", "tgt", "chan", "gmax", "U", "tau_rec", "tau_fac", "pfail", "delay")
    parts = {k: [] for k in keys}
    for q, col in enumerate(columns):
        off = (col_lo + q) * n1
        for k in keys:
            t = torch.as_tensor(np.ascontiguousarray(col[k]), device=device)
            if k in ("src", "tgt"):
                t = (t + off).to(torch.int32)
            parts[k].append(t)
    intra = tuple(torch.cat(parts[k]).contiguous() for k in keys)
    parts = None
    if not inter_column:
        return intra
    offs = np.concatenate([[0], np.cumsum(columns[0]["group_sizes"])])
    groups = [columns[0]["group_flat"][offs[g]:offs[g + 1]].astype(np.int64) for g in range(len(columns[0]["group_sizes"]))]
    inter = inter_column_synapses_device(groups, n1, n_columns, device, col_lo, col_lo + len(columns), seed)
    if inter is None:
        return intra
    return tuple(torch.cat([a, b]).contiguous() for a, b in zip(intra, inter))
