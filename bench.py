#!/usr/bin/env python
"""bench.py — throughput of the simulation hot path (BASELINE.json metric: neuron-updates/s + synaptic events/s,
realtime factor) on B200, with the HBM roofline of the state-update kernel and a CPU baseline timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload columns|column]
                    [--columns C] [--method euler|rk2|rk4] [--window S]

Default workload: BASELINE config 5 on one GPU — a ring of 1000 columns (1 003 000 neurons): the reference's default column
(bit-exact NetworkSetup output, seed 0) tiled 1000 times plus inter-column synapses drawn by the reference's rules
(3.16e8 synapses, delays up to 15 ms); `--workload columns` leaves the inter-column synapses out, `--workload column` is
the single 1003-neuron column (configs 0-1, cluster-resident plan).

A bench "step" = one window of S consecutive time steps (default 100 ms of biological time) of the whole network:
`value` times hhd_run on the device with everything resident in HBM; `e2e` times the user-facing call sequence with
host buffers — upload this window's background current (H2D), run, read back the window's spikes, the summed-current
LFP trace and the counters (D2H).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_NEURON_UPDATE = 240.0   # DESIGN.md §4 / SURVEY §8d: 8 state r+w (128) + 12 params (96) + last_spike, lastspike, flag (16)
BYTES_PER_SYN_EVENT = 96.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="multicolumn", choices=["columns", "column", "multicolumn"],
                    help="columns: tiled independent columns; multicolumn: ring of columns with inter-column synapses; column: one column")
    ap.add_argument("--columns", type=int, default=0, help="columns in the scaled network (0 = workload default)")
    ap.add_argument("--template", default="n1000", choices=["n1000", "n200"], help="one-column fixture that is tiled")
    ap.add_argument("--method", default="euler", choices=["euler", "rk2", "rk4", "rk23", "gsl_rk2"])
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--window", type=int, default=2000, help="time steps per bench step")
    ap.add_argument("--accum", default="fp64", choices=["fp64", "fixed"])
    ap.add_argument("--plan", default="auto", choices=["auto", "graph", "resident"])
    ap.add_argument("--idc-scale", type=float, default=1.0, help="scale the background current (0: silent network, timing experiments)")
    ap.add_argument("--instances", action="store_true", help="treat every column as an independent instance (config 2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the rk4 / high-activity / one-column legs")
    ap.add_argument("--distinct", type=int, default=-1, help="distinct column draws in the scaled network (0: one template tiled; -1: default)")
    ap.add_argument("--cache-dir", default=os.environ.get("HHD_COLUMN_CACHE", "/tmp/hhd_columns"), help="built columns are kept here between runs on one box")
    ap.add_argument("--no-config3", action="store_true", help="skip the 256-instance / dt 0.01 ms / 4 s leg (BASELINE config 3, ~20 s)")
    ap.add_argument("--config3-only", action="store_true", help="run only the config-3 leg and print its JSON")
    ap.add_argument("--high-idc-scale", type=float, default=1.5, help="background-current scale of the high-activity leg")
    return ap.parse_args()


# ---- workload ----------------------------------------------------------------------------------------------------
def one_column(template):
    """The one-column template: the reference's own NetworkSetup output (bit-exact golden fixture)."""
    from oracle_binding import default_idc, load_golden   # fixture loader only (no oracle code runs here)
    name = {"n1000": "net_n1000_s0_full", "n200": "net_n200_s0"}[template]
    if not os.path.exists(os.path.join(ROOT, "tests", "golden", name + ".npz")):
        name = "net_n200_s0"
    g = load_golden(name)
    from oracle_binding import golden_groups
    return dict(NeuPar=g["NeuPar"], V0=g["V0"], STypPar=g["STypPar"], SynPar=g["SynPar"], source_arr=g["source_arr"],
                target_arr=g["target_arr"], I_DC=default_idc(g), groups=[grp[0] for grp in golden_groups(g)],
                name="%s: %d neurons, %d synapses" % (name, g["NeuPar"].shape[1], g["SynPar"].shape[1]))


_DRAWS = {}
_DISTINCT_OK = [True]         # cleared when the column build fails: every leg then uses the tiled template
DEFAULT_DISTINCT = 1000      # column c of the scaled network is draw (c mod DISTINCT): with 1000 every column of the default workload is its own draw


def distinct_draws(args):
    if args.workload == "column" or args.instances or args.template != "n1000" or not _DISTINCT_OK[0]:
        return 0
    d = DEFAULT_DISTINCT if args.distinct < 0 else args.distinct
    d = min(d, args.columns or 1000)
    if args.distinct < 0:
        # a column takes 1.6 s of one core: keep the default build under ~150 s on hosts with few cores (16 cores build all 1000);
        # the limit depends on the host only, so both arms and every rank count the same
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import usable_cpus
        d = min(d, max(25, int(150 * usable_cpus() / 1.6) // 25 * 25))
    return d


def prebuild_columns(args, rank, world, ncol=None):
    """The distinct column draws this rank needs, built by forked workers BEFORE CUDA / NCCL exist in this process
    (scaled.build_distinct_columns: the reference's generator per column, seed = f(column), fast per-neuron path)."""
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import build_distinct_columns
    D = distinct_draws(args)
    if not D:
        return None
    ncol = ncol or args.columns or 1000
    per_rank = ncol // world
    need = sorted({p % D for p in range(rank * per_rank, (rank + 1) * per_rank)})
    t0 = time.perf_counter()
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import usable_cpus
    workers = max(1, usable_cpus() // max(1, world))
    runs, start = [], 0                                        # contiguous runs of draw indices -> one pool call each
    while start < len(need):
        end = start
        while end + 1 < len(need) and need[end + 1] == need[end] + 1:
            end += 1
        runs.append((need[start], need[end] + 1))
        start = end + 1
    draws = {}
    for lo, hi in runs:
        if all(c in _DRAWS for c in range(lo, hi)):            # built earlier in this process (the CPU sample reuses the GPU arm's columns)
            draws.update({c: _DRAWS[c] for c in range(lo, hi)})
            continue
        for c, col in zip(range(lo, hi), build_distinct_columns(1000, lo, hi, seed=0, workers=workers, cache_dir=args.cache_dir or None)):
            draws[c] = col
    _DRAWS.update(draws)
    return dict(draws=draws, D=D, seconds=time.perf_counter() - t0, workers=workers)


def build_workload(args, rank, world, device_synapses=False, prebuilt=None):
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import tile_columns
    col = one_column(args.template)
    col["I_DC"] = col["I_DC"] * args.idc_scale
    if prebuilt is not None and args.workload != "column":
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import distinct_columns_network
        ncol = args.columns or 1000
        per_rank = ncol // world
        cols = [prebuilt["draws"][p % prebuilt["D"]] for p in range(rank * per_rank, (rank + 1) * per_rank)]
        net = distinct_columns_network(cols)
        net["I_DC"] = net["I_DC"] * args.idc_scale
        net.update(SynPar=np.zeros((7, 0)), source_arr=[], target_arr=[], distinct_cols=cols, col=col, STypPar=col["STypPar"],
                   nsyn=int(sum(len(c["src"]) for c in cols)), columns_total=per_rank * world, columns_rank=per_rank, first_column=rank * per_rank,
                   template="%d distinct draws of the reference's 1003-neuron column, NetworkSetup(1000, 1) with one numpy seed per column"
                            % min(prebuilt["D"], ncol),
                   column_build=dict(seconds=prebuilt["seconds"], workers=prebuilt["workers"], columns_built=len(prebuilt["draws"])))
        return net
    if args.workload == "column":
        ncol = 1
    else:
        ncol = args.columns or 1000
    per_rank = ncol // world if args.workload != "column" else 1
    if device_synapses:
        # neurons on the host (96 B each), synapses generated straight into HBM by make_engine
        net = tile_columns(col["NeuPar"], col["V0"], np.zeros((7, 0)), [], [], col["I_DC"], per_rank)
        net["col"] = col
    else:
        net = tile_columns(col["NeuPar"], col["V0"], col["SynPar"], col["source_arr"], col["target_arr"], col["I_DC"], per_rank)
    net["nsyn"] = int(col["SynPar"].shape[1]) * per_rank
    net["STypPar"] = col["STypPar"]
    net["columns_total"] = per_rank * world
    net["columns_rank"] = per_rank
    net["template"] = col["name"]
    return net


def make_engine(cls, net, args, seed, host=False, **kw):
    from rescience_hass_hertag_durstewitz_2016_b200.engine import load_codes_network
    n = net["NeuPar"].shape[1]
    if host and "distinct_cols" in net:
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import distinct_synapses_device
        e = cls(n, args.dt, method=args.method, accum=args.accum, seed=seed, **kw)
        load_codes_network(e, net["NeuPar"], net["V0"], net["STypPar"], net["SynPar"], net["source_arr"], net["target_arr"], net["I_DC"])
        syn = distinct_synapses_device(net["distinct_cols"], 0, net["columns_total"], "cpu", inter_column=args.workload == "multicolumn", seed=1)
        net["nsyn"] = int(syn[0].numel())
        e.set_synapses(*[t.numpy() for t in syn])
        return e
    if host and "col" in net:
        # CPU arm: the same generators, tensors on the host
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import multicolumn_synapses_device, tile_synapses_device
        col = net["col"]
        e = cls(n, args.dt, method=args.method, accum=args.accum, seed=seed, **kw)
        load_codes_network(e, net["NeuPar"], net["V0"], net["STypPar"], net["SynPar"], net["source_arr"], net["target_arr"], net["I_DC"])
        if args.workload == "multicolumn":
            syn = multicolumn_synapses_device(col["SynPar"], col["source_arr"], col["target_arr"], col["groups"], net["columns_total"],
                                              net["n_per_column"], "cpu", 0, net["columns_total"], seed=1)
        else:
            syn = tile_synapses_device(col["SynPar"], col["source_arr"], col["target_arr"], net["columns_total"], net["n_per_column"], "cpu")
        net["nsyn"] = int(syn[0].numel())
        e.set_synapses(*[t.numpy() for t in syn])
        return e
    if cls.__name__ == "Engine":
        kw = dict(kw, plan=args.plan, instance_size=net["n_per_column"] if (args.instances or args.workload == "column") else 0)
    e = cls(n, args.dt, method=args.method, accum=args.accum, seed=seed, **kw)
    load_codes_network(e, net["NeuPar"], net["V0"], net["STypPar"], net["SynPar"], net["source_arr"], net["target_arr"], net["I_DC"])
    if "distinct_cols" in net:
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import distinct_synapses_device
        dev = "cuda:%d" % kw.get("device", 0)
        if kw.get("n_ranks", 1) > 1:                          # every neuron with a synapse somewhere in the ring
            e.set_has_outgoing(np.ones(kw["n_global"], dtype=np.uint8))
        syn = distinct_synapses_device(net["distinct_cols"], net["first_column"], net["columns_total"], dev,
                                       inter_column=args.workload == "multicolumn", seed=1)
        net["nsyn"] = int(syn[0].numel())
        e.set_synapses_dev(*syn, syn_id_base=kw.get("rank", 0) * (1 << 31))
        return e
    if "col" in net:
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import tile_synapses_device
        col = net["col"]
        dev = "cuda:%d" % kw.get("device", 0)
        off = kw.get("global_offset", 0)
        if kw.get("n_ranks", 1) > 1:
            has_out = np.zeros(net["n_per_column"], dtype=np.uint8)
            has_out[np.unique(col["source_arr"])] = 1
            e.set_has_outgoing(np.tile(has_out, kw["n_global"] // net["n_per_column"]))
        if args.workload == "multicolumn":
            from rescience_hass_hertag_durstewitz_2016_b200.scaled import multicolumn_synapses_device
            c0 = off // net["n_per_column"]
            syn = multicolumn_synapses_device(col["SynPar"], col["source_arr"], col["target_arr"], col["groups"], net["columns_total"],
                                              net["n_per_column"], dev, c0, c0 + net["columns_rank"], seed=1)
            net["nsyn"] = int(syn[0].numel())
            if kw.get("n_ranks", 1) > 1:                      # every PC / IN_CC_L5 / IN_F_L5 cell projects somewhere
                e.set_has_outgoing(np.ones(kw["n_global"], dtype=np.uint8))
            e.set_synapses_dev(*syn, syn_id_base=kw.get("rank", 0) * (1 << 30))
        else:
            e.set_synapses_dev(*tile_synapses_device(col["SynPar"], col["source_arr"], col["target_arr"], net["columns_rank"],
                                                     net["n_per_column"], dev, id_offset=off),
                               syn_id_base=kw.get("rank", 0) * net["nsyn"])
    return e


# ---- clocks ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the CPU arm ---------------------------------------------------------------------------------------------------
CPU_SAMPLE_COLUMNS = 40      # columns of the CPU sample of a multi-column workload (40 130 neurons, 1.26e7 synapses)


class CpuSample:
    """The CPU oracle (kind 'port': Brian 2, the reference's engine, cannot be installed here — the probe below says so in
    the JSON line) on a bounded sample of the SAME workload: for the ring of columns a ring of CPU_SAMPLE_COLUMNS columns
    with inter-column synapses drawn by the same generator (scaled.multicolumn_synapses_device on the host), for the
    single column the column itself.  CPU cost is linear in neurons, so neuron-updates/s of the sample is the figure for
    the whole network.  Built once; every call continues the same simulation."""

    def __init__(self, args, threads):
        os.environ["OMP_NUM_THREADS"] = str(threads)
        from oracle_binding import OracleEngine, oracle_lib
        import copy
        self.threads = threads
        self.lib = oracle_lib()
        self.lib.omp_set_num_threads(int(threads))     # the environment variable is only read when libgomp is loaded
        a2 = copy.copy(args)
        if args.workload != "column":
            a2.columns = min(args.columns or 1000, CPU_SAMPLE_COLUMNS)
        self.args = a2
        net = build_workload(a2, 0, 1, device_synapses=args.workload != "column",
                             prebuilt=prebuild_columns(a2, 0, 1, ncol=a2.columns) if args.workload != "column" else None)
        t0 = time.perf_counter()
        self.e = make_engine(OracleEngine, net, a2, seed=1, record_spikes=False, host=True)
        self.n = net["NeuPar"].shape[1]
        self.net = net
        self.setup_s = time.perf_counter() - t0
        self.e.run(3)                                   # first touch

    def probe(self, steps=12):
        self.lib.omp_set_num_threads(int(self.threads))
        t0 = time.perf_counter()
        self.e.run(steps)
        return (time.perf_counter() - t0) / steps

    def run(self, seconds, max_steps):
        per = max(self.probe(5), 1e-7)
        nsteps = int(max(4, min(max_steps, seconds / per)))
        c0 = self.e.counters()
        t0 = time.perf_counter()
        self.e.run(nsteps)
        el = time.perf_counter() - t0
        c1 = self.e.counters()
        what = ("ring of %d columns with inter-column synapses" % self.net["columns_total"]) if self.args.workload == "multicolumn" else \
               ("%d independent column(s)" % self.net["columns_total"])
        info = dict(value=self.n * nsteps / el, unit="neuron-updates/s", cores=self.threads, kind="port",
                    sample="%d time steps of a %s (%d neurons, %d synapses; same generator as the GPU workload), %.1f s wall, C oracle %s" %
                           (nsteps, what, self.n, self.net["nsyn"], el,
                            ("with OpenMP over neurons, %d threads" % self.threads) if self.threads > 1 else "single thread"),
                    syn_events_per_s=(c1["syn_events"] - c0["syn_events"]) / el, ms_per_time_step=el / nsteps * 1e3)
        return info, nsteps, el

    def close(self):
        self.e.close()


def best_cpu_sample(args):
    """All host threads if they help, else one: the probe decides, and `cores` reports the count that was used."""
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import usable_cpus
    ncpu = usable_cpus()
    one = CpuSample(args, 1)
    if ncpu == 1 or args.workload == "column":
        return one, {"1": one.probe()}
    t1 = one.probe()
    many = CpuSample(args, ncpu)
    tn = many.probe()
    probes = {"1": t1, str(ncpu): tn}
    if tn < t1:
        one.close()
        return many, probes
    many.close()
    return one, probes


def brian2_probe():
    """if Brian 2 is importable on the bench box the reference's own CortexNetwork.run would be the baseline."""
    try:
        import brian2  # noqa: F401
        return "importable (version %s) — the reference itself is not on this box (/root/reference is absent), so it is still the C port that is timed" % brian2.__version__
    except Exception as ex:
        return "not importable (%s): the CPU arm is the C restatement of Brian's schedule (oracle/hhd_oracle.c)" % type(ex).__name__


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    w = 1 if args.workload == "column" else world
    pre = None
    if distinct_draws(args):
        try:
            pre = prebuild_columns(args, 0, 1)
        except Exception as exc:
            _DISTINCT_OK[0] = False
            print("bench.py: building distinct columns failed (%s: %s); falling back to the tiled template column" % (type(exc).__name__, exc), file=sys.stderr)
    if pre is not None:
        # the whole workload's columns are built here too (and left in the column cache for the GPU arm): the synapse count of the
        # config is the real one; the timed sample below is a ring of the first CPU_SAMPLE_COLUMNS of them
        from rescience_hass_hertag_durstewitz_2016_b200.scaled import inter_column_count
        net = build_workload(args, 0, 1, device_synapses=True, prebuilt=pre)
        net["total_synapses"] = net["nsyn"] + (inter_column_count(net["groups"], net["columns_total"]) if args.workload == "multicolumn" else 0)
        net["total_neurons"] = net["NeuPar"].shape[1]
        net["NeuPar"] = net["NeuPar"][:, :net["NeuPar"].shape[1] // w]                  # per rank, as the GPU arm sees it (the l2 note)
        for k in [c for c in _DRAWS if c >= CPU_SAMPLE_COLUMNS]:
            del _DRAWS[k]
        net.pop("distinct_cols")
        net["distinct_cols"] = True
        pre = None
    else:
        net = build_workload(args, 0, w, device_synapses=True)   # config only
        net["nsyn"] = int(net["nsyn"] * (1.0766 if args.workload == "multicolumn" else 1.0))              # + inter-column synapses (measured ratio)
    cs, probes = best_cpu_sample(args)
    # one bench step = one bounded sample; K steps + W warm-ups must end within a few minutes
    budget = max(1.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    vals, ms = [], []
    info = None
    for it in range(args.warmup + args.steps):
        info, nsteps, el = cs.run(budget, args.window * 20)
        if it >= args.warmup:
            vals.append(info["value"])
            ms.append(el * 1e3)
    v = float(np.mean(vals))
    info["value"] = v
    info["thread_probe_s_per_time_step"] = probes
    info["brian2"] = brian2_probe()
    out = dict(metric="neuron_updates_per_s", value=v, unit="neuron-updates/s", n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=float(np.mean(ms)), higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", impl="reference", config=workload_config(args, net, world),
               cpu_baseline=info, e2e=dict(value=v, unit="neuron-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


def workload_config(args, net, world):
    n = net["NeuPar"].shape[1]
    how = "%s" if "distinct_cols" in net else "%s tiled"
    name = {"columns": ("%d-column PFC network (" + how + ", block-diagonal)") % (net["columns_total"], net["template"]),
            "multicolumn": ("ring of %d PFC columns (" + how + ") with inter-column synapses drawn by the reference's rules") % (net["columns_total"], net["template"]),
            "column": "one PFC column (%s)" % net["template"]}[args.workload]
    return dict(workload=name,
                neurons=net.get("total_neurons", n * (world if args.workload != "column" else 1)),
                synapses=net.get("total_synapses", net["nsyn"] * (world if args.workload != "column" else 1)),
                method=args.method, dt_ms=args.dt, time_steps_per_step=args.window, accumulate=args.accum,
                monitors="spikes + summed I_tot (LFP) every time step", background="250/200 pA constant current",
                l2="state+parameters %.0f MB per GPU vs 126 MB L2" % (n * BYTES_PER_NEURON_UPDATE / 1e6),
                parallelism="%d GPU(s), columns partitioned, spike indices exchanged every time step over IPC-mapped peer memory (NCCL all-gather as fallback)" % world if world > 1 else "1 GPU")


# ---- the other halves of the metric: the 1003-neuron column (configs 0-1) and the reference's default integrator -------------
PUBLISHED_BRIAN_MS_PER_STEP = 1.02     # codes_revision/Tests.py:10 (31 s bio + setup + analyses in 10 min 35 s; rk4; hardware unknown)


def column_leg(args, method, device, cpu=True):
    """One PFC column (the reference's own NetworkSetup output, seed 0), 2 s of biological time = 40 000 steps (BASELINE
    configs 0/1), cluster-resident plan, spikes + summed-I_tot monitor; CPU beside it: the C port (1 thread) and the NumPy
    oracle (structurally Brian's numpy target)."""
    import copy
    from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine
    a2 = copy.copy(args)
    a2.workload, a2.method, a2.plan, a2.columns = "column", method, "auto", 0
    net = build_workload(a2, 0, 1)
    n = net["NeuPar"].shape[1]
    steps = int(2000.0 / a2.dt)
    eng = make_engine(Engine, net, a2, seed=1, device=device)
    lfp = eng.add_state_monitor("I_tot", None, reduce="sum")
    eng.run(steps // 4)                                       # warm-up: 500 ms, the network leaves rest
    eng.clear_monitor(lfp)
    eng.clear_monitor(-1)
    c0 = eng.counters()
    eng.run(steps)
    c1 = eng.counters()
    dev_s = (c1["run_ms_device"] - c0["run_ms_device"]) * 1e-3
    eng.clear_monitor(lfp)
    eng.clear_monitor(-1)
    t0 = time.perf_counter()                                  # end to end: host buffers in and out
    eng.set_idc(net["I_DC"])
    eng.run(steps)
    i, sp = eng.spikes()
    trace = eng.read_monitor(lfp)
    e2e_s = time.perf_counter() - t0
    eng.close()
    out = dict(workload="one PFC column (%s), 2 s biological time" % net["template"], method=method, plan="cluster-resident (k_resident)",
               neurons=n, synapses=net["nsyn"], time_steps=steps, us_per_time_step=dev_s / steps * 1e6,
               neuron_updates_per_s=n * steps / dev_s, syn_events_per_s=(c1["syn_events"] - c0["syn_events"]) / dev_s,
               realtime_factor=2.0 / dev_s, mean_rate_hz=(c1["spikes"] - c0["spikes"]) / n / 2.0,
               e2e=dict(us_per_time_step=e2e_s / steps * 1e6, neuron_updates_per_s=n * steps / e2e_s, h2d_bytes=int(n * 8),
                        d2h_bytes=int(len(i) * 12 + trace.nbytes)),
               speedup_vs_published_brian_1p02_ms_per_step=(PUBLISHED_BRIAN_MS_PER_STEP * 1e3) / (dev_s / steps * 1e6) if method == "rk4" else None)
    if cpu:
        cs = CpuSample(a2, 1)
        info, nsteps, el = cs.run(3.0, steps)
        cs.close()
        out["cpu_baseline"] = info
        out["speedup_vs_c_port_1_thread"] = out["neuron_updates_per_s"] / info["value"]
        np_us = numpy_oracle_us_per_step(net, a2)
        out["numpy_oracle"] = dict(us_per_time_step=np_us, neuron_updates_per_s=n / (np_us * 1e-6), cores=1, kind="port",
                                   sample="300 time steps of the same column after 100 warm-up steps, oracle/numpy_oracle.py (SI units, one numpy statement per Brian code object)")
        out["speedup_vs_numpy_oracle"] = np_us / out["us_per_time_step"]
    return out


def config3_leg(args, device, n_instances=256, dt=0.01, duration_ms=4000.0, method="rk4"):
    """BASELINE config 3 as stated: 256 independent instances of the 1003-neuron column in ONE engine on one GPU, dt = 0.01 ms,
    4000 ms (codes/Main_poisson.py:29-30), even instances with the Poisson protocol (100 sources x 30 Hz x 100 ms, 2 nS, pCon 0.1
    onto PC_L23 at 1000 ms and PC_L5 at 2500 ms, :145-150), odd ones with the regular protocol (250 / 500 spikes in 5 ms, 0.1 nS,
    codes/Main_regular.py:165-172); each instance draws its own stimulus with its own numpy seed and has its own synapse indices
    (release failures).  Timed: the whole 400 000-step run, device time, after building; replicas only — no exchange."""
    import copy
    from oracle_binding import golden_groups, load_golden
    from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import POISSON_STIMULI, REGULAR_STIMULI, attach_stimuli, draw_instance_stimuli
    a2 = copy.copy(args)
    a2.workload, a2.method, a2.plan, a2.columns, a2.instances, a2.dt = "columns", method, "auto", n_instances, True, dt
    net = build_workload(a2, 0, 1, device_synapses=True)
    n, n1 = net["NeuPar"].shape[1], net["n_per_column"]
    g = load_golden("net_n1000_s0_full")
    groups = [[list(map(int, s)) for s in grp] for grp in golden_groups(g)]
    t0 = time.perf_counter()
    stim = draw_instance_stimuli(groups, n_instances, n1, dt, poisson=POISSON_STIMULI, regular=REGULAR_STIMULI, seed=0)
    eng = make_engine(Engine, net, a2, seed=1, device=device)
    attach_stimuli(eng, stim)
    setup_s = time.perf_counter() - t0
    steps = int(round(duration_ms / dt))
    c0 = eng.counters()
    t0 = time.perf_counter()
    eng.run(steps)
    i, sp = eng.spikes()                                      # the 256 spike monitors, read back once at the end
    wall = time.perf_counter() - t0
    c1 = eng.counters()
    eng.close()
    dev_s = (c1["run_ms_device"] - c0["run_ms_device"]) * 1e-3
    inst = np.asarray(i) // n1
    t_ms = np.asarray(sp) * dt
    in_stim = ((t_ms >= 1000) & (t_ms < 1100)) | ((t_ms >= 2500) & (t_ms < 2600))
    return dict(workload="BASELINE config 3: %d independent instances of %s, dt %.2f ms, %.0f ms, Poisson (even) / regular (odd) stimuli at 1000 and 2500 ms"
                         % (n_instances, net["template"], dt, duration_ms), method=method, neurons=n, synapses=int(net["nsyn"]), time_steps=steps,
                external_sources=int(stim["n_sources"]), external_spikes=int(len(stim["spk_t"])), external_synapses=int(len(stim["syn_src"])),
                us_per_time_step=dev_s / steps * 1e6, neuron_updates_per_s=n * steps / dev_s, syn_events_per_s=(c1["syn_events"] - c0["syn_events"]) / dev_s,
                ext_events=int(c1["ext_events"] - c0["ext_events"]), realtime_factor_per_instance=duration_ms * 1e-3 / dev_s,
                instance_seconds_per_wall_second=n_instances * duration_ms * 1e-3 / dev_s, mean_rate_hz=len(i) / n / (duration_ms * 1e-3),
                rate_hz_inside_stimulus_windows=float(in_stim.sum()) / n / 0.2, rate_hz_outside=float((~in_stim).sum()) / n / (duration_ms * 1e-3 - 0.2),
                spikes_per_instance_min_max=[int(np.bincount(inst, minlength=n_instances).min()), int(np.bincount(inst, minlength=n_instances).max())],
                e2e_wall_s=wall, setup_s=setup_s)


def numpy_oracle_us_per_step(net, a2, steps=300):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from numpy_oracle import NumpyOracle
    from rescience_hass_hertag_durstewitz_2016_b200.engine import load_codes_network
    e = NumpyOracle(net["NeuPar"].shape[1], a2.dt, method=a2.method, seed=1)
    load_codes_network(e, net["NeuPar"], net["V0"], net["STypPar"], net["SynPar"], net["source_arr"], net["target_arr"], net["I_DC"])
    e.run(100)
    t0 = time.perf_counter()
    e.run(steps)
    return (time.perf_counter() - t0) / steps * 1e6


def method_leg(args, method, net, n, device, idc_scale=1.0, windows=2):
    """The default 1M-neuron workload again with another integrator / background current: device-timed windows + kernel times."""
    import copy
    from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine
    a2 = copy.copy(args)
    a2.method = method
    net2 = dict(net)
    net2["I_DC"] = net["I_DC"] * idc_scale
    eng = make_engine(Engine, net2, a2, seed=1, device=device)
    lfp = eng.add_state_monitor("I_tot", None, reduce="sum")
    S = args.window
    for _ in range(2):
        eng.run(S)
        eng.clear_monitor(lfp)
        eng.clear_monitor(-1)
    c0 = eng.counters()
    for _ in range(windows):
        eng.run(S)
        eng.clear_monitor(lfp)
        eng.clear_monitor(-1)
    c1 = eng.counters()
    prof = eng.profile_kernels(min(S, 200))
    eng.close()
    dev_s = (c1["run_ms_device"] - c0["run_ms_device"]) * 1e-3
    ev = c1["syn_events"] - c0["syn_events"]
    steps = S * windows
    return dict(method=method, idc_scale=idc_scale, us_per_time_step=dev_s / steps * 1e6, neuron_updates_per_s=n * steps / dev_s,
                syn_events_per_s=ev / dev_s, mean_rate_hz=(c1["spikes"] - c0["spikes"]) / n / (steps * args.dt * 1e-3),
                realtime_factor=steps * args.dt * 1e-3 / dev_s, us_per_launch=prof, syn_events_per_time_step=ev / steps,
                frac_of_hbm_roofline_whole_step=None)


# ---- our arm -------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    prebuilt, build_error = None, None
    if not args.config3_only:
        try:
            prebuilt = prebuild_columns(args, rank, world)
        except Exception as exc:                               # no bench line is worse than a bench line on the tiled template
            build_error = "%s: %s" % (type(exc).__name__, exc)
            _DISTINCT_OK[0] = False
            print("bench.py: building distinct columns failed (%s); falling back to the tiled template column" % build_error, file=sys.stderr)
    import torch
    import torch.distributed as dist
    from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    if world > 1 and distinct_draws(args):                    # every rank or none
        ok = torch.tensor([0 if build_error else 1], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            prebuilt = None
            _DISTINCT_OK[0] = False

    if args.config3_only:
        print(json.dumps(config3_leg(args, local_rank)), flush=True)
        return
    net = build_workload(args, rank, world, device_synapses=True, prebuilt=prebuilt)
    n = net["NeuPar"].shape[1]
    t_setup = time.perf_counter()
    if world > 1:
        # one network partitioned by columns: same seed on every rank, global ids, spike all-gather inside libhhd (NCCL)
        eng = make_engine(Engine, net, args, seed=1, device=local_rank, rank=rank, n_ranks=world, n_global=n * world,
                          global_offset=rank * n)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(eng.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        eng.comm_init(bytes(uid.cpu().tolist()))
    else:
        eng = make_engine(Engine, net, args, seed=1, device=local_rank)
    lfp = eng.add_state_monitor("I_tot", None, reduce="sum")
    setup_s = time.perf_counter() - t_setup
    S = args.window
    idc = torch.from_numpy(np.ascontiguousarray(net["I_DC"])).pin_memory()   # this window's input, pinned host memory
    idc_np = idc.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def window_device():
        eng.run(S)
        eng.clear_monitor(lfp)
        eng.clear_monitor(-1)

    def window_e2e():
        eng.set_idc(idc_np)                   # H2D: background drive of this window
        eng.run(S)
        i, s = eng.spikes()                   # D2H: the window's spikes (SpikeMonitor.i / .t)
        trace = eng.read_monitor(lfp)         # D2H: summed I_tot per time step
        c = eng.counters()
        eng.clear_monitor(lfp)
        eng.clear_monitor(-1)
        return len(i), trace.nbytes, c

    for _ in range(args.warmup):
        window_device()
    # ---- device-timed leg: inputs resident, CUDA events inside hhd_run on the engine's stream ----------------
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    c0 = eng.counters()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        window_device()
    barrier()
    wall_dev = time.perf_counter() - w0
    c1 = eng.counters()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = c1["run_ms_device"] - c0["run_ms_device"]
    # ---- end-to-end leg ------------------------------------------------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        window_e2e()
    barrier()
    e0 = time.perf_counter()
    nspk_tot, d2h = 0, 0
    for _ in range(args.steps):
        nspk, tb, c2 = window_e2e()
        nspk_tot += nspk
        d2h += nspk * 8 + tb + 128
    barrier()
    e2e_s = time.perf_counter() - e0
    # ---- per-kernel times for the roofline (CUDA events between launches on the engine's stream) ----------------
    prof = eng.profile_kernels(min(S, 200))

    t = torch.tensor([dev_ms, e2e_s * 1e3, float(c1["syn_events"] - c0["syn_events"]), float(c1["spikes"] - c0["spikes"]),
                      float(c1["kernel_launches"] - c0["kernel_launches"]), float(n), float(net["nsyn"])], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = mx[0].item(), mx[1].item()
        syn_ev, spikes, launches, n_total, nsyn_total = sm[2].item(), sm[3].item(), sm[4].item(), sm[5].item(), sm[6].item()
    else:
        dev_ms, e2e_ms, syn_ev, spikes, launches, n_total, nsyn_total = [x.item() for x in t]
    net["total_neurons"], net["total_synapses"] = int(n_total), int(nsyn_total)
    if rank != 0:
        return
    updates = n_total * S * args.steps
    value = updates / (dev_ms * 1e-3)
    e2e_value = updates / (e2e_ms * 1e-3)
    bio_s = S * args.steps * args.dt * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BYTES_PER_NEURON_UPDATE * n / (prof["k_neuron"] * 1e-6) / 1e9
    traffic, traffic_source = None, None
    try:        # DRAM bytes per launch of the same kernel: NOT measured in this run (ncu cannot run inside a timed bench) but
        #         read from the committed `ncu --set full` capture of this very command, and labelled so
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if args.workload == "multicolumn" and n == 1003000 and args.method == "euler" and world == 1:
            traffic = tr["k_neuron"]["dram_bytes_read"] + tr["k_neuron"]["dram_bytes_write"]
            traffic_source = "committed ncu capture: " + tr["source"]
    except Exception:
        tr = {}
    time_steps = S * args.steps
    ev_per_step = syn_ev / time_steps / world                    # per rank and time step
    deliv_bytes = BYTES_PER_SYN_EVENT * ev_per_step
    deliv_achieved = deliv_bytes / (prof["k_synapse"] * 1e-6) / 1e9 if prof["k_synapse"] > 0 else None
    deliv_traffic = None
    if traffic is not None and "k_synapse" in tr:
        deliv_traffic = tr["k_synapse"]["dram_bytes_read"] + tr["k_synapse"]["dram_bytes_write"]
    step_bytes = BYTES_PER_NEURON_UPDATE * n + deliv_bytes
    step_us = dev_ms * 1e3 / time_steps
    out = dict(
        metric="neuron_updates_per_s", value=value, unit="neuron-updates/s", n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
        dtype="f64", data="synthetic", config=workload_config(args, net, world),
        syn_events_per_s=syn_ev / (dev_ms * 1e-3), realtime_factor=bio_s / (dev_ms * 1e-3),
        mean_rate_hz=spikes / n_total / bio_s, us_per_time_step=step_us,
        roofline=dict(bound="hbm", kernel="k_neuron<%s>" % args.method, achieved=achieved, peak=peak, unit="GB/s",
                      frac=achieved / peak, traffic=traffic, traffic_source=traffic_source,
                      peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                      algorithmic_bytes_per_launch=BYTES_PER_NEURON_UPDATE * n,
                      us_per_launch=prof, note="mean of %d launches, CUDA events between kernels on the engine's stream" % min(S, 200)),
        roofline_delivery=dict(bound="hbm", kernel="k_synapse (spike-time STP + delayed delivery)", achieved=deliv_achieved, peak=peak,
                               unit="GB/s", frac=(deliv_achieved / peak) if deliv_achieved else None, traffic=deliv_traffic,
                               traffic_source=traffic_source if deliv_traffic is not None else None,
                               algorithmic_bytes_per_launch=deliv_bytes, syn_events_per_launch=ev_per_step,
                               note="96 B per synaptic event (SURVEY 8d) x events per time step; at %.1f Hz the kernel is latency-bound, not "
                                    "bandwidth-bound — see `high_activity` for the regime where delivery traffic matters" % (spikes / n_total / bio_s)),
        roofline_whole_step=dict(achieved=step_bytes / (step_us * 1e-6) / 1e9 / 1.0, peak=peak, unit="GB/s",
                                 frac=step_bytes / (step_us * 1e-6) / 1e9 / peak,
                                 note="(240 B x neurons + 96 B x synaptic events) per rank / time per time step"),
        e2e=dict(value=e2e_value, unit="neuron-updates/s", h2d_bytes_per_step=int(n * 8), d2h_bytes_per_step=int(d2h / args.steps),
                 ms_per_step=e2e_ms / args.steps),
        gpu_launches=int(launches), clocks=clocks, setup_s=setup_s, wall_s_device_leg=wall_dev, column_build=net.get("column_build"))
    eng.close()
    if world == 1 and not args.no_extras and args.workload != "column":
        # the rest of BASELINE.json's metric on the same box, same process: the reference's default integrator at 1M neurons, a
        # high-activity variant for the delivery kernel, and the 1003-neuron column (euler + rk4) with its own CPU baselines
        extra = {}
        if args.method != "rk4":
            extra["rk4_1M"] = method_leg(args, "rk4", net, n, local_rank)
        extra["high_activity"] = method_leg(args, args.method, net, n, local_rank, idc_scale=args.high_idc_scale)
        for leg in extra.values():
            b = BYTES_PER_NEURON_UPDATE * n + BYTES_PER_SYN_EVENT * leg["syn_events_per_time_step"]
            leg["frac_of_hbm_roofline_whole_step"] = b / (leg["us_per_time_step"] * 1e-6) / 1e9 / peak
            leg["delivery_gbs_algorithmic"] = BYTES_PER_SYN_EVENT * leg["syn_events_per_time_step"] / (leg["us_per_launch"]["k_synapse"] * 1e-6) / 1e9
        extra["column_1k"] = {m: column_leg(args, m, local_rank, cpu=not args.no_cpu_baseline) for m in ("euler", "rk4")}
        if not args.no_config3:
            extra["config3_256_instances"] = config3_leg(args, local_rank)
        out.update(extra)
    if not args.no_cpu_baseline and world == 1:
        cs, probes = best_cpu_sample(args)
        out["cpu_baseline"], _, _ = cs.run(args.cpu_seconds, args.window * 20)
        out["cpu_baseline"]["thread_probe_s_per_time_step"] = probes
        out["cpu_baseline"]["brian2"] = brian2_probe()
        cs.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
