"""ctypes binding of the C ABI in include/hhd.h (libhhd.so, built from csrc/ by __graft_entry__.build()).

This is the thin layer the façades (CortexNetwork / Cortex look-alikes) sit on: it replaces what the reference does
with Brian 2 objects at codes/CortexNetwork.py:308-311, 389-482, 487-512, 570-622, 729.  There is no CPU fallback:
if libhhd.so is missing or no B200 is visible, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HHD_LIB") or os.path.join(_HERE, "csrc", "libhhd.so")   # HHD_LIB: kernel-variant experiments

ABI_VERSION = 1
METHODS = {"euler": 0, "rk2": 1, "rk4": 2, "rk23": 3, "gsl_rk2": 3}
ACCUM = {"fp64": 0, "fixed": 1}
PLANS = {"auto": 0, "graph": 1, "resident": 2}
REDUCE = {"none": 0, "sum": 1, "mean": 2}
PARAM_ROWS = ["C", "g_L", "E_L", "delta_T", "V_up", "tau_w", "b", "V_r", "V_T", "I_ref", "V_dep", "I_DC"]
VARS = ["V", "w", "g_AMPA_on", "g_AMPA_off", "g_GABA_on", "g_GABA_off", "g_NMDA_on", "g_NMDA_off",
        "g_AMPA", "g_GABA", "g_NMDA", "I_AMPA", "I_GABA", "I_NMDA", "I_syn", "I_inj", "I_tot", "I_EXC", "I_EXC_tot",
        "I_exp", "last_spike"]
VAR_ID = {n: i for i, n in enumerate(VARS)}
SYN_VARS = {"u": 0, "R": 1, "a": 2, "delay_steps": 3}
FUNCTIONS = [
    "abi_version", "create", "destroy", "last_error", "set_neurons", "set_has_outgoing", "set_channels", "set_synapses",
    "add_spike_source", "add_state_monitor", "set_monitor_active", "monitor_samples", "read_monitor",
    "read_monitor_dev", "clear_monitor", "run", "spike_count", "read_spikes", "read_state", "write_state",
    "read_syn_state", "get_counters", "current_step", "comm_unique_id", "comm_init", "profile_kernels", "set_idc", "set_synapses_dev", "set_timed_current", "set_noise",
]


class HhdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hhd error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_neurons", C.c_int32), ("n_global", C.c_int32), ("global_offset", C.c_int32),
        ("dt_ms", C.c_double), ("method", C.c_int32), ("accum", C.c_int32), ("plan", C.c_int32),
        ("refractory_ms", C.c_double), ("block_window_ms", C.c_double), ("last_spike0_ms", C.c_double),
        ("seed", C.c_uint64), ("device", C.c_int32), ("rank", C.c_int32), ("n_ranks", C.c_int32),
        ("record_spikes", C.c_int32), ("instance_size", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("neuron_updates", C.c_int64), ("syn_events", C.c_int64), ("syn_deliveries", C.c_int64),
        ("ext_events", C.c_int64), ("spikes", C.c_int64), ("w_crossings", C.c_int64), ("nonfinite", C.c_int64),
        ("kernel_launches", C.c_int64), ("run_ms_device", C.c_double), ("reserved", C.c_int64 * 6),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


def load_library(path=None):
    """Load libhhd.so.  Fails loudly: the CUDA extension IS the product."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback." % path)
    return C.CDLL(path)


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("expected %d values, got %d" % (n, a.size))
    return a


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Engine:
    """One network on one GPU (or one partition of a network, when n_ranks > 1)."""

    def __init__(self, n_neurons, dt_ms, method="rk4", accum="fp64", plan="auto", refractory_ms=5.0,
                 block_window_ms=5.0, last_spike0_ms=-1e8, seed=0, device=0, rank=0, n_ranks=1, n_global=None,
                 global_offset=0, record_spikes=True, instance_size=0, _lib=None, _prefix="hhd_"):
        self._lib = _lib if _lib is not None else load_library()
        self._prefix = _prefix
        self._h = C.c_void_p()
        self.n_neurons = int(n_neurons)
        self.n_global = int(n_global if n_global is not None else n_neurons)
        self.global_offset = int(global_offset)
        self.dt_ms = float(dt_ms)
        self.nsyn = 0
        self._monitors = {}
        self._mon_active, self._mon_runs, self._spike_epoch = {}, {}, 0
        if self._fn("abi_version", C.c_int)() != ABI_VERSION:
            raise ImportError("libhhd ABI version mismatch")
        cfg = Config(
            abi_version=ABI_VERSION, n_neurons=self.n_neurons, n_global=self.n_global,
            global_offset=self.global_offset, dt_ms=self.dt_ms,
            method=METHODS[method] if isinstance(method, str) else int(method),
            accum=ACCUM[accum] if isinstance(accum, str) else int(accum),
            plan=PLANS[plan] if isinstance(plan, str) else int(plan),
            refractory_ms=refractory_ms, block_window_ms=block_window_ms, last_spike0_ms=last_spike0_ms,
            seed=int(seed) & 0xFFFFFFFFFFFFFFFF, device=device, rank=rank, n_ranks=n_ranks,
            record_spikes=1 if record_spikes else 0, instance_size=int(instance_size))
        rc = self._fn("create")(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._fn("last_error", C.c_char_p)(None)
            self._h = C.c_void_p()
            raise HhdError(rc, (msg or b"").decode())

    # -- plumbing --------------------------------------------------------------------------------------------------
    def _fn(self, name, restype=C.c_int):
        f = getattr(self._lib, self._prefix + name)
        f.restype = restype
        return f

    def _check(self, rc):
        if rc != 0:
            msg = self._fn("last_error", C.c_char_p)(self._h)
            raise HhdError(rc, (msg or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._fn("destroy", None)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- model -----------------------------------------------------------------------------------------------------
    def set_neurons(self, params, V0, w0):
        """params: dict name->array (PARAM_ROWS) or array [12, N] in that row order; units pF nS mV ms pA."""
        n = self.n_neurons
        if isinstance(params, dict):
            rows = [_f64(params[k], n) for k in PARAM_ROWS]
        else:
            params = np.asarray(params, dtype=np.float64)
            rows = [_f64(params[i], n) for i in range(len(PARAM_ROWS))]
        arr = (_dp * len(rows))(*[_ptr(r, _dp) for r in rows])
        V0 = _f64(V0, n)
        w0 = _f64(w0, n)
        self._check(self._fn("set_neurons")(self._h, arr, _ptr(V0, _dp), _ptr(w0, _dp)))

    def set_idc(self, idc):
        a = _f64(idc, self.n_neurons)
        self._check(self._fn("set_idc")(self._h, _ptr(a, _dp)))

    def set_timed_current(self, table, col_of_neuron):
        """I_AC(t): table [n_steps, n_cols] in pA (row s = step s, last row held afterwards), col_of_neuron [N] (-1: none)."""
        table = np.ascontiguousarray(table, dtype=np.float64)
        if table.ndim != 2:
            raise ValueError("table must be [n_steps, n_cols]")
        col = np.ascontiguousarray(col_of_neuron, dtype=np.int32)
        if col.size != self.n_neurons:
            raise ValueError("col_of_neuron must have n_neurons entries")
        self._check(self._fn("set_timed_current")(self._h, C.c_int64(table.shape[0]), C.c_int32(table.shape[1]),
                                                   _ptr(table, _dp), _ptr(col, _ip)))

    def set_noise(self, sigma):
        """Conductance noise: sigma [N] = sqrt(2 * noise_D * nCon) in nS/sqrt(ms) (euler only)."""
        a = _f64(sigma, self.n_neurons)
        self._check(self._fn("set_noise")(self._h, _ptr(a, _dp)))

    def set_has_outgoing(self, has_out_global):
        a = np.ascontiguousarray(has_out_global, dtype=np.uint8)
        if a.size != self.n_global:
            raise ValueError("has_out must have n_global entries")
        self._check(self._fn("set_has_outgoing")(self._h, _ptr(a, _bp)))

    def set_channels(self, tau_on, tau_off, E_rev, gsyn, mg_fac, mg_slope, mg_half):
        a = [_f64(x, 3) for x in (tau_on, tau_off, E_rev, gsyn)]
        self._check(self._fn("set_channels")(self._h, *[_ptr(x, _dp) for x in a], C.c_double(mg_fac),
                                              C.c_double(mg_slope), C.c_double(mg_half)))

    def set_synapses(self, src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms, syn_id_base=0):
        src = np.ascontiguousarray(src, dtype=np.int32)
        n = src.size
        tgt = np.ascontiguousarray(tgt, dtype=np.int32)
        chan = np.ascontiguousarray(chan, dtype=np.uint8)
        d = [_f64(x, n) for x in (g_max, U, tau_rec, tau_fac, p_fail, delay_ms)]
        if tgt.size != n or chan.size != n:
            raise ValueError("synapse arrays differ in length")
        self._check(self._fn("set_synapses")(self._h, C.c_int64(n), _ptr(src, _ip), _ptr(tgt, _ip), _ptr(chan, _bp),
                                              *[_ptr(x, _dp) for x in d], C.c_int64(syn_id_base)))
        self.nsyn = n

    def set_synapses_dev(self, src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms, syn_id_base=0):
        """Same as set_synapses with torch CUDA tensors (int32, int32, uint8, 6 x float64) that already live on the GPU."""
        import torch
        ts = [src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms]
        want = [torch.int32, torch.int32, torch.uint8] + [torch.float64] * 6
        n = src.numel()
        for t, w in zip(ts, want):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == w and t.numel() == n):
                raise ValueError("set_synapses_dev needs contiguous CUDA tensors of dtype %s and equal length" % w)
        self._check(self._fn("set_synapses_dev")(self._h, C.c_int64(n), *[C.c_void_p(t.data_ptr()) for t in ts], C.c_int64(syn_id_base)))
        self.nsyn = n

    def add_spike_source(self, n_sources, spk_src, spk_t_ms, syn_src, syn_tgt, syn_chanmask, syn_gmax, syn_pfail):
        spk_src = np.ascontiguousarray(spk_src, dtype=np.int32)
        spk_t = _f64(spk_t_ms, spk_src.size)
        syn_src = np.ascontiguousarray(syn_src, dtype=np.int32)
        ns = syn_src.size
        syn_tgt = np.ascontiguousarray(syn_tgt, dtype=np.int32)
        mask = np.ascontiguousarray(syn_chanmask, dtype=np.uint8)
        g = _f64(syn_gmax, ns)
        pf = _f64(syn_pfail, ns)
        self._check(self._fn("add_spike_source")(
            self._h, C.c_int32(n_sources), C.c_int64(spk_src.size), _ptr(spk_src, _ip), _ptr(spk_t, _dp),
            C.c_int64(ns), _ptr(syn_src, _ip), _ptr(syn_tgt, _ip), _ptr(mask, _bp), _ptr(g, _dp), _ptr(pf, _dp)))

    # -- monitors --------------------------------------------------------------------------------------------------
    def add_state_monitor(self, var, idx=None, reduce="none", capacity_steps=0):
        var_id = VAR_ID[var] if isinstance(var, str) else int(var)
        mon = C.c_int32(-1)
        if idx is None:
            n_idx, p = self.n_neurons, None
        else:
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            n_idx, p = idx.size, _ptr(idx, _ip)
        red = REDUCE[reduce] if isinstance(reduce, str) else int(reduce)
        self._check(self._fn("add_state_monitor")(self._h, C.c_int32(var_id), C.c_int64(n_idx), p, C.c_int32(red),
                                                   C.c_int64(capacity_steps), C.byref(mon)))
        self._monitors[mon.value] = (n_idx if red == 0 else 1)
        self._mon_active[mon.value] = True
        self._mon_runs[mon.value] = []
        return mon.value

    def set_monitor_active(self, mon_id, active):
        self._check(self._fn("set_monitor_active")(self._h, C.c_int32(mon_id), C.c_int32(1 if active else 0)))
        self._mon_active[mon_id] = bool(active)
        self._spike_epoch += mon_id == -1

    def monitor_steps(self, mon_id):
        """Time step of every sample the monitor holds: a monitor that was switched off and on again (the revision's
        start/stop schedule, Brian's `monitor.active`) has gaps, so this is not first_step + arange(n)."""
        runs = self._mon_runs.get(mon_id, [])
        return np.concatenate([np.arange(a, a + n, dtype=np.int64) for a, n in runs]) if runs else np.zeros(0, dtype=np.int64)

    def set_spike_monitor_active(self, active):
        self.set_monitor_active(-1, active)

    def monitor_samples(self, mon_id):
        n = C.c_int64(0)
        first = C.c_int64(-1)
        self._check(self._fn("monitor_samples")(self._h, C.c_int32(mon_id), C.byref(n), C.byref(first)))
        return n.value, first.value

    def read_monitor(self, mon_id):
        """-> array [n_samples, n_idx] (n_idx = 1 for reduced monitors)."""
        n, _ = self.monitor_samples(mon_id)
        width = self._monitors[mon_id]
        out = np.empty((n, width), dtype=np.float64)
        got = C.c_int64(0)
        self._check(self._fn("read_monitor")(self._h, C.c_int32(mon_id), _ptr(out, _dp), C.c_int64(out.size), C.byref(got)))
        return out[:got.value]

    def clear_monitor(self, mon_id):
        self._check(self._fn("clear_monitor")(self._h, C.c_int32(mon_id)))
        if mon_id == -1:
            self._spike_epoch += 1
        else:
            self._mon_runs[mon_id] = []

    # -- run and results -------------------------------------------------------------------------------------------
    def run(self, n_steps):
        n_steps = int(n_steps)
        k0 = self.current_step() if self._mon_runs else 0
        self._check(self._fn("run")(self._h, C.c_int64(n_steps)))
        for m, runs in self._mon_runs.items():               # which steps each active state monitor sampled
            if self._mon_active.get(m) and n_steps > 0:
                if runs and runs[-1][0] + runs[-1][1] == k0:
                    runs[-1] = (runs[-1][0], runs[-1][1] + n_steps)
                else:
                    runs.append((k0, n_steps))

    def profile_kernels(self, n_steps):
        """Run n_steps with a CUDA event between kernels -> mean microseconds per launch: k_neuron, the exchange (only when
        it is the NCCL all-gather fallback; over peer memory it happens inside the two kernels), k_synapse."""
        us = (C.c_double * 3)()
        self._check(self._fn("profile_kernels")(self._h, C.c_int64(int(n_steps)), us))
        return {"k_neuron": us[0], "exchange": us[1], "k_synapse": us[2]}

    def spike_count(self):
        n = C.c_int64(0)
        self._check(self._fn("spike_count")(self._h, C.byref(n)))
        return n.value

    def spikes(self):
        """-> (neuron ids int32, steps int64), sorted by (step, neuron) like SpikeMonitor.i / .t."""
        n = self.spike_count()
        neuron = np.empty(n, dtype=np.int32)
        step = np.empty(n, dtype=np.int64)
        got = C.c_int64(0)
        self._check(self._fn("read_spikes")(self._h, _ptr(neuron, _ip), _ptr(step, _lp), C.c_int64(n), C.byref(got)))
        return neuron[:got.value], step[:got.value]

    def read_state(self, var):
        out = np.empty(self.n_neurons, dtype=np.float64)
        self._check(self._fn("read_state")(self._h, C.c_int32(VAR_ID[var] if isinstance(var, str) else int(var)), _ptr(out, _dp)))
        return out

    def write_state(self, var, values):
        a = _f64(values, self.n_neurons)
        self._check(self._fn("write_state")(self._h, C.c_int32(VAR_ID[var] if isinstance(var, str) else int(var)), _ptr(a, _dp)))

    def read_syn_state(self, which):
        out = np.empty(self.nsyn, dtype=np.float64)
        self._check(self._fn("read_syn_state")(self._h, C.c_int32(SYN_VARS[which] if isinstance(which, str) else int(which)), _ptr(out, _dp)))
        return out

    def counters(self):
        c = Counters()
        self._check(self._fn("get_counters")(self._h, C.byref(c)))
        return c.as_dict()

    def current_step(self):
        s = C.c_int64(0)
        self._check(self._fn("current_step")(self._h, C.byref(s)))
        return s.value

    # -- multi-GPU -------------------------------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        rc = self._fn("comm_unique_id")(buf)
        if rc != 0:
            raise HhdError(rc, "comm_unique_id failed")
        return bytes(buf)

    def comm_init(self, unique_id):
        buf = (C.c_uint8 * 128)(*unique_id)
        self._check(self._fn("comm_init")(self._h, buf))


# ---- helpers that map the reference's arrays onto the engine --------------------------------------------------------
def gsyn_factors(STypPar):
    """`Gsyn_X = STypPar[0,X]*tau_off*tau_on/(tau_off - tau_on)` (codes/CortexNetwork.py:129-131)."""
    S = np.asarray(STypPar, dtype=np.float64)
    return S[0] * S[2] * S[1] / (S[2] - S[1])


def channels_from_STypPar(STypPar):
    """Channel constants the codes/ façade formats into its equation string (codes/CortexNetwork.py:113-131)."""
    S = np.asarray(STypPar, dtype=np.float64)
    return dict(tau_on=S[1].copy(), tau_off=S[2].copy(), E_rev=S[3].copy(), gsyn=gsyn_factors(S),
                mg_fac=float(S[5, 2]), mg_slope=float(S[6, 2]), mg_half=float(S[7, 2]))


def neuron_rows_from_NeuPar(NeuPar, I_DC):
    """NeuPar rows C,g_L,E_L,delta_T,V_up,tau_w,a,b,V_r,V_T,I_ref,V_dep (codes/CortexNetwork.py:322-333) -> PARAM_ROWS."""
    P = np.asarray(NeuPar, dtype=np.float64)
    rows = [P[0], P[1], P[2], P[3], P[4], P[5], P[7], P[8], P[9], P[10], P[11], np.asarray(I_DC, dtype=np.float64)]
    return np.ascontiguousarray(np.stack(rows))


def load_codes_network(engine, NeuPar, V0, STypPar, SynPar, source_arr, target_arr, I_DC):
    """Feed the arrays `codes/NetworkSetup.NetworkSetup` returns into an engine (what CortexNetwork.__init__ does with
    Brian objects, codes/CortexNetwork.py:308-482).  SynPar rows: type,U,tau_rec,tau_fac,g_max,delay,p_fail."""
    engine.set_neurons(neuron_rows_from_NeuPar(NeuPar, I_DC), np.asarray(V0)[0], np.asarray(V0)[1])
    engine.set_channels(**channels_from_STypPar(STypPar))
    SynPar = np.asarray(SynPar, dtype=np.float64)
    if SynPar.size:
        chan = (SynPar[0].astype(np.int64) - 1).astype(np.uint8)    # 1 AMPA, 2 GABA, 3 NMDA (codes/CortexNetwork.py:411-413)
        engine.set_synapses(source_arr, target_arr, chan, SynPar[4], SynPar[1], SynPar[2], SynPar[3], SynPar[6], SynPar[5])
