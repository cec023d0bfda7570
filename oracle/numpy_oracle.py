"""numpy_oracle.py — TEST INFRASTRUCTURE (never imported by the product; see the header of hhd_oracle.c).

A second, independent CPU restatement of what Brian 2 executes per `dt` for the reference's equation strings, written the
way Brian's *numpy* code-generation target works: every code object is one vectorised numpy statement over all neurons /
all due synapses, **in SI units** (volt, ampere, siemens, farad, second — what Brian holds internally), with numpy's own
`exp`.  It shares nothing with the CUDA engine or the C oracle (no hhd_math.h, other units, other data structures), so an
error in the shared `hhd_exp` / `hhd_div`, in a unit conversion, or in the C oracle's reading of the schedule shows up as
a disagreement between the two oracles (tests/test_numpy_oracle.py).  It also is the "NumPy oracle" CPU baseline of
BASELINE.md §3.2 (structurally what Brian's numpy target does), timed by bench.py.

Follows, statement by statement:
  equations / subexpressions   /root/reference/codes/CortexNetwork.py:233-302   (noise-free branch)
  threshold / reset / event    :304-311    (`V > V_up`, `V = V_r; w += b`, `w_crossing`, `w = w_V - D0/tau_w`)
  synapse model, 13 pathways   :343-387, delays and orders :463-482, initial u = U, R = 1 :421-426
  Gsyn                         :129-131, 415-419
  schedule                     Brian 2: start -> groups -> thresholds -> synapses -> resets -> after_resets
  integrators                  Brian 2 ExplicitStateUpdater: euler, rk2 (midpoint), rk4
The only thing borrowed from the other side is the definition of the release-failure draw: Brian's sequential `rand()`
(p0) cannot be reproduced by anybody, so both oracles and the engine key a Philox4x32-10 draw by (seed, synapse, step);
the generator is re-implemented here in numpy from the published algorithm (Salmon et al., SC'11).
Increments are applied as Brian does, `g_post += inc` synapse by synapse in queue order (np.add.at) — NOT summed first as
the C oracle and the engine do; the two differ at rounding level when a target gets several increments in one step.
"""
import numpy as np

ms, mV, pA, nS, pF = 1e-3, 1e-3, 1e-12, 1e-9, 1e-12

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO32 = np.uint64(0xFFFFFFFF)


def philox_u01(seed, index, step, stream=0):
    """Philox4x32-10, counter = (index lo, index hi, step lo, step hi[23:0] | stream << 24), key = seed -> uniform in
    [0, 1) with 53 bits (27 from word 0, 26 from word 1).  index: uint64 array."""
    index = np.asarray(index, dtype=np.uint64)
    step = np.uint64(step)
    c0 = index & _LO32
    c1 = index >> np.uint64(32)
    c2 = np.full_like(c0, step & _LO32)
    c3 = np.full_like(c0, ((step >> np.uint64(32)) & np.uint64(0x00FFFFFF)) | (np.uint64(stream) << np.uint64(24)))
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & _LO32, n2, p0 & _LO32
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    bits = ((c0 >> np.uint64(5)) << np.uint64(26)) | (c1 >> np.uint64(6))
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


def timestep(t, dt):
    """Brian 2's `timestep(t, dt)` = int((t + 1e-3 dt) / dt)."""
    return int((t + 1e-3 * dt) / dt)


class NumpyOracle:
    """Same call surface as the ctypes `Engine` for what the parity tests and bench.py use (units at the boundary: ms, mV, pA,
    nS, pF like the reference's arrays; converted to SI here the way codes/CortexNetwork.py:322-336, 401-431 does)."""

    def __init__(self, n_neurons, dt_ms, method="rk4", refractory_ms=5.0, block_window_ms=5.0, last_spike0_ms=-1e8, seed=0,
                 record_spikes=True, **_ignored):
        self.N = int(n_neurons)
        self.dt_ms = float(dt_ms)
        self.dt = self.dt_ms * ms
        self.method = method
        self.refr_steps = timestep(refractory_ms * ms, self.dt)
        self.window = block_window_ms * ms
        self.seed = int(seed)
        self.k = 0
        self.record = record_spikes
        self.last_spike = np.full(self.N, last_spike0_ms * ms)     # `last_spike` of the NEURON group, written by pathway p5
        self.lastspike_step = np.full(self.N, -(1 << 40), dtype=np.int64)   # Brian's own `lastspike` (refractoriness)
        self.spk_i, self.spk_k = [], []
        self.nsyn = 0
        self.n_spikes = self.n_events = self.n_wcross = 0
        self.lfp = None

    # -- model ---------------------------------------------------------------------------------------------------------
    def set_neurons(self, rows, V0, w0):
        r = np.asarray(rows, dtype=np.float64)
        (self.C, self.g_L, self.E_L, self.delta_T, self.V_up, self.tau_w, self.b, self.V_r, self.V_T, self.I_ref, self.V_dep,
         self.I_DC) = (r[0] * pF, r[1] * nS, r[2] * mV, r[3] * mV, r[4] * mV, r[5] * ms, r[6] * pA, r[7] * mV, r[8] * mV,
                       r[9] * pA, r[10] * mV, r[11] * pA)
        self.V = np.asarray(V0, dtype=np.float64) * mV
        self.w = np.asarray(w0, dtype=np.float64) * pA
        self.g = np.zeros((6, self.N))       # AMPA_on, AMPA_off, GABA_on, GABA_off, NMDA_on, NMDA_off (siemens)

    def set_idc(self, idc):
        self.I_DC = np.asarray(idc, dtype=np.float64) * pA

    def set_channels(self, tau_on, tau_off, E_rev, gsyn, mg_fac, mg_slope, mg_half):
        self.tau_on = np.asarray(tau_on, dtype=np.float64) * ms
        self.tau_off = np.asarray(tau_off, dtype=np.float64) * ms
        self.E = np.asarray(E_rev, dtype=np.float64) * mV
        self.gsyn = np.asarray(gsyn, dtype=np.float64)             # dimensionless in Brian (`Gsyn: 1`)
        self.mg = (float(mg_fac), float(mg_slope), float(mg_half) * mV)

    def set_synapses(self, src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms, syn_id_base=0):
        src = np.asarray(src, dtype=np.int64)
        order = np.argsort(src, kind="stable")                     # a spiking neuron's synapses in creation order
        self.row_syn = order
        self.row_ptr = np.searchsorted(src[order], np.arange(self.N + 1))
        self.tgt = np.asarray(tgt, dtype=np.int64)
        self.chan = np.asarray(chan, dtype=np.int64)
        self.g_max = np.asarray(g_max, dtype=np.float64) * nS
        self.U = np.asarray(U, dtype=np.float64)
        self.tau_rec = np.asarray(tau_rec, dtype=np.float64) * ms
        self.tau_fac = np.asarray(tau_fac, dtype=np.float64) * ms
        self.p_fail = np.asarray(p_fail, dtype=np.float64)
        d = np.asarray(delay_ms, dtype=np.float64)
        self.delay_steps = np.floor(d / self.dt_ms + 0.5).astype(np.int64)      # Brian's C++ spike queue: int(delay/dt + 0.5)
        self.nsyn = len(src)
        self.u = self.U.copy()                                      # :421-426
        self.R = np.ones(self.nsyn)
        self.a_syn = np.zeros(self.nsyn)
        self.failure = np.zeros(self.nsyn)
        self.has_out = np.diff(self.row_ptr) > 0
        self.syn_id_base = int(syn_id_base)
        self.maxd = int(self.delay_steps.max()) if self.nsyn else 0
        self.queue = [[] for _ in range(self.maxd + 1)]

    def add_state_monitor(self, var, idx=None, reduce="none", capacity_steps=0):
        if var != "I_tot" or reduce != "sum" or idx is not None:
            raise NotImplementedError("the NumPy oracle records the summed I_tot only")
        self.lfp = []
        return 0

    def read_monitor(self, mon_id):
        return np.asarray(self.lfp, dtype=np.float64)[:, None] / pA

    # -- the equations (codes/CortexNetwork.py:233-272) ---------------------------------------------------------------------
    def _sub(self, V, g):
        g_AMPA, g_GABA, g_NMDA = g[1] - g[0], g[3] - g[2], g[5] - g[4]
        I_AMPA = g_AMPA * (self.E[0] - V)
        Mg = 1.0 / (1.0 + self.mg[0] * np.exp(self.mg[1] * (self.mg[2] - V) / mV))
        I_GABA = g_GABA * (self.E[1] - V)
        I_NMDA = g_NMDA * (self.E[2] - V) * Mg
        I_syn = I_AMPA + I_NMDA + I_GABA
        I_tot = I_syn + self.I_DC
        ex = np.exp((V - self.V_T) / self.delta_T)
        I_exp = self.g_L * self.delta_T * ex
        w_V = I_tot + I_exp - self.g_L * (V - self.E_L)
        D0 = (self.C / self.g_L) * w_V
        dD0 = self.C * (ex - 1.0)
        return I_tot, ex, I_exp, w_V, D0, dD0

    def _f(self, V, w, g, t):
        with np.errstate(over="ignore", invalid="ignore"):
            I_tot, ex, I_exp, w_V, D0, dD0 = self._sub(V, g)
            blk = (I_tot >= self.I_ref).astype(np.float64) * ((t - self.last_spike) < self.window).astype(np.float64)
            dV = blk * (-self.g_L / self.C) * (V - self.V_dep) + (1.0 - blk) * (I_tot + I_exp - self.g_L * (V - self.E_L) - w) / self.C
            gate = ((w > w_V - D0 / self.tau_w) & (w < w_V + D0 / self.tau_w) & (V <= self.V_T) & (I_tot < self.I_ref)).astype(np.float64)
            dw = gate * -(self.g_L * (1.0 - ex) + dD0 / self.tau_w) * dV
        dg = np.empty_like(g)
        dg[0::2] = -(1.0 / self.tau_on)[:, None] * g[0::2]
        dg[1::2] = -(1.0 / self.tau_off)[:, None] * g[1::2]
        return dV, dw, dg

    def _integrate(self, t):
        dt, V, w, g = self.dt, self.V, self.w, self.g
        k1 = [dt * d for d in self._f(V, w, g, t)]
        if self.method == "euler":
            return V + k1[0], w + k1[1], g + k1[2]
        k2 = [dt * d for d in self._f(V + k1[0] / 2, w + k1[1] / 2, g + k1[2] / 2, t + dt / 2)]
        if self.method == "rk2":
            return V + k2[0], w + k2[1], g + k2[2]
        if self.method != "rk4":
            raise NotImplementedError(self.method)
        k3 = [dt * d for d in self._f(V + k2[0] / 2, w + k2[1] / 2, g + k2[2] / 2, t + dt / 2)]
        k4 = [dt * d for d in self._f(V + k3[0], w + k3[1], g + k3[2], t + dt)]
        return tuple(x + (a + 2 * b + 2 * c + d) / 6 for x, a, b, c, d in zip((V, w, g), k1, k2, k3, k4))

    # -- one time step in Brian's slot order ------------------------------------------------------------------------------
    def step(self):
        k = self.k
        t = k * self.dt
        if self.lfp is not None:                                           # start: StateMonitor samples the pre-update state
            self.lfp.append(float(np.sum(self._sub(self.V, self.g)[0])))
        not_refractory = (k - self.lastspike_step) >= self.refr_steps      # groups: stateupdater sets not_refractory, integrates
        self.V, self.w, self.g = self._integrate(t)
        with np.errstate(over="ignore", invalid="ignore"):                 # thresholds
            spikes = np.nonzero((self.V > self.V_up) & not_refractory)[0]
            _, _, _, w_V, D0, _ = self._sub(self.V, self.g)
            wcross = np.nonzero((self.w > w_V - D0 / self.tau_w) & (self.w < w_V + D0 / self.tau_w) & (self.V <= self.V_T))[0]
        self.lastspike_step[spikes] = k
        if self.record and len(spikes):
            self.spk_i.append(spikes.copy())
            self.spk_k.append(np.full(len(spikes), k, dtype=np.int64))
        self.n_spikes += len(spikes)
        self.n_wcross += len(wcross)
        if self.nsyn:                                                      # synapses
            for j in spikes:
                s = self.row_syn[self.row_ptr[j]:self.row_ptr[j + 1]]
                if len(s) == 0:
                    continue
                self.failure[s] = (philox_u01(self.seed, (s + self.syn_id_base).astype(np.uint64), k) < self.p_fail[s]).astype(np.float64)  # p0
                dts = t - self.last_spike[j]
                u1 = self.U[s] + self.u[s] * (1 - self.U[s]) * np.exp(-dts / self.tau_fac[s])          # p1
                R1 = 1 + (self.R[s] - self.u[s] * self.R[s] - 1) * np.exp(-dts / self.tau_rec[s])      # p2 (old u)
                self.u[s] = u1                                                                          # p3
                self.R[s] = R1                                                                          # p4
                self.last_spike[j] = t                                                                  # p5
                self.a_syn[s] = self.u[s] * self.R[s]                                                   # p6
                d = self.delay_steps[s]
                for dd in np.unique(d):                                    # push into the queue slot k + delay_steps
                    self.queue[(k + dd) % (self.maxd + 1)].append(s[d == dd])
                self.n_events += len(s)
            slot = self.queue[k % (self.maxd + 1)]
            if slot:                                                       # p7a..p9b, values read NOW
                s = np.concatenate(slot)
                slot.clear()
                inc = self.g_max[s] * self.a_syn[s] * (1 - self.failure[s]) * self.gsyn[self.chan[s]]
                np.add.at(self.g, (2 * self.chan[s], self.tgt[s]), inc)
                np.add.at(self.g, (2 * self.chan[s] + 1, self.tgt[s]), inc)
        if len(spikes):                                                    # resets
            self.V[spikes] = self.V_r[spikes]
            self.w[spikes] = self.w[spikes] + self.b[spikes]
        if len(wcross):                                                    # after_resets: w = w_V - D0/tau_w on the current state
            _, _, _, w_V, D0, _ = self._sub(self.V, self.g)
            self.w[wcross] = (w_V - D0 / self.tau_w)[wcross]
        self.k += 1

    def run(self, n_steps):
        for _ in range(int(n_steps)):
            self.step()

    # -- results ---------------------------------------------------------------------------------------------------------
    def spikes(self):
        if not self.spk_i:
            return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64)
        return np.concatenate(self.spk_i).astype(np.int32), np.concatenate(self.spk_k)

    def read_state(self, var):
        names = ["g_AMPA_on", "g_AMPA_off", "g_GABA_on", "g_GABA_off", "g_NMDA_on", "g_NMDA_off"]
        if var == "V":
            return self.V / mV
        if var == "w":
            return self.w / pA
        if var == "last_spike":
            return self.last_spike / ms
        return self.g[names.index(var)] / nS

    def read_syn_state(self, which):
        if which == "a":       # the amplitude of the next delivery in nS (0 before the first presynaptic spike), as include/hhd.h defines it
            return self.g_max * self.a_syn * (1 - self.failure) * self.gsyn[self.chan] / nS
        return {"u": self.u, "R": self.R, "delay_steps": self.delay_steps.astype(np.float64)}[which].copy()

    def counters(self):
        return dict(steps=self.k, neuron_updates=self.k * self.N, syn_events=self.n_events, spikes=self.n_spikes,
                    w_crossings=self.n_wcross, nonfinite=int((~np.isfinite(self.V)).sum()))

    def close(self):
        pass
