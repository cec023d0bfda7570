"""Per-neuron derived quantities of the network generator for MANY neurons at once: the current I_ref that gives an onset
rate of 200 Hz, the onset latencies used to tell IN-L from IN-L-d cells, and the onset/steady rate ratio used to tell IN-CL
from IN-CL-AC cells (codes/NetworkSetup.py:112-149 with FRsimpAdEx :444-553, IntPoints_simpAdEx :559-606, Define_I_ref
:609-618 of the reference).

The reference does this neuron by neuron with scipy: `fmin` (Nelder-Mead, ~60 evaluations) around `quad` (QUADPACK, ~25 Python
callbacks per integral) — 90 ms per neuron, 25 CPU-hours for the 1000-column network of BASELINE config 5.  `NetworkSetup`
keeps that path (bit-exact with the reference, tests/test_network_setup.py).  This module is the builder's fast path
(`sparse_network(..., fast=True)`, used for networks the reference could never build): the same closed-form inter-spike
interval, evaluated for all neurons of a group as array expressions —
  * the three pieces of the interval (w constant / w on the nullcline band / w constant again) by composite Gauss-Legendre
    quadrature (8 panels x 16 nodes) instead of adaptive QUADPACK,
  * the point where the trajectory meets the band by bisection on the convex function the reference hands to `fsolve`
    (only the lower of its two roots is ever used),
  * I_ref by bisection on the monotone onset f-I curve instead of Nelder-Mead on (rate - 200)^2 —
so it is NOT bit-identical: I_ref agrees with the reference's value to within the tolerance `fmin` itself stops at (a few
0.01 pA, tests/test_fast_neuron.py), and a neuron sitting exactly on a re-typing threshold could be typed differently.
"""
import numpy as np

_PANELS, _NODES = 8, 16
_GX, _GW = np.polynomial.legendre.leggauss(_NODES)


def _integrate(f, a, b):
    """Integral of f(V) (V: [n, m] array -> [n, m]) from a[n] to b[n]; zero-width intervals give 0."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    edges = a[:, None] + (b - a)[:, None] * np.linspace(0.0, 1.0, _PANELS + 1)[None, :]
    mid = 0.5 * (edges[:, 1:] + edges[:, :-1])
    half = 0.5 * (edges[:, 1:] - edges[:, :-1])
    V = (mid[:, :, None] + half[:, :, None] * _GX[None, None, :]).reshape(len(a), -1)
    W = (half[:, :, None] * _GW[None, None, :]).reshape(len(a), -1)
    with np.errstate(all="ignore"):
        return np.sum(f(V) * W, axis=1)


def firing_rate(par, I, w0=None, V0=None):
    """Vectorised FRsimpAdEx: par = NeuPar rows [>= 10, n] (C, g_L, E_L, delta_T, V_up, tau_w, a, b, V_r, V_T), I[n] pA.
    w0=None: steady-state rate (w_r = w_end + b); else onset rate from (w0, V0 or V_r).  nan where the neuron does not fire."""
    Cm, gL, EL, sf, Vup, tcw, _, b, Vr, Vth = (np.asarray(par[r], dtype=np.float64) for r in range(10))
    n = len(Cm)
    I = np.broadcast_to(np.asarray(I, dtype=np.float64), (n,)).copy()
    with np.errstate(all="ignore"):
        tau = Cm / gL
        f = tau / tcw
        X_Vth = f * (I + gL * sf - gL * (Vth - EL))
        w_end = -gL * (Vth - EL) + gL * sf + I - X_Vth
        w_r = np.where(b != 0, w_end + b, 0.0) if w0 is None else np.broadcast_to(np.asarray(w0, dtype=np.float64), (n,)).copy()
        V_r = Vr.copy() if V0 is None else np.broadcast_to(np.asarray(V0, dtype=np.float64), (n,)).copy()
        silent = (X_Vth <= 0) | (f >= 1) | (b < 0) | (Cm <= 0) | (gL <= 0) | (tcw <= 0) | (sf <= 0)
        e_r = np.exp((V_r - Vth) / sf)
        X_Vr = f * (I + gL * sf * e_r - gL * (V_r - EL))
        wV_Vr = -gL * (V_r - EL) + gL * sf * e_r + I
        w1, w2 = wV_Vr - X_Vr, wV_Vr + X_Vr

        above = V_r >= Vth
        caseA1 = above & (w_r >= wV_Vr)
        caseA2 = above & ~caseA1
        caseB1 = ~above & (w1 < w_r) & (w_r < w2)
        caseB2 = ~above & ~caseB1
        sign = np.where(caseA1 | (caseB2 & (w_r > w1)), 1.0, -1.0)
        w_ref = -gL * (Vth - EL) + gL * sf + I + sign * X_Vth
        crosses = (caseA1 | caseB2) & (w_r > w_ref)            # the trajectory w = w_r meets the band below V_T

        # lower root of G(V) = (1 + sign f)(I - g_L (V - E_L) + g_L sf exp((V - V_T)/sf)) - w_r: G falls monotonically for V < V_T
        fac = 1.0 + sign * (Cm / (gL * tcw))
        G = lambda V: fac * (I - gL * (V - EL) + gL * sf * np.exp((V - Vth) / sf)) - w_r
        hi = Vth.copy()
        lo = Vth - 1.0
        for _ in range(60):                                    # expand until G(lo) > 0
            need = crosses & (G(lo) <= 0)
            if not need.any():
                break
            lo = np.where(need, Vth - 2.0 * (Vth - lo), lo)
        for _ in range(70):
            mid = 0.5 * (lo + hi)
            pos = G(mid) > 0
            lo = np.where(pos, mid, lo)
            hi = np.where(pos, hi, mid)
        Vlb = 0.5 * (lo + hi)

        m = f * gL - gL
        si, sj, sk = m, (1 - f) * I - m * EL, (1 - f) * gL * sf

        def fixed_w(w):
            return lambda V: 1.0 / ((1 / Cm)[:, None] * (I[:, None] - w[:, None] + (gL * sf)[:, None] * np.exp((V - Vth[:, None]) / sf[:, None])
                                                          - gL[:, None] * (V - EL[:, None])))

        def on_band(V):
            e = np.exp((V - Vth[:, None]) / sf[:, None])
            return 1.0 / ((1 / Cm)[:, None] * (I[:, None] - (si[:, None] * V + sj[:, None] + sk[:, None] * e) + (gL * sf)[:, None] * e
                                               - gL[:, None] * (V - EL[:, None])))

        # piece 1: w = w_r from V_r to the band (or to V_T); piece 2: along the band up to V_T; piece 3: w = w_stop up to V_up
        a1 = V_r
        b1 = np.where(caseA1 | caseB2, np.where(crosses, Vlb, Vth), V_r)
        a2 = np.where(crosses, Vlb, np.where(caseB1, V_r, Vth))
        b2 = np.where(crosses | caseB1, Vth, a2)
        w_stop = np.where(caseA2 | (caseB2 & ~crosses), w_r, w_end)
        V1b = np.where(caseA2, V_r, Vth)
        b3 = np.where(V1b >= Vup, V1b, Vup)
        t = _integrate(fixed_w(w_r), a1, b1) + _integrate(on_band, a2, b2) + _integrate(fixed_w(w_stop), V1b, b3)
        rate = 1000.0 / t
    rate[silent | ~np.isfinite(rate)] = np.nan
    return rate


def reference_current(par, target_hz=200.0):
    """I_ref: the current whose ONSET rate (w = 0, V = V_r) is `target_hz` — what the reference finds with
    fmin(Define_I_ref, I_rheo + 100).  Bisection on the monotone onset f-I curve."""
    gL, EL, sf, Vth = (np.asarray(par[r], dtype=np.float64) for r in (1, 2, 3, 9))
    n = len(gL)
    lo = gL * (Vth - EL) - gL * sf                     # rheobase: rate 0
    hi = lo + 100.0
    rate = lambda I: np.nan_to_num(firing_rate(par, I, 0.0, None), nan=0.0)
    width = 100.0
    for _ in range(40):                                # widen until the rate at `hi` exceeds the target
        low = rate(hi) < target_hz
        if not low.any():
            break
        width *= 2.0
        lo = np.where(low, hi, lo)
        hi = np.where(low, hi + width, hi)
    # Illinois variant of regula falsi on g(I) = rate(I) - target (monotone): ~8 evaluations instead of 48 bisections
    g_lo, g_hi = rate(lo) - target_hz, rate(hi) - target_hz
    side = np.zeros(n, dtype=np.int8)
    for _ in range(40):
        with np.errstate(all="ignore"):
            x = (lo * g_hi - hi * g_lo) / (g_hi - g_lo)
        x = np.where(np.isfinite(x) & (x > lo) & (x < hi), x, 0.5 * (lo + hi))
        g = rate(x) - target_hz
        neg = g < 0
        # the end that is kept twice in a row gets its function value halved
        g_hi = np.where(neg & (side == -1), 0.5 * g_hi, g_hi)
        g_lo = np.where(~neg & (side == 1), 0.5 * g_lo, g_lo)
        lo, g_lo = np.where(neg, x, lo), np.where(neg, g, g_lo)
        hi, g_hi = np.where(neg, hi, x), np.where(neg, g_hi, g)
        side = np.where(neg, -1, 1).astype(np.int8)
        if np.all((hi - lo < 1e-9 * np.maximum(1.0, np.abs(hi))) | (np.abs(g) < 1e-10)):
            break
    return np.where(np.abs(g_lo) < np.abs(g_hi), lo, hi)


def per_neuron(par, latency=True, adaptation=True):
    """-> (I_ref[n], t_lat[n], t_lat_LIF[n], adaptation ratio[n]) as `_per_neuron` of NetworkSetup.py gives them one by one.
    The generator uses the latencies only to re-type IN-L cells and the ratio only to re-type IN-CL cells: pass False for the
    other groups and zeros come back."""
    par = np.asarray(par, dtype=np.float64)
    C, gL, EL, sf, _, _, _, _, _, Vth = (par[r] for r in range(10))
    n = par.shape[1]
    Iref = reference_current(par)
    I = 500.0
    t_lat, t_lif, med = np.zeros(n), np.zeros(n), np.zeros(n)
    with np.errstate(all="ignore"):
        if latency:
            t_lat = _integrate(lambda V: C[:, None] / (I - gL[:, None] * (V - EL[:, None]) + (gL * sf)[:, None] * np.exp((V - Vth[:, None]) / sf[:, None])),
                               EL, Vth / 2)
            t_lif = C * np.log(I / (I + gL * (EL - Vth / 2))) / gL
        if adaptation:
            cur = np.arange(25, 301, 25)
            onset = np.stack([firing_rate(par, float(c), 0.0, EL) for c in cur])
            steady = np.stack([firing_rate(par, float(c), None, None) for c in cur])
            ok = np.nan_to_num(steady, nan=0.0) > 0
            ratio = np.where(ok, onset / steady, np.nan)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                med = np.nanmedian(ratio, axis=0)                   # median over the currents with a steady rate ...
            med[(ok & np.isnan(onset)).any(axis=0)] = np.nan         # ... which np.median turns into nan if an onset rate is missing
            med = np.nan_to_num(med, nan=0.0)
    return Iref, t_lat, t_lif, med
