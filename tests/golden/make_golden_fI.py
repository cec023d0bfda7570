"""Known answers for the single-neuron integrator, produced by the REFERENCE's own closed-form simpAdEx firing rate
`FRsimpAdEx` (/root/reference/codes/NetworkSetup.py:444-553, imported unmodified; numpy + scipy only).

SURVEY §4 names this function as the analytic known answer for the state update: a neuron with constant current and
no synapses, started at (V0, w0), must fire its first spike after 1000 / FRsimpAdEx(par, I, w0, V0) ms as dt -> 0.

Cases (each = parameter column, I, V0, w0, reference rate):
  * >= 50 cells of the reference's default column (fixture net_n1000_s0 = the reference's own NetworkSetup, seed 0),
    three currents above rheobase each, from rest (V0 = E_L, w0 = 0: the network's initial state, codes/NetworkSetup.py:151-154),
    from the adapted state (V0 = V_r, w0 = the reference's steady-state reset value, i.e. FRsimpAdEx's default arguments)
    and from a partially adapted state;
  * the four hand-picked parameter sets of codes/SingleNeuron.py:190-294 with their currents I1 = 56, I2 = 37 pA and
    multiples.
Cells with V_up < V_T are left out: the closed form integrates up to V_T before it looks at V_up (NetworkSetup.py:536-542),
so it does not describe them.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_fI.py
"""
import os
import sys
import warnings

import numpy as np

REF = "/root/reference/codes"
HERE = os.path.dirname(os.path.abspath(__file__))

SINGLE_NEURON_SETS = {   # codes/SingleNeuron.py:190-294 (C, g_L, E_L, delta_T, V_up, tau_w, a, b, V_r, V_T, I_ref, V_dep)
    "params1": [242.635, 7.369, -80.251, 24.709, -42.869, 102.958, 0, 6.039, -67.339, -48.045, 1212, -67.339],
    "params4": [242.635, 7.369, -80.251, 24.709, -42.869, 102.958, 0, 6.039, -72.722, -48.045, 1212, -67.339],
    "params2": [247.943, 7.456, -80.634, 23.163, -48.416, 83.243, 0, 7.526, -72.722, -52.574, 1212, -72.722],
    "params3": [247.943, 7.456, -80.634, 23.163, -48.416, 83.243, 0, 7.526, -67.339, -52.574, 1212, -72.722],
}


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, REF)
    from NetworkSetup import FRsimpAdEx   # the reference's own function
    g = np.load(os.path.join(HERE, "net_n1000_s0.npz"), allow_pickle=True)
    P = g["NeuPar"]
    rng = np.random.RandomState(2016)
    ok = np.nonzero(P[4] >= P[9])[0]                      # V_up >= V_T
    cells = np.sort(rng.choice(ok, 64, replace=False))
    pars, cur, V0s, w0s, rates, kinds, origin = [], [], [], [], [], [], []

    def steady_w(p, I):
        # the value FRsimpAdEx itself uses as w_r when no w0 is given (NetworkSetup.py:449-458)
        Cm, gL, EL, sf, Vup, tcw, a, b, Vr, Vth = p[:10]
        f = (Cm / gL) / tcw
        X_Vth = f * (I + gL * sf - gL * (Vth - EL))
        return -gL * (Vth - EL) + gL * sf + I - X_Vth + b

    def add(p, I, V0, w0, kind, who):
        fr = FRsimpAdEx(p, I, w0, V0)
        if fr is None or not np.isfinite(fr) or fr <= 0.5 or fr > 150:    # ISI between the refractory period and 2 s
            return
        pars.append(np.asarray(p, dtype=np.float64)); cur.append(I); V0s.append(V0); w0s.append(w0)
        rates.append(float(fr)); kinds.append(kind); origin.append(who)

    def cases(p, I, who):
        if I >= p[10]:                                    # the I_tot >= I_ref branch is not part of the closed form
            return
        add(p, I, p[2], 0.0, 0, who)                      # from rest
        ws = steady_w(p, I)
        chk = FRsimpAdEx(p, I)                            # reference default = adapted state
        fr = FRsimpAdEx(p, I, ws, p[8])
        assert chk is None or abs(chk - fr) <= 1e-12 * abs(fr), (chk, fr)
        add(p, I, p[8], ws, 1, who)                       # adapted (steady state)
        add(p, I, p[8], 0.5 * ws, 2, who)                 # partially adapted

    for c in cells:
        p = P[:, c]
        I_SN = p[1] * (p[9] - p[2] - p[3])                # rheobase (codes/CortexNetwork.py:339)
        for fac in (1.3, 2.0, 3.5):
            cases(p, I_SN * fac if I_SN > 0 else 50.0 * fac, int(c))
    for k, (name, p) in enumerate(SINGLE_NEURON_SETS.items()):
        p = np.asarray(p, dtype=np.float64)
        for I in (56.0, 37.0, 112.0, 74.0, 200.0):
            cases(p, I, -(k + 1))
    out = dict(params=np.array(pars).T.copy(), I=np.array(cur), V0=np.array(V0s), w0=np.array(w0s),
               rate_hz=np.array(rates), kind=np.array(kinds, dtype=np.int32), origin=np.array(origin, dtype=np.int32),
               STypPar=g["STypPar"])
    np.savez_compressed(os.path.join(HERE, "fI_golden.npz"), **out)
    print("cases:", len(cur), "cells:", len(set(origin)), "rates %.2f..%.2f Hz" % (min(rates), max(rates)))


if __name__ == "__main__":
    main()
