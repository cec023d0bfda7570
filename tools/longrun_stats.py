"""Long-run statistics of the reference column on the GPU vs the paper's numbers (content.tex:239, 318)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import default_idc, golden_groups, load_golden  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine, load_codes_network  # noqa: E402


def stats(i, s, n, dt, t_total_ms, groups):
    t = s * dt
    keep = t >= 1000.0
    i, t = i[keep], t[keep]
    order = np.argsort(i, kind="stable")
    i, t = i[order], t[order]
    b = np.searchsorted(i, np.arange(n + 1))
    means, cvs = [], []
    counts = np.diff(b)
    for q in range(n):
        tr = t[b[q]:b[q + 1]]
        if len(tr) > 20:
            isi = np.diff(tr)
            means.append(isi.mean())
            cvs.append(isi.std() / isi.mean())
    pc = np.concatenate([groups[0][0], groups[7][0]])
    dur_s = (t_total_ms - 1000.0) / 1000.0
    thr = 10.0 * dur_s / 30.0          # > 10 spikes per 30 s
    return dict(mean_isi=float(np.mean(means)), sd_isi=float(np.std(means)), cv=float(np.mean(cvs)), sd_cv=float(np.std(cvs)),
                n_trains=len(means), frac_spiking=float((counts > thr).mean()), frac_spiking_pc=float((counts[pc] > thr).mean()),
                rate=float(len(i) / n / dur_s))


if __name__ == "__main__":
    method = sys.argv[1] if len(sys.argv) > 1 else "rk4"
    bio_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 31000.0
    g = load_golden("net_n1000_s0_full")
    n = g["NeuPar"].shape[1]
    for seed in (0, 1, 2):
        e = Engine(n, 0.05, method=method, seed=seed)
        load_codes_network(e, g["NeuPar"], g["V0"], g["STypPar"], g["SynPar"], g["source_arr"], g["target_arr"], default_idc(g))
        t0 = time.time()
        e.run(int(bio_ms / 0.05))
        el = time.time() - t0
        i, s = e.spikes()
        st = stats(i, s, n, 0.05, bio_ms, golden_groups(g))
        st.update(seed=seed, method=method, wall_s=round(el, 2), nonfinite=e.counters()["nonfinite"])
        print(st, flush=True)
        e.close()
