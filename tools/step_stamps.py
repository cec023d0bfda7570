"""Where a time step of a partitioned run goes: HHD_DEBUG_TIMES=gpurun_out/stamps torchrun ... bench.py ...; python tools/step_stamps.py gpurun_out/stamps.rank*
Stamps (%globaltimer, ns) per step and rank: 0 k_neuron's CTA 0 starts, 1 its last CTA is done, 2 spike counts sent to the peers,
3 k_synapse's CTA 0 starts, 4 CTA 0 has done its share of the local spikes and seen every rank's count, 5 ... its share of the
remote spikes, 6 CTA 0 of k_synapse done."""
import sys

import numpy as np

NAMES = ["k_neuron (first CTA start -> last CTA done)", "counts to peers", "launch edge -> k_synapse starts", "local spikes (CTA 0) + wait for counts",
         "remote spikes (CTA 0)", "rest of CTA 0", "other CTAs' tails + launch edge -> next k_neuron"]
for path in sys.argv[1:]:
    a = np.loadtxt(path, dtype=np.int64)
    t = a[:, 1:8].astype(np.float64)
    order = np.argsort(t[:, 0])
    t = t[order]
    ok = (t > 0).all(axis=1)
    t = t[ok]
    t = t[len(t) // 4:]                                    # steady part of the ring buffer
    d = np.diff(t, axis=1)
    nxt = t[1:, 0] - t[:-1, 6]
    period = np.diff(t[:, 0])
    keep = (period > 0) & (period < 500e3)
    print("%s: %d steps, period mean %.2f us (median %.2f)" % (path, keep.sum(), period[keep].mean() / 1e3, np.median(period[keep]) / 1e3))
    for j in range(6):
        print("   %-48s mean %6.2f  median %6.2f us" % (NAMES[j], d[:-1][keep, j].mean() / 1e3, np.median(d[:-1][keep, j]) / 1e3))
    print("   %-48s mean %6.2f  median %6.2f us" % (NAMES[6], nxt[keep].mean() / 1e3, np.median(nxt[keep]) / 1e3))
