"""Time the device spike-train analyses at the reference's scale (1003 neurons, 30 s, 2 ms bins) next to the pandas loop
of the reference restated on the CPU (oracle, on a subset, extrapolated by the pair count)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import analysis_oracle as ao  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200.codes_revision import AnalysesTools as AT  # noqa: E402

rng = np.random.default_rng(0)
n, t_ms, dt = 1003, 30000.0, 0.05
trains = [np.sort(rng.choice(np.arange(int(t_ms / dt)), size=rng.poisson(3.0 * (0.2 + 2 * rng.random()) * t_ms / 1000), replace=False)) * dt
          for _ in range(n)]
idcs = [q for q in range(n) if len(trains[q]) > 2]
lags = list(range(-10, 11))
AT.get_correlations_from_spike_trains(trains, (1000, t_ms), idcs=idcs[:8], delta_t=2, lags=[0])      # warm-up (context, module load)
t0 = time.perf_counter()
res = AT.get_correlations_from_spike_trains(trains, (1000, t_ms), idcs=idcs, delta_t=2, lags=lags)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
AT.get_ISI_stats_from_spike_trains(trains, neuron_idcs=idcs)
t_isi = time.perf_counter() - t0
sub = idcs[:24]
t0 = time.perf_counter()
ao.correlations_revision(trains, (1000, t_ms), sub, 2, [0])
t_cpu = (time.perf_counter() - t0) * (len(idcs) / len(sub)) ** 2 * len(lags)
print("neurons %d, bins %d, lags %d: device path %.2f s (ISI stats %.3f s); np.corrcoef loop (extrapolated from %d neurons, 1 lag) %.0f s"
      % (len(idcs), int((t_ms - 1000) // 2), len(lags), t_dev, t_isi, len(sub), t_cpu))
print("cross mean at lag 0: %.6f" % res[3][lags.index(0)])
