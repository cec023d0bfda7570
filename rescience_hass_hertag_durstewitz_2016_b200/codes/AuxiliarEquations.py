"""Spike-train statistics of the reference's `codes/AuxiliarEquations.py`, same names and arguments, computed on the
device (libhhd.so, `hhd_an_*`): the binary vectors are bit-packed there and every ordered pair and lag is an AND + POPC
over them; the coefficients are then the reference's own expressions on those counts (analysis.pearson_binary).
Spike times are in ms, as the reference passes them (`spikemonitor.t / ms`).

  set_binary_vector                    codes/AuxiliarEquations.py:537-548
  pop_rate                             :422-443
  Pearson_correlation / Zero_lag_...   :597-620
  Pearson_correlation_group            :622-656   (mean over the ordered pairs i != j, per lag)
  Zero_lag_Pearson_correlation_group   :658-690   (list over the ordered pairs i != j)
"""
import numpy as np
from scipy.ndimage import gaussian_filter1d as gf1d

from .. import analysis as an


def _group_bits(spike_group, t0, t1, dt, device):
    idcs = list(range(len(spike_group)))
    neuron, t = an.spikes_from_trains(spike_group, idcs)
    n_bins = int((t1 - t0) / dt)                                          # len(t_vector), :545
    bits, _, _ = an.binned_trains(neuron, t, idcs, t0, t1, dt, an.BIN_TRUNC, n_bins, device=device)
    return bits, n_bins


def set_binary_vector(spike_times, t0, t1, dt, device=0):
    n_bins = int((t1 - t0) / dt)
    neuron, t = an.spikes_from_trains([spike_times], [0])
    _, bins, _ = an.binned_trains(neuron, t, [0], t0, t1, dt, an.BIN_TRUNC, n_bins, device=device, bins=True)
    return bins[0].astype(np.float64)


def pop_rate(sp_list, t0, t1, dt, sigma=False, device=0):
    Ntbins = int(round((t1 - t0) / dt, 0))
    Nsparrays = len(sp_list)
    t = np.arange(t0, t1, dt)
    idcs = list(range(Nsparrays))
    neuron, ts = an.spikes_from_trains(sp_list, idcs)
    _, _, counts = an.binned_trains(neuron, ts, idcs, t0, t1, dt, an.BIN_FLOORDIV, Ntbins, device=device, counts=True)
    rate_arr = (counts.astype(np.float64) / Nsparrays) * 1000 / dt          # np.mean(sp_array, axis=0)*1000/dt
    if type(sigma) is not bool:
        rate_arr = gf1d(rate_arr, sigma=sigma)
    return t, rate_arr


def _lag_bins(lag0, lag1, t_bin):
    lag_array = np.arange(lag0, lag1, t_bin)
    return lag_array, (lag_array / t_bin).astype(int)


def _pair_matrices(bits, n_bins, lag_bins, device):
    """[n_lags, N, N] of Pearson(x_i, x_j, lag)."""
    absl = sorted(set(abs(int(l)) for l in lag_bins))
    sxy, head, tail = an.lagged_counts(bits, n_bins, absl, device=device)
    pos = {L: q for q, L in enumerate(absl)}
    return np.stack([an.pearson_binary(sxy[pos[abs(int(l))]], head[pos[abs(int(l))]], tail[pos[abs(int(l))]], n_bins, int(l))
                     for l in lag_bins])


def Pearson_correlation(spike_times1, spike_times2, t0, t1, t_bin, lag0, lag1, device=0):
    lag_array, lag_bins = _lag_bins(lag0, lag1, t_bin)
    bits, n_bins = _group_bits([spike_times1, spike_times2], t0, t1, t_bin, device)
    return lag_array, _pair_matrices(bits, n_bins, lag_bins, device)[:, 0, 1]


def Zero_lag_Pearson_correlation(spike_times1, spike_times2, t0, t1, t_bin, device=0):
    bits, n_bins = _group_bits([spike_times1, spike_times2], t0, t1, t_bin, device)
    return _pair_matrices(bits, n_bins, [0], device)[0, 0, 1]


def Pearson_correlation_group(spike_group, t0, t1, t_bin, lag0, lag1, device=0):
    lag_array, lag_bins = _lag_bins(lag0, lag1, t_bin)
    Ngroup = len(spike_group)
    bits, n_bins = _group_bits(spike_group, t0, t1, t_bin, device)
    r = _pair_matrices(bits, n_bins, lag_bins, device)                      # [n_lags, N, N]
    off = ~np.eye(Ngroup, dtype=bool)
    corr_group = np.ascontiguousarray(r[:, off].T)                          # rows k in the reference's (i, j != i) order
    return lag_array, np.mean(corr_group, axis=0)


def Zero_lag_Pearson_correlation_group(spike_group, t0, t1, t_bin, device=0):
    Ngroup = len(spike_group)
    bits, n_bins = _group_bits(spike_group, t0, t1, t_bin, device)
    r = _pair_matrices(bits, n_bins, [0], device)[0]
    return list(r[~np.eye(Ngroup, dtype=bool)])
