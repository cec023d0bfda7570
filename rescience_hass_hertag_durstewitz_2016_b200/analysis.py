"""Spike-train analyses: ctypes binding of the `hhd_an_*` entry points of libhhd.so (include/hhd.h) and the reference's
formulas evaluated on the integer counts they return.

The device computes what is quadratic in neurons / linear in spikes — binary trains, S_xy(i, j, lag) for every ordered
pair, window sums, per-neuron ISI moments; the coefficients follow here, operand by operand as the reference writes them
(codes/AuxiliarEquations.py:550-595; pandas' Series.corr of codes_revision/AnalysesTools.py:231-240 as the Pearson
formula on the same counts).  No CPU fallback: without the library or a device every function raises.
"""
import ctypes as C

import numpy as np

from .engine import HhdError, load_library

BIN_TRUNC, BIN_FLOORDIV = 0, 1
FUNCTIONS = ["an_words", "an_binned_trains", "an_lagged_counts", "an_isi_stats"]   # the hhd_an_* entry points bound here
_lib = None


def _library():
    global _lib
    if _lib is None:
        lib = load_library()
        ip, lp, dp = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        up, bp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
        lib.hhd_an_words.argtypes = [C.c_int64]
        lib.hhd_an_words.restype = C.c_int
        lib.hhd_an_binned_trains.argtypes = [C.c_int32, C.c_int64, ip, dp, C.c_int32, ip, C.c_double, C.c_double,
                                             C.c_double, C.c_int32, C.c_int64, up, bp, ip]
        lib.hhd_an_lagged_counts.argtypes = [C.c_int32, C.c_int32, C.c_int64, up, C.c_int32, ip, ip, ip, ip]
        lib.hhd_an_isi_stats.argtypes = [C.c_int32, C.c_int64, ip, dp, C.c_int32, ip, C.c_int32, C.c_double,
                                         C.c_double, dp, dp, ip]
        for f in (lib.hhd_an_binned_trains, lib.hhd_an_lagged_counts, lib.hhd_an_isi_stats):
            f.restype = C.c_int
        _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise HhdError(rc, "%s failed (no CUDA device, or bad arguments); there is no CPU fallback" % what)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None and a.size else C.cast(None, C.POINTER(t))


def spikes_from_trains(trains, idcs):
    """(neuron, t_ms) arrays in time order from per-neuron arrays of spike times in ms (`SpikeMonitor.spike_trains()` values
    / ms).  Row q of every result refers to trains[idcs[q]]; the floats are passed on untouched."""
    neuron, t = [], []
    for q, i in enumerate(idcs):
        tr = np.asarray(trains[i], dtype=np.float64).ravel()
        neuron.append(np.full(tr.size, q, np.int32))
        t.append(tr)
    neuron = np.concatenate(neuron) if neuron else np.zeros(0, np.int32)
    t = np.concatenate(t) if t else np.zeros(0)
    order = np.lexsort((neuron, t))
    return np.ascontiguousarray(neuron[order]), np.ascontiguousarray(t[order])


def binned_trains(neuron, t_ms, sel, t0, t1, bin_ms, rule, n_bins, device=0, bins=False, counts=False):
    """Bit-packed binary trains [len(sel), words] (+ the 0/1 matrix, + neurons active per bin)."""
    lib = _library()
    neuron = np.ascontiguousarray(neuron, np.int32)
    t_ms = np.ascontiguousarray(t_ms, np.float64)
    sel = np.ascontiguousarray(sel, np.int32)
    n_bins = int(n_bins)
    words = lib.hhd_an_words(n_bins)
    bits = np.zeros((sel.size, words), np.uint64)
    out_bins = np.zeros((sel.size, n_bins), np.uint8) if bins else None
    out_counts = np.zeros(n_bins, np.int32) if counts else None
    _check(lib.hhd_an_binned_trains(device, neuron.size, _p(neuron, C.c_int32), _p(t_ms, C.c_double), sel.size,
                                    _p(sel, C.c_int32), float(t0), float(t1), float(bin_ms), int(rule), n_bins,
                                    _p(bits, C.c_uint64), _p(out_bins, C.c_uint8) if bins else C.cast(None, C.POINTER(C.c_uint8)),
                                    _p(out_counts, C.c_int32) if counts else C.cast(None, C.POINTER(C.c_int32))),
           "hhd_an_binned_trains")
    return bits, out_bins, out_counts


def pack_bits(bin_arr):
    """Bit-packed rows (hhd_an layout: bit b of word b // 64, one spare zero word) of a 0/1 matrix [n, n_bins]."""
    bin_arr = np.asarray(bin_arr)
    n, n_bins = bin_arr.shape
    words = (n_bins + 63) // 64 + 1
    padded = np.zeros((n, words * 64), np.uint8)
    padded[:, :n_bins] = bin_arr != 0
    return np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little").view(np.uint64))


def lagged_counts(bits, n_bins, lags, device=0):
    """sxy[l, i, j] = sum_{t < n_bins - L} x_i[t] x_j[t + L], head[l, i] = sum_{t < n_bins - L} x_i[t], tail[l, i] =
    sum_{t >= L} x_i[t] for the non-negative lags L (in bins)."""
    lib = _library()
    bits = np.ascontiguousarray(bits, np.uint64)
    n = bits.shape[0]
    lags = np.ascontiguousarray(lags, np.int32)
    sxy = np.zeros((lags.size, n, n), np.int32)
    head = np.zeros((lags.size, n), np.int32)
    tail = np.zeros((lags.size, n), np.int32)
    _check(lib.hhd_an_lagged_counts(device, n, int(n_bins), _p(bits, C.c_uint64), lags.size, _p(lags, C.c_int32),
                                    _p(sxy, C.c_int32), _p(head, C.c_int32), _p(tail, C.c_int32)), "hhd_an_lagged_counts")
    return sxy, head, tail


def isi_stats(neuron, t_ms, sel, tlim=None, device=0):
    """Per selected neuron: ISI mean [ms], CV (np.std / np.mean) and number of intervals."""
    lib = _library()
    neuron = np.ascontiguousarray(neuron, np.int32)
    t_ms = np.ascontiguousarray(t_ms, np.float64)
    sel = np.ascontiguousarray(sel, np.int32)
    mean = np.zeros(sel.size)
    cv = np.zeros(sel.size)
    count = np.zeros(sel.size, np.int32)
    t0, t1 = (0.0, 0.0) if tlim is None else tlim
    _check(lib.hhd_an_isi_stats(device, neuron.size, _p(neuron, C.c_int32), _p(t_ms, C.c_double), sel.size,
                                _p(sel, C.c_int32), 0 if tlim is None else 1, float(t0), float(t1), _p(mean, C.c_double),
                                _p(cv, C.c_double), _p(count, C.c_int32)), "hhd_an_isi_stats")
    return mean, cv, count


# ---- the reference's formulas on the counts ---------------------------------------------------------------------------
def pearson_binary(sxy, head, tail, n_bins, lag):
    """codes/AuxiliarEquations.py:571-595 `Pearson(t_vec1, t_vec2, lag)` for every ordered pair at once: entry [i, j] is
    Pearson(x_i, x_j, lag).  sxy/head/tail are the slices of `lagged_counts` for |lag|."""
    L = abs(int(lag))
    n = n_bins - L                                    # len(vector1)
    if lag > 0:                                       # vector1 = t_vec1[:-lag], vector2 = t_vec2[lag:]
        s1, s2, joint = head[:, None], tail[None, :], sxy
    elif lag < 0:                                     # vector1 = t_vec1[-lag:], vector2 = t_vec2[:lag]
        s1, s2, joint = tail[:, None], head[None, :], sxy.T
    else:
        s1, s2, joint = head[:, None], head[None, :], sxy
    with np.errstate(divide="ignore", invalid="ignore"):
        px = s1.astype(np.float64) / n                # np.sum(t_vector) / len(t_vector)
        py = s2.astype(np.float64) / n
        pxy = joint.astype(np.float64) / n            # (vector1 @ vector2) / len(vector1)
        r = (pxy - px * py) / np.sqrt(px * (1 - px) * py * (1 - py))
    return np.where((px == 0) | (py == 0), 0.0, r)


def pearson_shifted(sxy, head, tail, n_bins, lag):
    """Pearson coefficient of x_i and x_j.shift(lag) over the rows where both exist — what
    `df[target].corr(df[c].shift(periods=lag))` computes (codes_revision/AnalysesTools.py:231-240) — for every ordered pair:
    entry [i, j] is target i, column j.  NaN where a window has no variance, as np.corrcoef gives."""
    L = abs(int(lag))
    n = float(n_bins - L)
    if lag > 0:                                       # pairs (x_i[t + L], x_j[t])
        sx, sy, joint = tail[:, None], head[None, :], sxy.T
    elif lag < 0:                                     # pairs (x_i[t], x_j[t + L])
        sx, sy, joint = head[:, None], tail[None, :], sxy
    else:
        sx, sy, joint = head[:, None], head[None, :], sxy
    sx = sx.astype(np.float64)
    sy = sy.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        num = n * joint.astype(np.float64) - sx * sy
        den = np.sqrt((n * sx - sx * sx) * (n * sy - sy * sy))
        return num / den
