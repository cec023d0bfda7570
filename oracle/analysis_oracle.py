"""CPU restatement of the reference's spike-train analyses — TEST INFRASTRUCTURE ONLY (tests/ may import it; the product
computes these on the device through libhhd.so and never calls this file).  Plain loops over numpy vectors, each function
citing the reference lines it follows.  Pinned against tests/golden/analysis_golden.npz, which the reference's own functions
produced (tests/golden/make_golden_analysis.py).
"""
import numpy as np
from scipy.ndimage import gaussian_filter1d


def set_binary_vector(spike_times, t0, t1, dt):
    """codes/AuxiliarEquations.py:537-548"""
    s = np.asarray(spike_times, dtype=np.float64)
    s = s[(s >= t0) & (s < t1)] - t0
    v = np.zeros(int((t1 - t0) / dt))
    v[(s / dt).astype(int)] = 1
    return v


def pearson(v1, v2, lag):
    """codes/AuxiliarEquations.py:550-595 (p_vector, p_joint_vector, Pearson)"""
    if lag > 0:
        a, b = v1[:-lag], v2[lag:]
    elif lag < 0:
        a, b = v1[-lag:], v2[:lag]
    else:
        a, b = v1, v2
    px, py = np.sum(a) / len(a), np.sum(b) / len(b)
    pxy = (a @ b) / len(a)
    if px == 0 or py == 0:
        return 0
    with np.errstate(divide="ignore", invalid="ignore"):
        return (pxy - px * py) / np.sqrt(px * (1 - px) * py * (1 - py))


def pearson_group(group, t0, t1, t_bin, lag0, lag1):
    """codes/AuxiliarEquations.py:622-656: mean over the ordered pairs i != j, per lag"""
    lag_array = np.arange(lag0, lag1, t_bin)
    lag_bins = (lag_array / t_bin).astype(int)
    vec = [set_binary_vector(s, t0, t1, t_bin) for s in group]
    rows = []
    for i in range(len(group)):
        for j in range(len(group)):
            if i != j:
                rows.append([pearson(vec[i], vec[j], l) for l in lag_bins])
    return lag_array, np.mean(np.asarray(rows, dtype=np.float64), axis=0)


def zero_lag_group(group, t0, t1, t_bin):
    """codes/AuxiliarEquations.py:658-690"""
    vec = [set_binary_vector(s, t0, t1, t_bin) for s in group]
    return [pearson(vec[i], vec[j], 0) for i in range(len(group)) for j in range(len(group)) if i != j]


def pop_rate(sp_list, t0, t1, dt, sigma=False):
    """codes/AuxiliarEquations.py:422-443"""
    n_bins = int(round((t1 - t0) / dt, 0))
    arr = np.zeros((len(sp_list), n_bins))
    for i, s in enumerate(sp_list):
        s = np.asarray(s, dtype=np.float64)
        s = s[(s >= t0) & (s < t1)] - t0
        arr[i, (s // dt).astype(int)] = 1
    rate = np.mean(arr, axis=0) * 1000 / dt
    if type(sigma) is not bool:
        rate = gaussian_filter1d(rate, sigma=sigma)
    return np.arange(t0, t1, dt), rate


def binned_trains_revision(trains, idcs, tlim, delta_t):
    """codes_revision/AnalysesTools.py:205-221"""
    t0, t1 = tlim
    t1 = np.ceil(t1).astype(int)
    arr = np.zeros((len(idcs), int((t1 - t0) // delta_t)))
    for q, i in enumerate(idcs):
        tr = np.asarray(trains[i], dtype=np.float64)
        tr = tr[(tr >= t0) & (tr < t1)]
        arr[q, ((tr - t0) // delta_t).astype(int)] = 1
    return arr.astype(int)


def shifted_corr(x, y, lag):
    """`x.corr(y.shift(periods=lag))` of two pandas Series (codes_revision/AnalysesTools.py:231-240): Pearson coefficient
    (np.corrcoef) over the rows where the shifted series exists."""
    n = len(x)
    if lag > 0:
        a, b = x[lag:], y[:n - lag]
    elif lag < 0:
        a, b = x[:n + lag], y[-lag:]
    else:
        a, b = x, y
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.corrcoef(a.astype(np.float64), b.astype(np.float64))[0, 1]


def correlations_revision(trains, tlim, idcs, delta_t, lags):
    """codes_revision/AnalysesTools.py:227-262 (ISIcorrelations): per lag, mean and std of the auto- and of the
    cross-correlations over all ordered (target, column) pairs"""
    arr = binned_trains_revision(trains, idcs, tlim, delta_t)
    out = [[], [], [], []]
    for l in lags:
        auto, cross = [], []
        for i in range(arr.shape[0]):
            for j in range(arr.shape[0]):
                (auto if i == j else cross).append(shifted_corr(arr[i], arr[j], l))
        for o, v in zip(out, (np.mean(auto), np.std(auto), np.mean(cross), np.std(cross))):
            o.append(v)
    return (np.asarray(lags) * delta_t,) + tuple(np.asarray(o) for o in out)


def lagged_counts(arr, lags):
    """The integer statistics hhd_an_lagged_counts returns, by dot products."""
    arr = np.asarray(arr, dtype=np.int64)
    n, T = arr.shape
    sxy = np.zeros((len(lags), n, n), np.int32)
    head = np.zeros((len(lags), n), np.int32)
    tail = np.zeros((len(lags), n), np.int32)
    for q, L in enumerate(lags):
        sxy[q] = arr[:, :T - L] @ arr[:, L:].T
        head[q] = arr[:, :T - L].sum(axis=1)
        tail[q] = arr[:, L:].sum(axis=1)
    return sxy, head, tail


def isi_stats(trains, tlim=None):
    """codes_revision/AnalysesTools.py:407-429"""
    mean, cv = [], []
    for tr in trains:
        tr = np.asarray(tr, dtype=np.float64)
        if tlim is not None:
            tr = tr[(tr >= tlim[0]) & (tr < tlim[1])]
        isi = np.diff(tr)
        with np.errstate(divide="ignore", invalid="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mean.append(np.mean(isi))
                cv.append(np.std(isi) / np.mean(isi))
    return np.asarray(mean), np.asarray(cv)
