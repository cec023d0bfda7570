"""The revision's spike-train statistics (`codes_revision/AnalysesTools.py`), same names and arguments, with the N^2
`Series.corr` calls per lag (the 81-minute step of Tests.py:67) replaced by one device call per analysis (libhhd.so,
`hhd_an_*`) and the Pearson formula on its integer counts (analysis.pearson_shifted).

  get_correlations / get_correlations_from_spike_trains / ISIcorrelations   codes_revision/AnalysesTools.py:174-295
  get_ISI_stats / get_ISI_stats_from_spike_trains / ISIstats               :392-441
  get_LFP_SPD_from_Itotarray                                                :375-389 (host: one FFT)
"""
import numpy as np
from scipy.ndimage import gaussian_filter1d as gf1d
from scipy.signal import periodogram

from .. import analysis as an
from .. import units


def _binned(spike_trains, idcs, tlim, delta_t, device):
    if idcs is None:
        idcs = np.arange(len(spike_trains))
    elif isinstance(idcs, (int, np.integer)):
        idcs = [idcs]
    t0, t1 = tlim
    t1 = np.ceil(t1).astype(int)
    n_bins = int((t1 - t0) // delta_t)
    neuron, t = an.spikes_from_trains(spike_trains, list(idcs))
    bits, _, _ = an.binned_trains(neuron, t, np.arange(len(idcs)), t0, t1, delta_t, an.BIN_FLOORDIV, n_bins, device=device)
    return bits, n_bins


def _correlations(bits, n_bins, delta_t, lags, file):
    if isinstance(lags, (int, np.integer)):
        lags = [lags]
    lags = [int(l) for l in lags]
    absl = sorted(set(abs(l) for l in lags))
    sxy, head, tail = an.lagged_counts(bits, n_bins, absl)
    pos = {L: q for q, L in enumerate(absl)}
    n = bits.shape[0]
    off = ~np.eye(n, dtype=bool)
    auto_mean, auto_std, cross_mean, cross_std = [], [], [], []
    for l in lags:
        q = pos[abs(l)]
        r = an.pearson_shifted(sxy[q], head[q], tail[q], n_bins, l)      # [target, column]
        auto, cross = np.diagonal(r).copy(), r[off]                       # the reference's list order: target-major
        auto_mean.append(np.mean(auto))
        auto_std.append(np.std(auto))
        cross_mean.append(np.mean(cross))
        cross_std.append(np.std(cross))
    lags = np.asarray(lags) * delta_t
    out = lags, np.array(auto_mean), np.array(auto_std), np.array(cross_mean), np.array(cross_std)
    if file is not None:
        with open(file, 'w') as f:
            print('Auto-correlations:', file=f, end='\n\n')
            for lag in zip(lags, out[1], out[2]):
                print('lag {} ms --> mean: {:.4f}, std: {:.4f}'.format(*lag), file=f)
            print(file=f)
            print('Auto-correlations:', file=f, end='\n\n')
            for lag in zip(lags, out[3], out[4]):
                print('lag {} ms --> mean: {:.4f}, std: {:.4f}'.format(*lag), file=f)
    return out


def ISIcorrelations(bin_arr, delta_t, lags, file=None, display=False, display_interval=10):
    bin_arr = np.asarray(bin_arr)
    return _correlations(an.pack_bits(bin_arr), bin_arr.shape[1], delta_t, lags, file)


def get_correlations_from_spike_trains(spike_trains, tlim, idcs=None, delta_t=2, lags=0, file=None, display=False,
                                       display_interval=10, device=0):
    bits, n_bins = _binned(spike_trains, idcs, tlim, delta_t, device)
    return _correlations(bits, n_bins, delta_t, lags, file)


def _trains_ms(cortex):
    return {k: v / units.ms for k, v in cortex.spikemonitor.spike_trains().items()}


def get_correlations(cortex, tlim=None, idcs=None, delta_t=2, lags=0, file=None, display=False, display_interval=10, device=0):
    if idcs is None:
        idcs = np.arange(cortex.neurons.N)
    if tlim is None:
        tlim = (cortex.transient, cortex.neurons.t / units.ms)
    return get_correlations_from_spike_trains(_trains_ms(cortex), tlim, idcs, delta_t, lags, file, device=device)


def ISIstats(spike_trains, savetxt=None, device=0):
    idcs = list(range(len(spike_trains)))
    neuron, t = an.spikes_from_trains(spike_trains, idcs)
    mean, cv, _ = an.isi_stats(neuron, t, idcs, device=device)
    ISImean, ISICV = list(mean), list(cv)
    if savetxt is not None:
        with open(savetxt, 'w') as f:
            print('ISI stats\n', file=f)
            print('Cells:', len(ISImean), file=f)
            print('ISImean mean: {:.3f} ms'.format(np.mean(ISImean)), file=f)
            print('ISImean std: {:.3f}ms'.format(np.std(ISImean)), file=f)
            print('ISICV mean: {:.3f}'.format(np.mean(ISICV)), file=f)
            print('ISICV std: {:.3f}'.format(np.std(ISICV)), file=f)
    return ISImean, ISICV


def get_ISI_stats_from_spike_trains(spike_trains, neuron_idcs=None, tlim=None, savetxt=None, device=0):
    if neuron_idcs is None:
        neuron_idcs = np.arange(len(spike_trains))
    trains = [np.asarray(spike_trains[idc]) for idc in neuron_idcs]
    if tlim is not None:
        t0, t1 = tlim
        trains = [train[(train >= t0) & (train < t1)] for train in trains]
    return ISIstats(trains, savetxt, device=device)


def get_ISI_stats(cortex, neuron_idcs=None, tlim=None, savetxt=None, device=0):
    trains = _trains_ms(cortex)
    if neuron_idcs is None:
        neuron_idcs = np.arange(len(trains))
    return get_ISI_stats_from_spike_trains(trains, neuron_idcs, tlim, savetxt, device=device)


def get_LFP_SPD_from_Itotarray(Itotarray, fq, log=True, sigma=0):
    frequency, power = periodogram(Itotarray, fq)
    if log:
        power = np.log(power[frequency > 0])
        frequency = np.log(frequency[frequency > 0])
    if sigma > 0:
        power = gf1d(power, sigma)
    return frequency, power


# ---- membrane potential and LFP (host side: they work on recorded traces, not on spike lists) ----------------------------------
def _without_spikes(Varr, V_T, spike_jump=1):
    """A V trace with its spikes cut out (codes_revision/AnalysesTools.py:92-123): the trace is split where |V| jumps by at least
    `spike_jump` mV from one sample to the next (the reset); every piece but the last ends at its last sample below V_T, i.e. loses
    the upswing into the spike — or vanishes if it never is below V_T."""
    after = np.where(np.diff(np.abs(Varr)) >= spike_jump)[0] + 1
    pieces, start = [], 0
    for stop in after:
        piece = Varr[start:stop]
        below = np.where(piece < V_T)[0]
        pieces.append(piece[:below[-1] + 1] if len(below) else piece[:0])
        start = stop
    pieces.append(Varr[start:])
    return np.concatenate(pieces)


def get_V_stats(cortex, neuron_idcs=None, tlim=None, file=None):
    """-> Vmean, Vstd, mean(V - V_T), and the same three with the spikes cut out, one entry per neuron (mV); `file`: the reference's
    text report (codes_revision/AnalysesTools.py:88-172).  Needs a 'V' monitor (`set_neuron_monitors('V', 'V', ...)`)."""
    V = np.asarray(cortex.recorded.V.V) / units.mV
    if neuron_idcs is not None:
        V = V[neuron_idcs]
    else:
        neuron_idcs = np.arange(cortex.neurons.N)
    if tlim is not None:
        t = np.asarray(cortex.recorded.V.t) / units.ms
        V = V[:, (t >= tlim[0]) & (t < tlim[1])]
    rows = [[] for _ in range(6)]
    for i in range(V.shape[0]):
        V_T = cortex.network.membr_params.loc[dict(param="V_T", cell_index=neuron_idcs[i])].values
        cut = _without_spikes(V[i], V_T)
        for row, val in zip(rows, (np.mean(V[i]), np.std(V[i]), np.mean(V[i] - V_T), np.mean(cut), np.std(cut), np.mean(cut - V_T))):
            row.append(val)
    if file is not None:
        names = ("Vmean", "Vstd", "V minus V_T mean")
        with open(file, "w") as f:
            print("V correlations", file=f)
            print("{} cells".format(len(neuron_idcs)), end="\n\n", file=f)
            for title, block, suffix in (("Full V arrays", rows[:3], ""), ("V arrays after spike extraction", rows[3:], "_extracted")):
                print(title, file=f)
                for name, row in zip(names, block):
                    print("{}{} --> mean: {:.2f} mV, std: {:.2f} mV".format(name, suffix, np.mean(row), np.std(row)), file=f)
                print(file=f)
    return tuple(rows)


def get_LFP(cortex, invert_Itot=True, population_agroupate=None):
    """The recorded summed current as the LFP proxy (codes_revision/AnalysesTools.py:343-356): -> t (ms), LFP (pA)."""
    LFP = np.asarray(cortex.recorded.var("I_tot")) / units.pA
    t_LFP = np.asarray(cortex.recorded.t("I_tot")) / units.ms
    if population_agroupate == "mean":
        LFP = np.mean(LFP, axis=0)
    elif population_agroupate == "sum":
        LFP = np.sum(LFP, axis=0)
    return t_LFP, (-LFP if invert_Itot else LFP)


def get_LFP_SPD(cortex, log=True, sigma=0):
    """Periodogram of the LFP proxy, sampling rate from the monitor's time axis (:358-373)."""
    t, LFP = get_LFP(cortex, invert_Itot=False)
    return get_LFP_SPD_from_Itotarray(LFP, 1000 / (t[1] - t[0]), log, sigma)
