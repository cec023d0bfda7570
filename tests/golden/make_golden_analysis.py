"""Golden vectors for the spike-train analyses, produced by the REFERENCE's own functions
(/root/reference/codes/AuxiliarEquations.py and codes_revision/AnalysesTools.py, imported with stubs for the modules
they import but do not need here: brian2, matplotlib, scipy.integrate.quadrature).  Run in the build container:
    python tests/golden/make_golden_analysis.py
writes tests/golden/analysis_golden.npz (inputs + outputs)."""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    _stub("matplotlib", pyplot=_stub("matplotlib.pyplot"))
    _stub("brian2", ms=1e-3, pA=1e-12)
    import scipy.integrate
    if not hasattr(scipy.integrate, "quadrature"):
        scipy.integrate.quadrature = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    sys.path.insert(0, os.path.join(REF, "codes"))
    import AuxiliarEquations as AE
    sys.path.pop(0)
    sys.path.insert(0, os.path.join(REF, "codes_revision"))
    import AnalysesTools as AT
    sys.path.pop(0)

    rng = np.random.default_rng(20161)
    dt = 0.05
    n, T = 7, 400.0
    trains = []
    for i in range(n):
        k = int(rng.integers(0, 40)) if i != 3 else 0                 # one silent train
        steps = np.sort(rng.choice(np.arange(int(T / dt)), size=k, replace=False))
        trains.append(steps * dt)
    trains[5] = np.concatenate([trains[5], [120.0, 120.05, 121.95]])   # two spikes in one bin, bin edges
    trains[5].sort()
    out = {"n": n, "dt": dt}
    for i, t in enumerate(trains):
        out["train_%d" % i] = t

    t0, t1, t_bin = 10.0, 390.0, 2.0
    out["codes_args"] = np.array([t0, t1, t_bin, -10.0, 12.0])
    out["codes_binary_5"] = AE.set_binary_vector(trains[5], t0, t1, t_bin)
    group = [trains[i] for i in (0, 1, 2, 4, 5, 6)]
    lag_array, corr = AE.Pearson_correlation_group(group, t0, t1, t_bin, -10.0, 12.0)
    out["codes_group_lags"], out["codes_group_corr"] = lag_array, corr
    out["codes_zero_lag_group"] = np.asarray(AE.Zero_lag_Pearson_correlation_group(group, t0, t1, t_bin))
    la, c01 = AE.Pearson_correlation(trains[0], trains[1], t0, t1, t_bin, -10.0, 12.0)
    out["codes_pair01_corr"] = c01
    tt, rate = AE.pop_rate(trains, 0.0, 400.0, 5.0)
    out["codes_pop_rate_t"], out["codes_pop_rate"] = tt, rate
    tt, rate = AE.pop_rate(trains, 0.0, 400.0, 5.0, sigma=2.0)
    out["codes_pop_rate_sigma2"] = rate

    idcs = [0, 1, 2, 4, 5, 6]
    lags = [-3, 0, 1, 4]
    res = AT.get_correlations_from_spike_trains(trains, (10, 389.3), idcs=idcs, delta_t=2, lags=lags)
    for name, v in zip(("lags", "auto_mean", "auto_std", "cross_mean", "cross_std"), res):
        out["rev_corr_" + name] = np.asarray(v)
    out["rev_corr_args"] = np.array([10, 389.3, 2])
    out["rev_corr_idcs"] = np.array(idcs)
    out["rev_corr_lags_in"] = np.array(lags)
    mean, cv = AT.get_ISI_stats_from_spike_trains(trains, neuron_idcs=[0, 1, 2, 4, 5, 6], tlim=(5.0, 395.0))
    out["rev_isi_mean"], out["rev_isi_cv"] = np.asarray(mean), np.asarray(cv)
    # ---- the reference's own data + script: codes/Original_spiketime.txt analysed as codes/Original_ISI.py:11-57 does
    with open(os.path.join(REF, "codes", "Original_spiketime.txt")) as f:
        spike_list = [np.asarray(tr.split(",")).astype(float) for tr in f.read().split(";")]
    stset = []
    for tr in spike_list:
        tr = tr[tr >= 1000]
        if len(tr) > 20:
            stset.append(tr)
    stset = stset[:48]                                               # a subset keeps the fixture small
    out["orig_n"] = len(stset)
    out["orig_trains"] = np.concatenate(stset)
    out["orig_offsets"] = np.cumsum([0] + [len(tr) for tr in stset])
    out["orig_isi_mean"] = np.asarray([np.mean(np.diff(tr)) for tr in stset])
    out["orig_isi_cv"] = np.asarray([np.std(np.diff(tr)) / np.mean(np.diff(tr)) for tr in stset])
    lag, corr = AE.Pearson_correlation_group(stset, 1000, 61000, 2, -30, 30)
    out["orig_corr_lags"], out["orig_corr"] = lag, corr
    # ---- membrane-potential statistics with spike excision and the LFP spectrum (codes_revision/AnalysesTools.py:88-172, 343-373) on a
    # stand-in `cortex` that exposes what those functions touch (recorded.V.V / .t, recorded.var / .t, neurons.N, network.membr_params)
    class _Loc:
        def __init__(self, vt):
            self.vt = vt

        def __getitem__(self, sel):
            return types.SimpleNamespace(values=self.vt[sel["cell_index"]])

    rngv = np.random.default_rng(77)
    nV, nT = 6, 4000
    V_T = np.array([-50.0, -48.0, -52.0, -45.0, -55.0, -50.0])
    Vtr = np.zeros((nV, nT))
    for i in range(nV):
        v, trace = -70.0, []
        for k in range(nT):
            v += rngv.normal(0.02 * (i + 1), 0.4) + (0.25 * (v - V_T[i]) if v > V_T[i] else 0.0)
            if v > -20.0:
                trace.append(-20.0 + rngv.random())
                v = -75.0 - 5 * rngv.random()                      # reset: |V| jumps by far more than 1 mV
            else:
                trace.append(v)
        Vtr[i] = trace
    Vtr[4] = -80.0 + 0.01 * rngv.normal(size=nT)                   # one cell that never spikes
    tV = np.arange(nT) * 0.05
    Itot = np.cumsum(rngv.normal(size=(3, nT)), axis=1) + 200.0
    fake = types.SimpleNamespace(
        recorded=types.SimpleNamespace(V=types.SimpleNamespace(V=Vtr * 1e-3, t=tV * 1e-3),
                                       var=lambda name: Itot * 1e-12, t=lambda name: tV * 1e-3),
        neurons=types.SimpleNamespace(N=nV),
        network=types.SimpleNamespace(membr_params=types.SimpleNamespace(loc=_Loc(V_T))))
    import brian2 as _b
    _b.mV = 1e-3
    AT.br2.mV, AT.br2.ms, AT.br2.pA = 1e-3, 1e-3, 1e-12
    out["vstats_V"], out["vstats_t"], out["vstats_VT"], out["lfp_Itot"] = Vtr, tV, V_T, Itot
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        out["vstats_all"] = np.asarray(AT.get_V_stats(fake))
        out["vstats_sel"] = np.asarray(AT.get_V_stats(fake, neuron_idcs=[1, 3, 4], tlim=(20.0, 150.0)))
    t_lfp, lfp = AT.get_LFP(fake, invert_Itot=True, population_agroupate="sum")
    out["lfp_t"], out["lfp_sum_inverted"] = t_lfp, lfp
    fake1 = types.SimpleNamespace(recorded=types.SimpleNamespace(var=lambda name: Itot[0] * 1e-12, t=lambda name: tV * 1e-3))
    fq, pw = AT.get_LFP_SPD(fake1, log=True, sigma=2)
    out["lfp_spd_f"], out["lfp_spd_p"] = fq, pw
    np.savez_compressed(os.path.join(HERE, "analysis_golden.npz"), **out)
    print("wrote analysis_golden.npz:", sorted(out))


if __name__ == "__main__":
    main()
