"""Drop-in for `codes/CortexNetwork.py` of the reference: same constructor, `neuron_monitors`, `poisson_input`,
`regular_input`, `run`, group/synapse index helpers and monitor attributes — with the Brian 2 objects replaced by
the sm_100a engine behind include/hhd.h.

    from CortexNetwork import *            # as codes/Main.py:1 does; put this directory first on sys.path
    cortex = CortexNetwork(NeuPar, V0, STypPar, SynPar, SPMtx, group_distr, constant_current, fluctuating_current,
                           scales2, noise, method, time_step, simseed, simulation_dir)      # codes/Main.py:593
    cortex.neuron_monitors(NeuronMonitor); cortex.poisson_input(PoissonStimuli); cortex.run(Duration*ms)

What is NOT here: the ~3600 lines of statistics and matplotlib figures of the reference class (host-side
post-processing, out of scope); they read only `spikemonitor`, `neuronmonitors`, `NeuPar`, `group_distr`, which keep
their meaning, so the reference's own analysis methods can be mixed in unchanged.
"""
import os

import numpy as np

from ..engine import Engine, channels_from_STypPar, gsyn_factors, neuron_rows_from_NeuPar
from ..monitors import SpikeMonitorView, StateMonitorView
from ..units import *  # noqa: F401,F403  (the scripts get ms, mV, pA ... through `from CortexNetwork import *`)
from .. import stimuli, units

GROUP_NAMES = ["PC_L23", "IN_L_L23", "IN_L_d_L23", "IN_CL_L23", "IN_CL_AC_L23", "IN_CC_L23", "IN_F_L23",
               "PC_L5", "IN_L_L5", "IN_L_d_L5", "IN_CL_L5", "IN_CL_AC_L5", "IN_CC_L5", "IN_F_L5"]
_ALIASES = {"PC": [0, 7], "IN": [1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13], "IN_L": [1, 8], "IN_L_d": [2, 9],
            "IN_CL": [3, 10], "IN_CL_AC": [4, 11], "IN_CC": [5, 12], "IN_F": [6, 13], "IN_L23": [1, 2, 3, 4, 5, 6],
            "IN_L5": [8, 9, 10, 11, 12, 13], "L_23": list(range(7)), "L_5": list(range(7, 14)), "all": list(range(14))}
_SEED = {"value": None}


def seed(value=None):
    """Brian's `seed()` (codes/CortexNetwork.py:98, codes/Main.py:641): remembered and used as the engine's Philox key."""
    _SEED["value"] = value


class _Clock:
    dt = 0.1 * units.ms


defaultclock = _Clock()          # `defaultclock.dt = ...` in codes/Main.py:640 is a no-op there as well


class BrianLogger:               # codes/Main.py:10
    @staticmethod
    def suppress_name(*a, **k):
        pass


def sourcetarget(SPMtx):
    """(target_arr, source_arr) from the dense synapse-index matrix, as codes/CortexNetwork.py:4378-4409 builds them:
    entry [target, source, 0] is the 1-based index of the first synapse of the pair, [.., 1] of its NMDA twin, which is
    only meaningful where layer 0 is set."""
    S = np.asarray(SPMtx)
    n = int(np.max(S)) if S.size else 0
    target_arr = np.zeros(n, dtype=np.int64)
    source_arr = np.zeros(n, dtype=np.int64)
    tg, sr = np.nonzero(S[:, :, 0] > 0)
    first = S[tg, sr, 0].astype(np.int64) - 1
    target_arr[first] = tg
    source_arr[first] = sr
    twin = S[tg, sr, 1].astype(np.int64)
    has = twin > 0
    target_arr[twin[has] - 1] = tg[has]
    source_arr[twin[has] - 1] = sr[has]
    return target_arr, source_arr


class _Group:
    """The few NeuronGroup attributes scripts touch (N, t, parameters in SI)."""

    def __init__(self, owner):
        self._o = owner

    @property
    def N(self):
        return self._o.NeuPar.shape[1]

    @property
    def t(self):
        return self._o.engine.current_step() * self._o.time_step * units.ms

    @property
    def V(self):
        return self._o.engine.read_state("V") * units.mV

    @property
    def w(self):
        return self._o.engine.read_state("w") * units.pA


class CortexNetwork:
    def __init__(self, NeuPar, V0, STypPar, SynPar, SPMtx, group_distr, constant_current, fluctuating_current,
                 scales2, noise, method, time_step, simseed, simulation_dir, accum="fp64", device=0, plan="auto",
                 _engine_cls=None):
        self.noise = noise if (type(noise) is int or type(noise) is float) else None   # codes/CortexNetwork.py:157
        seed(simseed)
        self.time_step = float(time_step)
        self.group_distr = group_distr
        self.NeuPar = np.asarray(NeuPar, dtype=np.float64)
        self.SynPar = np.array(SynPar, dtype=np.float64, copy=True)
        self.STypPar = np.asarray(STypPar, dtype=np.float64)
        self.sim_dir = simulation_dir
        if simulation_dir and not os.path.isdir(simulation_dir):
            os.makedirs(simulation_dir, exist_ok=True)
        g = gsyn_factors(self.STypPar)
        self.Gsyn_AMPA, self.Gsyn_GABA, self.Gsyn_NMDA = (float(x) for x in g)
        n = self.NeuPar.shape[1]

        # group.I_DC[...] = constant_current[stripe][group]*pA     (codes/CortexNetwork.py:314-318)
        self.I_DC = np.zeros(n)
        if len(constant_current):
            for stripe in range(len(group_distr[0])):
                for grp in range(len(group_distr)):
                    idx = np.asarray(group_distr[grp][stripe]).astype(int)
                    self.I_DC[idx] = constant_current[stripe][grp]

        if isinstance(SPMtx, (tuple, list)) and len(SPMtx) == 2:      # sparse form: (target_arr, source_arr)
            self.target_arr, self.source_arr = (np.asarray(a).astype(np.int64) for a in SPMtx)
        else:
            self.target_arr, self.source_arr = sourcetarget(SPMtx)

        # scales2: rescale g_max by source / target group after loading (codes/CortexNetwork.py:434-460)
        if self.SynPar.size:
            sA, sG, sN, tA, tG, tN = scales2
            for lst, stype, as_source in ((sA, 1, True), (sG, 2, True), (sN, 3, True), (tA, 1, False), (tG, 2, False), (tN, 3, False)):
                for group, stripe, scale in lst:
                    name = ["all", "all", group, stripe] if as_source else [group, stripe, "all", "all"]
                    idc = self.group_name_to_synapses_idc([name])
                    idc = idc[self.SynPar[0, idc] == stype] if len(idc) else idc
                    self.SynPar[4, idc] *= scale

        method = {"gsl_rk2": "rk23"}.get(method, method)
        if method not in ("euler", "rk2", "rk4", "rk23"):
            raise ValueError("integration method %r is not available (euler, rk2, rk4, gsl_rk2)" % (method,))
        s = simseed if simseed is not None else int.from_bytes(os.urandom(8), "little")
        self.engine = (_engine_cls or Engine)(n, self.time_step, method=method, accum=accum, plan=plan, refractory_ms=5.0,
                             block_window_ms=5.0, last_spike0_ms=-1e8, seed=s, device=device)
        V0 = np.asarray(V0, dtype=np.float64)
        self.engine.set_neurons(neuron_rows_from_NeuPar(self.NeuPar, self.I_DC), V0[0], V0[1])
        self.engine.set_channels(**channels_from_STypPar(self.STypPar))
        if self.SynPar.size:
            chan = (self.SynPar[0].astype(np.int64) - 1).astype(np.uint8)
            self.engine.set_synapses(self.source_arr, self.target_arr, chan, self.SynPar[4], self.SynPar[1],
                                     self.SynPar[2], self.SynPar[3], self.SynPar[6], self.SynPar[5])
        if self.noise is not None:
            # group.noise_D = noise; group.nCon[i] = incoming + outgoing synapses of i   (codes/CortexNetwork.py:391-394, 838-842);
            # the equations add xi * sqrt(2 * noise_D * nCon) * nS/sqrt(ms) to dg_X_off/dt (:182-189)
            nCon = np.bincount(self.target_arr, minlength=n) + np.bincount(self.source_arr, minlength=n)
            self.engine.set_noise(np.sqrt(2 * self.noise * nCon.astype(np.float64)))
        # fluctuating_current = [start_ms, stop_ms, AC[stripe][group] expression strings in t (ms)] -> TimedArray `ta`
        # (codes/CortexNetwork.py:133-151): zeros before `start`, one sample per time step, last value held afterwards
        self.ta = []
        if len(fluctuating_current):
            start_f, stop_f, AC = fluctuating_current
            nsteps = int(round(stop_f / self.time_step, 0))
            nstart = int(round(start_f / self.time_step, 0))
            ns = {k: getattr(np, k) for k in dir(np) if not k.startswith("_")}
            cols, col_of = [], np.full(n, -1, dtype=np.int32)
            for stripe in range(len(group_distr[0])):
                for grp in range(len(group_distr)):
                    expr = AC[stripe][grp]
                    if expr != "":
                        tt = np.linspace(start_f, stop_f, nsteps - nstart, endpoint=False)
                        vals = np.asarray([eval(expr, dict(ns), {"t": t}) for t in tt], dtype=np.float64)
                        col_of[np.asarray(group_distr[grp][stripe]).astype(int)] = len(cols)
                        cols.append(np.concatenate([np.zeros(nstart), vals]))
            if cols:
                self.ta = np.stack(cols, axis=1)
                self.engine.set_timed_current(self.ta, col_of)
        self.group = _Group(self)
        self.I_SN = self.NeuPar[1] * units.nS * (self.NeuPar[9] - self.NeuPar[2] - self.NeuPar[3]) * units.mV   # :339
        self.spikemonitor = SpikeMonitorView(self.engine, n)
        self.neuronmonitors = {}
        self.synapsesmonitors = {}
        self._n_ext_sources = 0

    # ---- monitors ---------------------------------------------------------------------------------------------------
    def neuron_monitors(self, MonitorsSetup):
        """MonitorsSetup: [[name, variables, [[group, stripe], ...]], ...]   (codes/CortexNetwork.py:499-512)"""
        for monitor in MonitorsSetup:
            sel = monitor[2]
            if len(sel) == 1 and sel[0][0] == "all" and sel[0][1] == "all":
                idc = True
            else:
                idc = self.group_name_to_neuron_idc(sel)
            self.neuronmonitors[monitor[0]] = StateMonitorView(self.engine, monitor[1], idc)

    def synapses_monitors(self, MonitorsSetup):
        # the reference records from the neuron group here as well (codes/CortexNetwork.py:521)
        raise NotImplementedError("synapses_monitors records neuron variables at synapse indices in the reference "
                                  "(codes/CortexNetwork.py:521) and is unused by every script")

    # ---- external input -------------------------------------------------------------------------------------------------
    def poisson_input(self, PoissonInput):
        """[[N, type, rate_Hz, g_max_nS, p_fail, start_ms, stop_ms, [[group, stripe, pCon], ...]], ...]
        (codes/CortexNetwork.py:525-622); drawn by stimuli.draw_poisson_input with np.random in the reference's order."""
        s = stimuli.draw_poisson_input(PoissonInput, self.group_distr, self.time_step)
        stimuli.attach(self.engine, s, self.time_step)
        self.poisson_spikes = (s["spk_src"], s["spk_t"] * units.ms)

    def regular_input(self, RegularInput):
        """[[N, type, n_spikes, g_max_nS, p_fail, start_ms, stop_ms, targets], ...]   (codes/CortexNetwork.py:624-719)"""
        s = stimuli.draw_regular_input(RegularInput, self.group_distr, self.time_step)
        stimuli.attach(self.engine, s, self.time_step)
        self.regular_spikes = (s["spk_src"], s["spk_t"] * units.ms)

    # ---- run --------------------------------------------------------------------------------------------------------------
    def run(self, t):
        """`t` in seconds, i.e. `Duration*ms` as the scripts write it (codes/CortexNetwork.py:721-729)."""
        n_steps = int(round(float(t) / (self.time_step * units.ms)))
        self.engine.run(n_steps)

    # ---- index helpers (codes/CortexNetwork.py:820-984) ---------------------------------------------------------------
    def get_source_from_target(self, ind):
        return self.source_arr[np.where(np.isin(self.target_arr, ind))[0]]        # scalar or array, as the reference (:820-826)

    def get_target_from_source(self, ind):
        return self.target_arr[np.where(np.isin(self.source_arr, ind))[0]]

    def get_connections_count(self, ind):
        """incoming + outgoing synapses (codes/CortexNetwork.py:838-842: the nCon of the noise term)"""
        return len(self.get_source_from_target(ind)) + len(self.get_target_from_source(ind))

    def get_neurons_from_synapses(self, ind):
        return self.target_arr[ind], self.source_arr[ind]

    def group_name_to_group_idc(self, group):
        if group in _ALIASES:
            return list(_ALIASES[group])
        return [GROUP_NAMES.index(group)]

    def group_idc_to_group_name(self, idc):
        return GROUP_NAMES[idc]

    def neuron_idc_to_group_name(self, idc):
        for group in range(len(self.group_distr)):
            for stripe in range(len(self.group_distr[0])):
                if idc in self.group_distr[group][stripe]:
                    return self.group_idc_to_group_name(group), stripe

    def group_name_to_neuron_idc(self, name):
        idcs = []
        for grname, stripe in name:
            if grname == "all" and stripe == "all":
                for gr in range(len(self.group_distr)):
                    for strp in range(len(self.group_distr[gr])):
                        idcs.extend(self.group_distr[gr][strp])
                continue
            if isinstance(grname, str):
                groups = self.group_name_to_group_idc(grname)
            elif isinstance(grname, (int, np.integer)):
                groups = [int(grname)]
            else:
                groups = list(grname)
            stripes = range(len(self.group_distr[0])) if stripe == "all" else [stripe]
            for g in groups:
                for s in stripes:
                    idcs.extend(self.group_distr[g][s])
        return np.unique(np.asarray(idcs)).astype(int)

    def group_name_to_synapses_idc(self, names, neuron_index=False):
        both = []
        nsyn = len(self.source_arr)
        for name in names:
            if neuron_index:
                target_idc, source_idc = name
                if isinstance(target_idc, str) and target_idc == "all":
                    target_idc = np.arange(self.NeuPar.shape[1])
                if isinstance(source_idc, str) and source_idc == "all":
                    source_idc = np.arange(self.NeuPar.shape[1])
            else:
                tname, tstripe, sname, sstripe = name
                target_idc = self.group_name_to_neuron_idc([[tname, tstripe]])
                source_idc = self.group_name_to_neuron_idc([[sname, sstripe]])
            hit = np.isin(self.target_arr, target_idc) & np.isin(self.source_arr, source_idc)
            both.extend(np.arange(nsyn)[hit])
        return np.asarray(sorted(both), dtype=np.int64)

    def _typed(self, names, stype, neuron_index):
        idc = self.group_name_to_synapses_idc(names, neuron_index=neuron_index)
        return idc[self.SynPar[0, idc] == stype] if len(idc) else idc

    def group_name_to_AMPAsynapses_idc(self, names, neuron_index=False):
        return self._typed(names, 1, neuron_index)

    def group_name_to_GABAsynapses_idc(self, names, neuron_index=False):
        return self._typed(names, 2, neuron_index)

    def group_name_to_NMDAsynapses_idc(self, names, neuron_index=False):
        return self._typed(names, 3, neuron_index)

    # ---- a first analysis, enough to check the published statistics (codes/CortexNetwork.py:2713-2765) ---------------
    def ISI_stats(self, start_ms=0.0, min_spikes=20):
        """Mean, std and CV of the inter-spike intervals of neurons with more than `min_spikes` spikes after `start_ms`
        (the selection of codes/Original_ISI.py:13-40 and content.tex:318)."""
        trains = self.spikemonitor.spike_trains()
        means, cvs = [], []
        for tr in trains.values():
            tr = tr[tr >= start_ms * units.ms] / units.ms
            if len(tr) > min_spikes:
                isi = np.diff(tr)
                means.append(isi.mean())
                cvs.append(isi.std() / isi.mean())
        return np.asarray(means), np.asarray(cvs)
