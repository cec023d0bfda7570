"""`Cortex` of codes_revision/CortexSetup.py on the B200 engine: same constructor / `setup` / `load`, `run(t)` in ms with
the transient, the monitor schedule and the long-run monitors, `set_neuron_monitors`, `set_longrun_neuron_monitors`,
`set_regular_stimuli`, `set_poisson_stimuli`, `set_custom_stimuli`, index helpers, `spiking_count/rate/idcs`, `rheobase`.

Engine-side simplifications that keep the caller-visible behaviour:
  * Brian runs segment by segment because monitors can only be toggled between `net.run` calls
    (CortexSetup.py:126-189); the engine toggles monitors between `hhd_run` calls the same way;
  * long-run monitors spill 5 s chunks to `longrun/*.npy` and re-create the StateMonitor to bound host memory
    (:310-368); here a population-aggregated long-run monitor is a device-side reduction (one double per step), so
    nothing needs spilling — the holder is filled after the run with the same attributes (`.t`, `.<var>`).
"""
import numpy as np

from ..engine import Engine
from ..monitors import SpikeMonitorView, StateMonitorView
from .. import units
from .NetworkSetup import CHANNEL_NAMES, Network, network_setup


class NetworkHolder(dict):
    """Attribute-and-item holder (the reference's dataclass-based BaseClass)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class VarHolder(NetworkHolder):
    def var(self, var):
        return getattr(self[var], var)

    def t(self, var):
        return self[var].t


class _Neurons:
    def __init__(self, owner):
        self._o = owner

    @property
    def N(self):
        return self._o.engine.n_neurons

    @property
    def t(self):
        return self._o.engine.current_step() * self._o.dt_ms * units.ms


class Cortex:
    @staticmethod
    def setup(Ncells, Nstripes, constant_stimuli, method, dt, transient=0, basics_scales=None, fluctuant_stimuli=None, seed=None, **kw):
        network = network_setup(Ncells, Nstripes, basics_scales, seed)
        return Cortex(network, constant_stimuli, method, dt, transient, fluctuant_stimuli, seed, **kw)

    @staticmethod
    def load(path, constant_stimuli, method, dt, transient=0, fluctuant_stimuli=None, **kw):
        network = Network.load(path)
        return Cortex(network, constant_stimuli, method, dt, transient, fluctuant_stimuli, network.seed, **kw)

    def __init__(self, network, constant_stimuli, method, dt, transient=0, fluctuant_stimuli=None, seed=None,
                 last_spike0_ms=None, mg_fac=None, reference_quirks=False, accum="fp64", device=0, _engine_cls=None):
        """The model variant comes from the network (Network.channel_constants): a network drawn by the revision's generator
        carries the revision's BasicsSetup tree — NMDA Mg factor 1 (BasicsSetup.py:591), last_spike = -1000 ms
        (CortexSetup.py:254), block term relaxing to V_r (BasicsSetup.py:737); a network wrapped from codes/ arrays
        (Network.from_codes) keeps the codes/ channel table (Mg factor 0.33).  `mg_fac` / `last_spike0_ms` override either.
        reference_quirks=True additionally loads tau_fac from the tau_rec row, as BasicsSetup.py:447 makes the reference do.
        Not reproduced: the revision sums I_syn = I_AMPA + I_GABA + I_NMDA (:727), the engine (I_AMPA + I_NMDA) + I_GABA as
        codes/ does — a last-bit difference in I_tot."""
        if fluctuant_stimuli is not None and isinstance(fluctuant_stimuli[-1], (int, float)):
            fluctuant_stimuli = [fluctuant_stimuli]
        self.fluctuant_stimuli = fluctuant_stimuli
        np.random.seed(seed)
        self.network = network
        self.method = method
        self.dt_ms = float(dt)
        self.dt = self.dt_ms * units.ms
        if constant_stimuli is not None and isinstance(constant_stimuli[-1], (int, float)):
            constant_stimuli = [constant_stimuli]
        self.constant_stimuli = constant_stimuli
        self.transient = transient
        self.seed = seed
        n = network.basics.struct.Ncells_total
        I_DC = np.zeros(n)
        if constant_stimuli is not None:
            for target, value in constant_stimuli:
                I_DC[self.neuron_idcs(target)] = value
        P = network.membr_params.values
        rows = np.stack([P[0], P[1], P[2], P[3], P[4], P[5], P[6], P[7], P[8], network.refractory_current.values,
                         P[7],                  # the revision's block term relaxes to V_r (BasicsSetup.py:737) = V_dep of codes/
                         I_DC])
        m = {"gsl_rk2": "rk23"}.get(method, method)
        model = network.channel_constants()
        ch = {k: model[k] for k in ("tau_on", "tau_off", "E_rev", "gsyn", "mg_fac", "mg_slope", "mg_half")}
        if mg_fac is not None:
            ch["mg_fac"] = float(mg_fac)
        if last_spike0_ms is None:
            last_spike0_ms = model["last_spike0_ms"]
        self.reference_quirks = bool(reference_quirks)
        s = seed if seed is not None else int.from_bytes(__import__("os").urandom(8), "little")
        self.engine = (_engine_cls or Engine)(n, self.dt_ms, method=m, accum=accum, refractory_ms=5.0, block_window_ms=5.0,
                                              last_spike0_ms=last_spike0_ms, seed=s, device=device,
                                              record_spikes=not (transient > 0))
        self.engine.set_neurons(rows, P[2], np.zeros(n))                     # V = E_L, w = 0 (CortexSetup.py:256-258)
        self.engine.set_channels(**ch)
        self.gsyn_amp = dict(zip(CHANNEL_NAMES, ch["gsyn"]))
        sp = network.syn_params
        if network.syn_pairs.shape[1] > 0:
            self.engine.set_synapses(network.syn_pairs[1], network.syn_pairs[0], sp["channel"].values[0].astype(np.uint8),
                                     sp["spiking"].values[0], sp["STSP_params"].values[0], sp["STSP_params"].values[1],
                                     sp["STSP_params"].values[1 if (reference_quirks and model["tau_fac_from"] == "tau_rec") else 2],
                                     sp["spiking"].values[1], sp["delay"].values[0])
        # fluctuant_stimuli = [[target, function(t_ms) -> pA, start_ms, stop_ms], ...] -> TimedArray (CortexSetup.py:193-213).
        # As in the reference every entry re-creates the array, so only the LAST entry takes effect.
        self.fluctuant_array = None
        if fluctuant_stimuli is not None:
            for target, function, start, stop in fluctuant_stimuli:
                idcs = self.neuron_idcs(target)
                n_total = round(stop / self.dt_ms)
                n_start = round(start / self.dt_ms)
                arr = np.asarray([0] * n_start + [function(t) for t in np.linspace(start, stop, n_total - n_start, endpoint=False)], dtype=np.float64)
                col_of = np.full(n, -1, dtype=np.int32)
                col_of[idcs] = 0
                self.fluctuant_array = (arr[:, None], col_of)
            self.engine.set_timed_current(*self.fluctuant_array)
        self.neurons = _Neurons(self)
        self.spikemonitor = SpikeMonitorView(self.engine, n)
        self._spikes_on = not (transient > 0)
        self.external_stimuli = NetworkHolder()
        self.neuron_monitors = NetworkHolder()
        self.longrun_neuron_monitors = NetworkHolder()
        self.recorded = VarHolder()
        self.monitor_schedule = {}
        self.longrun_monitor_control = []

    def save(self, path):
        self.network.save(path)

    # ---- run (CortexSetup.py:109-190): segments between schedule points -----------------------------------------------
    def run(self, t, erase_longrun=True):
        t0 = self.neurons.t / units.ms
        points = list(self.monitor_schedule.keys()) + [self.transient]
        for lr in self.longrun_monitor_control:
            points.append(lr["start"])
            if lr["stop"] is not None:
                points.append(lr["stop"])
        points = np.asarray(points, dtype=float)
        points = np.unique(points[(points >= t0) & (points < t0 + t)])
        points = np.unique(np.concatenate([points, [t0, t0 + t]]))
        self._apply_schedule(t0)
        for a, b in zip(points[:-1], points[1:]):
            n_steps = int(round(b / self.dt_ms)) - int(round(a / self.dt_ms))
            self.engine.run(n_steps)
            if b >= self.transient and not self._spikes_on:
                self.engine.set_spike_monitor_active(True)            # spikemonitor.active = True  (:175)
                self._spikes_on = True
            self._apply_schedule(b)
        self._fill_longrun()

    def _apply_schedule(self, t1):
        if t1 in self.monitor_schedule:
            for mon in self.monitor_schedule[t1]["start"]:
                mon.set_active(True)
            for mon in self.monitor_schedule[t1]["stop"]:
                mon.set_active(False)
        for lr in self.longrun_monitor_control:
            if t1 == lr["start"]:
                lr["view"].set_active(True)
            if lr["stop"] is not None and t1 == lr["stop"]:
                lr["view"].set_active(False)

    def _fill_longrun(self):
        for lr in self.longrun_monitor_control:
            view, holder = lr["view"], lr["monitor"]
            holder["t"] = view.t
            for var in view.variables:
                vals = getattr(view, var)
                holder[var] = vals[0] if lr["population_agroupate"] is not None else vals

    # ---- monitors (:501-607) --------------------------------------------------------------------------------------------
    def _schedule(self, view, variables, start, stop):
        if self.transient > 0 or (start is not None and start > 0):
            view.set_active(False)
        for var in variables:
            self.recorded[var] = view
        if start is None:
            start = self.transient
        self.monitor_schedule.setdefault(start, dict(start=[], stop=[]))["start"].append(view)
        if stop is not None:
            self.monitor_schedule.setdefault(stop, dict(start=[], stop=[]))["stop"].append(view)

    def set_neuron_monitors(self, name, variables, groupstripe_list, start=None, stop=None):
        variables = [variables] if isinstance(variables, str) else list(variables)
        view = StateMonitorView(self.engine, variables, self.neuron_idcs(groupstripe_list))
        self.neuron_monitors[name] = view
        self._schedule(view, variables, start, stop)

    def set_longrun_neuron_monitors(self, name, variables, groupstripe_list, interval, start=None, stop=None, population_agroupate=None):
        """`interval` (the spill period of the reference) is accepted and ignored: aggregated monitors are reduced on
        the device, so a 61 s run of Sum(I_tot) is 9.8 MB (Tests.py:29 uses interval=5000, 'sum')."""
        variables = [variables] if isinstance(variables, str) else list(variables)
        if population_agroupate not in (None, "sum", "mean"):
            raise ValueError("population_agroupate must be None, 'sum' or 'mean'")
        view = StateMonitorView(self.engine, variables, self.neuron_idcs(groupstripe_list), reduce=population_agroupate or "none")
        view.set_active(False)
        holder = NetworkHolder()
        holder["t"] = None
        for var in variables:
            holder[var] = None
            self.recorded[var] = holder
        self.longrun_neuron_monitors[name] = view
        self.neuron_monitors[name] = holder
        if start is None:
            start = self.transient
        if start <= self.neurons.t / units.ms:
            view.set_active(True)
        self.longrun_monitor_control.append(dict(start=start, stop=stop, interval=interval, view=view, monitor=holder,
                                                 population_agroupate=population_agroupate))

    # ---- stimuli (:370-500) ------------------------------------------------------------------------------------------------
    def set_custom_stimuli(self, name, Nsource, channel, spike_idcs, spike_times, pairs_connected, gmax, pfail):
        """One input synapse per (pair, channel) (:377-382); spike_times in ms; pairs_connected[0] targets, [1] sources."""
        channel = [channel] if isinstance(channel, str) else list(channel)
        pairs = np.asarray(pairs_connected, dtype=np.int64)
        nc = pairs.shape[1]
        gmax = [gmax] if isinstance(gmax, (int, float)) else list(gmax)
        pfail = [pfail] if isinstance(pfail, (int, float)) else list(pfail)
        g = np.asarray(gmax * nc if len(gmax) == 1 else gmax, dtype=float)
        pf = np.asarray(pfail * nc if len(pfail) == 1 else pfail, dtype=float)
        mask = np.concatenate([np.full(nc, 1 << CHANNEL_NAMES.index(c), dtype=np.uint8) for c in channel])
        reps = len(channel)
        self.engine.add_spike_source(int(Nsource), np.asarray(spike_idcs, dtype=np.int32), np.asarray(spike_times, dtype=float),
                                     np.tile(pairs[1], reps), np.tile(pairs[0], reps), mask, np.tile(g, reps), np.tile(pf, reps))
        self.external_stimuli[name] = NetworkHolder(spike_idcs=np.asarray(spike_idcs), spike_times=np.asarray(spike_times), pairs=pairs)

    def _random_pairs(self, Nsource, target_idc, pCon):
        target_idc = np.asarray(target_idc)
        n_conn = int(round(pCon * len(target_idc) * Nsource, 0))
        local = np.arange(Nsource * len(target_idc))
        np.random.shuffle(local)
        local = local[:n_conn]
        return np.stack([target_idc[local // Nsource], np.arange(Nsource)[local % Nsource]]).astype(int)

    def set_poisson_stimuli(self, stimulator_name, Nsource, channel, target_idc, pCon, rate, start, stop, gmax, pfail):
        times, idcs = [], []
        for src in range(Nsource):
            lambd = 1000 / rate
            upper = int(round(2 * (stop - start) / lambd, 0))
            # np.clip(x, dt, NaN) under the reference's numpy: lower bound only (see codes/CortexNetwork.py port)
            arr = start + np.cumsum(np.maximum(-lambd * np.log(1 - np.random.rand(upper)), self.dt_ms))
            arr = arr[arr < stop]
            times.extend(arr)
            idcs.extend([src] * len(arr))
        pairs = self._random_pairs(Nsource, target_idc, pCon)
        self.set_custom_stimuli(stimulator_name, Nsource, channel, np.asarray(idcs), times, pairs, gmax, pfail)

    def set_regular_stimuli(self, stimulator_name, Nsource, channel, target_idc, pCon, rate, start, stop, gmax, pfail):
        n_spikes = round((stop - start) * rate / 1000)
        if 1000 / rate < self.dt_ms:
            raise ValueError("ERROR: time step is larger than spikes intervals")
        times = list(np.linspace(start, stop, n_spikes, endpoint=False)) * Nsource
        idcs = np.concatenate([[src] * n_spikes for src in range(Nsource)]).astype(int)
        pairs = self._random_pairs(Nsource, target_idc, pCon)
        self.set_custom_stimuli(stimulator_name, Nsource, channel, idcs, times, pairs, gmax, pfail)

    # ---- helpers (:608-700) --------------------------------------------------------------------------------------------------
    def group_idcs(self, group):
        return self.network.group_idcs(group)

    def neuron_idcs(self, groupstripe_list):
        return self.network.neuron_idcs(groupstripe_list)

    def syn_idcs_from_neurons(self, target, source, channel=None):
        return self.network.syn_idcs_from_neurons(target, source, channel)

    def syn_idcs_from_groups(self, t, s, channel=None):
        return self.network.syn_idcs_from_groups(t, s, channel)

    def rheobase(self, neuron_idcs=None):
        return self.network.rheobase(neuron_idcs)

    def spiking_count(self, neuron_idcs=None, tlim=None, delta_t=None):
        if tlim is None and delta_t is None:
            count = self.spikemonitor.count
            return count if neuron_idcs is None else count[neuron_idcs]
        t0, t1 = (self.transient, self.neurons.t / units.ms) if tlim is None else tlim
        if neuron_idcs is None:
            neuron_idcs = np.arange(self.neurons.N)
        trains = self.spikemonitor.spike_trains()
        if delta_t is None:
            return np.asarray([np.count_nonzero((trains[i] / units.ms >= t0) & (trains[i] / units.ms < t1)) for i in neuron_idcs])
        n_int = int(np.ceil((t1 - t0) / delta_t))
        edges = t0 + delta_t * np.arange(n_int + 1)
        return np.asarray([np.histogram(trains[i] / units.ms, bins=edges)[0] for i in neuron_idcs])

    def spiking_rate(self, neuron_idcs=None, tlim=None):
        count = self.spiking_count(neuron_idcs, tlim)
        if tlim is None:
            return count / (self.neurons.t / units.second - self.transient / 1000)
        return count / ((tlim[1] - tlim[0]) / 1000)

    def spiking_idcs(self, comparisons, groupstripe=None, tlim=None):
        if isinstance(comparisons[0], np.ufunc):
            comparisons = [comparisons]
        rate = self.spiking_rate(tlim=tlim)
        keep = np.ones(self.neurons.N, dtype=bool)
        for relation, value in comparisons:
            keep &= relation(rate, value)
        if groupstripe is not None:
            keep &= np.isin(np.arange(self.neurons.N), self.neuron_idcs(groupstripe))
        return np.arange(self.neurons.N)[keep]
