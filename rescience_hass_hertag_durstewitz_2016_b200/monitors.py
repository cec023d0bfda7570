"""Monitor objects with the attribute surface the reference's analysis code reads from Brian 2 monitors
(SURVEY.md §8b "Python-side attribute contract"): SpikeMonitor -> .i .t .t_ .count .num_spikes .spike_trains();
StateMonitor -> .<var> as [n_recorded, n_samples], .t, .record."""
import numpy as np

from . import units


class SpikeMonitorView:
    """What `SpikeMonitor(group)` exposes (codes/CortexNetwork.py:487; consumers :2727-2765, 4300-4301;
    codes_revision/AnalysesTools.py:189-194, 394).  Values are fetched from the engine on access."""

    def __init__(self, engine, n_neurons, id_offset=0):
        self._e = engine
        self._n = n_neurons
        self._off = id_offset
        self._cache = None
        self._cache_step = -1

    def _fetch(self):
        step = (self._e.current_step(), getattr(self._e, "_spike_epoch", 0))     # cleared or toggled monitors invalidate too
        if self._cache is None or self._cache_step != step:
            i, s = self._e.spikes()
            self._cache = (i.astype(np.int64) - self._off, s.astype(np.float64) * self._e.dt_ms * units.ms)
            self._cache_step = step
        return self._cache

    @property
    def i(self):
        return self._fetch()[0]

    @property
    def t(self):
        return self._fetch()[1]

    t_ = t

    @property
    def num_spikes(self):
        return len(self._fetch()[0])

    @property
    def count(self):
        return np.bincount(self._fetch()[0], minlength=self._n).astype(np.int32)

    def spike_trains(self):
        i, t = self._fetch()
        order = np.argsort(i, kind="stable")
        i_s, t_s = i[order], t[order]
        bounds = np.searchsorted(i_s, np.arange(self._n + 1))
        return {n: t_s[bounds[n]:bounds[n + 1]] for n in range(self._n)}

    def all_values(self):
        return {"t": self.spike_trains()}


class StateMonitorView:
    """What `StateMonitor(group, variables, record)` exposes (codes/CortexNetwork.py:510; revision CortexSetup.py:532):
    `mon.V` is [n_recorded, n_samples] in SI, `mon.t` the sample times, `mon.record` the recorded neuron indices."""

    def __init__(self, engine, variables, record, reduce="none"):
        self._e = engine
        self.variables = [variables] if isinstance(variables, str) else list(variables)
        n = engine.n_neurons
        if record is True or record is None:
            self.record = np.arange(n)
            idx = None
        else:
            self.record = np.unique(np.asarray(record, dtype=np.int64))
            idx = self.record.astype(np.int32)
        self._ids = {v: engine.add_state_monitor(v, idx, reduce=reduce) for v in self.variables}

    def _samples(self):
        return self._e.monitor_samples(self._ids[self.variables[0]])

    @property
    def t(self):
        n, first = self._samples()
        steps = self._e.monitor_steps(self._ids[self.variables[0]]) if hasattr(self._e, "monitor_steps") else np.zeros(0)
        if len(steps) != n:                         # monitor capacity reached, or an engine that does not track runs
            steps = max(first, 0) + np.arange(n)
        return steps * self._e.dt_ms * units.ms

    t_ = t

    @property
    def N(self):
        return self._samples()[0]

    def __getattr__(self, name):
        ids = self.__dict__.get("_ids", {})
        base = name[:-1] if name.endswith("_") and name[:-1] in ids else name
        if base in ids:
            return self._e.read_monitor(ids[base]).T * units.SI_SCALE.get(base, 1.0)
        raise AttributeError(name)

    def set_active(self, active):
        for m in self._ids.values():
            self._e.set_monitor_active(m, active)
