"""`SingleNeuron` of codes/SingleNeuron.py:12-107 on the B200 engine: one simplified-AdEx neuron driven by a constant current, with
a spike monitor and a state monitor of V and w at every step.

    neuron = SingleNeuron(params, I=56)          # params: the 12 NeuPar rows of one cell (pF nS mV mV mV ms nS pA mV mV pA mV)
    neuron.run(5000 * ms)
    neuron.spikemonitor.t / ms,  neuron.neuronmonitor.V[0] / mV,  neuron.neuronmonitor.w[0] / pA,  neuron.I_SN

Two things differ from the network's neuron model and are reproduced: the drive is the model variable `I` (no synaptic currents),
and the reset statement itself writes `last_spike = t` (:52) — in the network only the synaptic pathway p5 does — so the
depolarisation-block term can switch on after a spike when I >= I_ref; the engine gets that by marking the neuron as one that has
the pathway (`set_has_outgoing`).  The plotting methods of the reference (matplotlib) are replaced by `nullclines`, which returns
the curves `phase_portrait` draws.
"""
import numpy as np

from .. import units
from ..engine import Engine
from ..monitors import SpikeMonitorView, StateMonitorView


class SingleNeuron:
    def __init__(self, params, I=0, V0=False, w0=0, method="rk4", time_step=0.05, device=0, _engine_cls=None):
        self.params = params
        self.I = I
        self.time_step = time_step
        p = [float(x) for x in params]
        V_init = float(V0) if type(V0) in (int, float) else p[2]                 # E_L unless a number is given (:60-63)
        w_init = float(w0) if type(w0) in (int, float) else 0.0
        self.engine = (_engine_cls or Engine)(1, float(time_step), method={"gsl_rk2": "rk23"}.get(method, method), seed=0,
                                              **({} if _engine_cls else {"device": device}))
        rows = np.array([[p[0]], [p[1]], [p[2]], [p[3]], [p[4]], [p[5]], [p[7]], [p[8]], [p[9]], [p[10]], [p[11]], [float(I)]])
        self.engine.set_neurons(rows, [V_init], [w_init])
        # no synapses: every channel conductance stays 0, so I_tot = I_DC = I
        self.engine.set_channels(tau_on=[1.4, 3.0, 4.3], tau_off=[10.0, 40.0, 75.0], E_rev=[0.0, -70.0, 0.0], gsyn=[1.0, 1.0, 1.0],
                                 mg_fac=0.33, mg_slope=0.0625, mg_half=0.0)
        self.engine.set_has_outgoing(np.ones(1, dtype=np.uint8))                  # `last_spike = t` in the reset statement (:52)
        self.I_SN = p[1] * (p[9] - p[2] - p[3])                                   # rheobase g_L (V_T - E_L - delta_T), pA (:93)
        self.spikemonitor = SpikeMonitorView(self.engine, 1)
        self.neuronmonitor = StateMonitorView(self.engine, ["V", "w"], True)

    def run(self, t):
        """t in seconds (write `5000 * ms` with units.ms, as the reference does with Brian's)."""
        self.engine.run(int(round(float(t) / (self.time_step * units.ms))))

    def nullclines(self, I=False, V_range=False, n=400):
        """-> V, w_V(V) (the V-nullcline, where dV/dt = 0) and the two borders w_V -/+ D0/tau_w of the band inside which w moves
        along it — what `phase_portrait` (:123-180) draws; V in mV, currents in pA."""
        C, gL, EL, dT, Vup, tauw, _, _, _, VT = [float(x) for x in self.params[:10]]
        I = self.I if type(I) is bool else I
        lo, hi = (EL - 20.0, Vup) if type(V_range) is bool else V_range
        V = np.linspace(lo, hi, n)
        w_V = I + gL * dT * np.exp((V - VT) / dT) - gL * (V - EL)
        D0 = (C / gL) * w_V
        return V, w_V, w_V - D0 / tauw, w_V + D0 / tauw

    def close(self):
        self.engine.close()
