"""Known-answer tests that pin the CPU oracle's reading of the reference model (it has no golden trajectories):
hand-computed Tsodyks–Markram sequences, delay / refractory bookkeeping, monitor timing, and the documented
`last_spike` quirk.  Plain numpy / math on the right-hand side, no shared code with the oracle."""
import math

import numpy as np
import pytest

from oracle_binding import OracleEngine, load_golden

# a regular-spiking cell: C=200 pF, g_L=10 nS, E_L=-70, delta_T=2, V_up=-40, tau_w=100, b=50, V_r=-60, V_T=-50
BASE = dict(C=200.0, g_L=10.0, E_L=-70.0, delta_T=2.0, V_up=-40.0, tau_w=100.0, b=50.0, V_r=-60.0, V_T=-50.0,
            I_ref=1e9, V_dep=-60.0, I_DC=0.0)
CH = dict(tau_on=[1.4, 3.0, 4.3], tau_off=[10.0, 40.0, 75.0], E_rev=[0.0, -70.0, 0.0], gsyn=[1.0, 1.0, 1.0],
          mg_fac=0.33, mg_slope=0.0625, mg_half=0.0)


def engine(n, idc, dt=0.05, method="euler", **kw):
    e = OracleEngine(n, dt, method=method, **kw)
    p = {k: np.full(n, v) for k, v in BASE.items()}
    p["I_DC"] = np.asarray(idc, dtype=float)
    e.set_neurons(p, np.full(n, -70.0), np.zeros(n))
    e.set_channels(**CH)
    return e


def test_subthreshold_relaxation_matches_closed_form():
    """Far below V_T and with w pinned by the w-nullcline projection out of reach, V relaxes like a leaky integrator;
    rk4 at dt=0.05 must match exp(-t/tau) to 1e-9."""
    e = engine(1, [50.0], method="rk4")
    e.run(400)
    V = e.read_state("V")[0]
    tau = BASE["C"] / BASE["g_L"]
    # dV/dt = (I + I_exp - g_L (V-E_L) - w)/C with I_exp ~ 1e-4 pA and w = 0 (dw gate closed: w not inside the band)
    Vinf = BASE["E_L"] + 50.0 / BASE["g_L"]
    expect = Vinf + (-70.0 - Vinf) * math.exp(-400 * 0.05 / tau)
    assert abs(V - expect) < 5e-3       # residual: I_exp and the adaptation current that builds up
    assert e.counters()["spikes"] == 0


def test_refractory_is_100_steps_and_gates_only_the_threshold():
    """NeuronGroup(refractory=5*ms) at dt = 0.05: a cell pushed above V_up 99 steps after its spike must wait for step
    +100; V keeps integrating meanwhile (no `unless refractory` flag in the equations)."""
    e = engine(1, [700.0])
    e.run(3000)
    k = int(e.spikes()[1][0])
    e = engine(1, [700.0])
    e.run(k + 1)                                    # the spike of step k has been reset
    assert e.read_state("V")[0] == BASE["V_r"]
    e.run(98)
    assert e.current_step() == k + 99
    e.write_state("V", [-39.9])                     # above V_up = -40
    e.run(1)                                        # step k+99: over threshold but refractory
    assert e.read_state("V")[0] > -39.9 and list(e.spikes()[1]) == [k]
    e.run(1)                                        # step k+100: allowed again
    assert list(e.spikes()[1]) == [k, k + 100]


def two_cell(delay_ms, U=0.25, tau_rec=500.0, tau_fac=200.0, pfail=0.0, gmax=2.0, chan=0, dt=0.05, idc0=700.0, **kw):
    e = engine(2, [idc0, 0.0], dt=dt, **kw)
    e.set_synapses([0], [1], [chan], [gmax], [U], [tau_rec], [tau_fac], [pfail], [delay_ms])
    return e


def test_stp_sequence_matches_hand_recursion():
    U, trec, tfac = 0.25, 500.0, 200.0
    e = two_cell(1.0, U, trec, tfac)
    e.run(6000)
    _, steps = e.spikes()
    assert len(steps) >= 5
    u, R, last = U, 1.0, -1e8
    for k in steps:
        t = k * 0.05
        u1 = U + u * (1 - U) * math.exp(-(t - last) / tfac)
        R1 = 1 + (R - u * R - 1) * math.exp(-(t - last) / trec)
        u, R, last = u1, R1, t
    assert e.read_syn_state("u")[0] == pytest.approx(u, rel=1e-12)
    assert e.read_syn_state("R")[0] == pytest.approx(R, rel=1e-12)
    assert e.read_syn_state("a")[0] == pytest.approx(2.0 * u * R, rel=1e-12)
    # first spike: u1 = U exactly (exp(-1e8/tau) = 0), R1 = 1 + (1 - U - 1)*0 = 1
    # facilitation makes u grow with every spike at this rate
    assert u > U


def test_delay_bookkeeping_first_effect_one_step_after_delivery():
    """Spike at step k, delay d steps -> increment lands in the `synapses` slot of step k+d, i.e. the StateMonitor
    (slot `start`) first sees it at step k+d+1."""
    for delay_ms, d in ((0.0, 0), (0.02, 0), (0.03, 1), (1.0, 20), (3.55, 71)):
        e = two_cell(delay_ms)
        m = e.add_state_monitor("g_AMPA_off", [1])
        e.run(1500)
        _, steps = e.spikes()
        g = e.read_monitor(m)[:, 0]
        k = int(steps[0])
        assert e.read_syn_state("delay_steps")[0] == d
        assert np.all(g[:k + d + 1] == 0.0)
        assert g[k + d + 1] == pytest.approx(2.0 * 0.25, rel=1e-12)       # gmax * U * R(=1) * Gsyn(=1)


def test_release_failure_rate_and_amplitude():
    """p0: failure = int(rand() < failure_rate) per (spike, synapse); a failed release delivers exactly 0."""
    n = 300
    e = engine(n + 1, [700.0 + i for i in range(n)] + [0.0], seed=123)
    e.set_synapses(np.arange(n), np.full(n, n), np.zeros(n), np.full(n, 2.0), np.full(n, 0.25), np.full(n, 500.0),
                   np.full(n, 200.0), np.full(n, 0.3), np.full(n, 1.0))
    seen, events, failed = 0, 0, 0
    for _ in range(150):
        e.run(40)
        i, _ = e.spikes()
        new = i[seen:]
        new = new[new < n]                               # the target itself may fire
        seen = len(i)
        a = e.read_syn_state("a")
        events += len(new)
        failed += int(np.sum(a[new] == 0.0))
        assert np.all(a[new] >= 0.0)
    assert events > 800
    assert abs(failed / events - 0.3) < 0.05


def test_last_spike_is_written_by_the_pathway_not_by_the_reset():
    """codes/CortexNetwork.py:365 `last_spike_pre = t`: only neurons with outgoing synapses get last_spike (F6)."""
    e = engine(2, [700.0, 700.0])
    e.set_synapses([0], [1], [0], [0.0], [0.25], [500.0], [200.0], [0.0], [1.0])
    e.run(3000)
    i, steps = e.spikes()
    ls = e.read_state("last_spike")
    assert ls[0] == pytest.approx(steps[i == 0][-1] * 0.05)
    assert ls[1] == -1e8 and np.any(i == 1)


def test_channels_and_nmda_mg_block():
    """One presynaptic spike into each channel; currents follow I = g (E - V) [* Mg for NMDA]."""
    for chan, E in ((0, 0.0), (1, -70.0), (2, 0.0)):
        e = two_cell(0.0, chan=chan)
        mg = e.add_state_monitor(["g_AMPA", "g_GABA", "g_NMDA"][chan], [1])
        mi = e.add_state_monitor(["I_AMPA", "I_GABA", "I_NMDA"][chan], [1])
        mv = e.add_state_monitor("V", [1])
        e.run(1200)
        g, I, V = (e.read_monitor(x)[:, 0] for x in (mg, mi, mv))
        k = int(np.argmax(g > 0)) + 40
        expect = g[k] * (E - V[k])
        if chan == 2:
            expect *= 1.0 / (1.0 + 0.33 * math.exp(0.0625 * (0.0 - V[k])))
        assert g[k] > 0
        assert I[k] == pytest.approx(expect, rel=1e-12, abs=1e-12)


def test_lfp_monitor_is_the_sum_of_I_tot():
    g = load_golden("net_n200_s0")
    from helpers import make_engine
    e = make_engine("cpu", g, method="rk2")
    ms = e.add_state_monitor("I_tot", None, reduce="sum")
    ma = e.add_state_monitor("I_tot", None)
    mm = e.add_state_monitor("I_tot", None, reduce="mean")
    e.run(300)
    s, a, mmean = e.read_monitor(ms), e.read_monitor(ma), e.read_monitor(mm)
    assert np.allclose(s[:, 0], a.sum(axis=1), rtol=1e-12)
    assert np.allclose(mmean[:, 0], a.mean(axis=1), rtol=1e-12)


def test_depolarisation_block_term():
    """I_tot >= I_ref within 5 ms of the last spike: dV = -(g_L/C)(V - V_dep) (codes/CortexNetwork.py:251)."""
    e = OracleEngine(2, 0.05, method="euler")
    p = {k: np.full(2, v) for k, v in BASE.items()}
    p["I_DC"] = np.array([3000.0, 0.0])
    p["I_ref"] = np.array([2000.0, 1e9])
    e.set_neurons(p, np.full(2, -70.0), np.zeros(2))
    e.set_channels(**CH)
    e.set_synapses([0], [1], [0], [0.0], [0.25], [500.0], [200.0], [0.0], [1.0])   # gives neuron 0 a last_spike
    m = e.add_state_monitor("V", [0])
    e.run(2000)
    _, steps = e.spikes()
    V = e.read_monitor(m)[:, 0]
    k = int(steps[0])
    # after the reset V = V_r = V_dep, so the block term holds it there exactly for the 5 ms window
    assert np.all(V[k + 1:k + 100] == BASE["V_r"])
    assert V[k + 105] > BASE["V_r"]


def test_golden_fixture_digests():
    """The fixtures are the reference's own NetworkSetup output; their digests pin them (tests/golden/make_golden_network.py)."""
    import hashlib
    for name in ("net_n200_s0", "net_n60x5_s1"):
        g = load_golden(name)
        for key, arr in (("sha_NeuPar", g["NeuPar"]), ("sha_V0", g["V0"]), ("sha_SynPar", g["SynPar"]),
                         ("sha_target", g["target_arr"].astype(np.int64)), ("sha_source", g["source_arr"].astype(np.int64))):
            assert hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest() == str(g[key]), (name, key)


def test_timed_current_follows_timedarray_semantics():
    """I_AC = ta(t, i) (codes/CortexNetwork.py:133-151): row s during step s, the last row held afterwards, neurons
    without a column get 0; a sinusoidal current makes a passive cell oscillate with the analytic gain."""
    n, dt, steps = 3, 0.05, 4000
    e = engine(n, [0.0, 0.0, 100.0], dt=dt, method="rk4")
    tt = np.arange(steps) * dt
    table = np.stack([20.0 * np.sin(2 * np.pi * tt / 50.0), np.where(tt < 10.0, 0.0, 30.0)], axis=1)
    e.set_timed_current(table, [0, 1, -1])
    mi = e.add_state_monitor("I_inj", None)
    mv = e.add_state_monitor("V", [0])
    e.run(steps + 200)
    I = e.read_monitor(mi)
    assert np.array_equal(I[:steps, 0], table[:, 0]) and np.array_equal(I[:steps, 1], table[:, 1])
    assert np.all(I[steps:, 0] == table[-1, 0]) and np.all(I[steps:, 1] == 30.0)      # last row held
    assert np.all(I[:, 2] == 100.0)                                                    # I_DC only
    V = e.read_monitor(mv)[2000:steps, 0]                   # steady oscillation of a leaky integrator
    tau = BASE["C"] / BASE["g_L"]
    gain = 1.0 / (BASE["g_L"] * math.sqrt(1.0 + (2 * math.pi * tau / 50.0) ** 2))
    assert (V.max() - V.min()) / 2 == pytest.approx(20.0 * gain, rel=0.02)


def test_conductance_noise_is_an_ou_process():
    """`noise` variant (codes/CortexNetwork.py:182-189): g_X_off is an Ornstein-Uhlenbeck process, Euler-Maruyama as Brian's
    stochastic euler does it: stationary variance sigma^2 tau / 2 (discretisation: / (1 - dt / (2 tau))), zero mean, the three
    channels independent; the injected neuron is clamped far below threshold so nothing else moves the conductances."""
    n, dt, steps = 256, 0.05, 40000
    e = engine(n, np.zeros(n), dt=dt, method="euler", seed=5)
    sigma = np.full(n, 0.05)
    sigma[: n // 2] = 0.0
    e.set_noise(sigma)
    mons = [e.add_state_monitor(v, np.arange(n, dtype=np.int32)) for v in ("g_AMPA_off", "g_GABA_off", "g_NMDA_off", "g_AMPA_on")]
    e.run(steps)
    tau = {"g_AMPA_off": 10.0, "g_GABA_off": 40.0, "g_NMDA_off": 75.0}
    tr = {v: e.read_monitor(m) for v, m in zip(("g_AMPA_off", "g_GABA_off", "g_NMDA_off", "g_AMPA_on"), mons)}
    assert not tr["g_AMPA_on"].any()                         # only the _off equations carry noise
    for v, t in tau.items():
        x = tr[v][int(10 * t / dt):]                          # after ten time constants
        assert not x[:, : n // 2].any()                       # sigma = 0: untouched
        var = x[:, n // 2:].var()
        want = 0.05 ** 2 * t / 2 / (1 - dt / (2 * t))
        n_eff = (n // 2) * (x.shape[0] * dt) / (2 * t)         # independent samples: one per two time constants and neuron
        assert abs(var / want - 1) < 4 * np.sqrt(2 / n_eff), (v, var, want)
        assert abs(x[:, n // 2:].mean()) < 4 * np.sqrt(want / (x[:, n // 2:].size * dt / (2 * t)))
    a, b = tr["g_AMPA_off"][4000:, n // 2:], tr["g_GABA_off"][4000:, n // 2:]
    assert abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 0.02
