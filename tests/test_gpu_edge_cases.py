"""Edge cases of the CUDA engine through the C ABI: ragged sizes, empty inputs, zero-delay and long-delay synapses,
dt = 0.01 ms, error behaviour.  Same bit-parity bar as test_gpu_parity.py."""
import numpy as np
import pytest

from helpers import make_engine
from oracle_binding import OracleEngine, default_idc, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine, HhdError

pytestmark = pytest.mark.gpu

BASE = dict(C=200.0, g_L=10.0, E_L=-70.0, delta_T=2.0, V_up=-40.0, tau_w=100.0, b=50.0, V_r=-60.0, V_T=-50.0,
            I_ref=1e9, V_dep=-60.0, I_DC=0.0)
CH = dict(tau_on=[1.4, 3.0, 4.3], tau_off=[10.0, 40.0, 75.0], E_rev=[0.0, -70.0, 0.0], gsyn=[1.6, 3.2, 5.0],
          mg_fac=0.33, mg_slope=0.0625, mg_half=0.0)
STATE = ["V", "w", "g_AMPA_on", "g_AMPA_off", "g_GABA_on", "g_GABA_off", "g_NMDA_on", "g_NMDA_off"]


def small(cls, n, idc, dt=0.05, **kw):
    e = cls(n, dt, **kw)
    p = {k: np.full(n, v) for k, v in BASE.items()}
    rng = np.random.default_rng(1)
    p["C"] = 150 + 100 * rng.random(n)
    p["I_DC"] = np.asarray(idc, dtype=float)
    e.set_neurons(p, np.full(n, -70.0), np.zeros(n))
    e.set_channels(**CH)
    return e


def random_synapses(n, nsyn, rng, max_delay_ms):
    return dict(src=rng.integers(0, n, nsyn), tgt=rng.integers(0, n, nsyn), chan=rng.integers(0, 3, nsyn),
                g_max=rng.random(nsyn) * 3, U=0.1 + 0.5 * rng.random(nsyn), tau_rec=50 + 500 * rng.random(nsyn),
                tau_fac=10 + 300 * rng.random(nsyn), p_fail=np.full(nsyn, 0.3), delay_ms=rng.random(nsyn) * max_delay_ms)


def same(gpu, cpu, syn=True):
    gi, gs = gpu.spikes()
    ci, cs = cpu.spikes()
    assert np.array_equal(gi, ci) and np.array_equal(gs, cs)
    for v in STATE + ["last_spike"]:      # equal_nan: with I_ref = 1e9 (no depolarisation block) strongly driven cells run away
        assert np.array_equal(gpu.read_state(v), cpu.read_state(v), equal_nan=True), v     # to NaN — in both
    if syn:
        for v in ("u", "R", "a"):
            assert np.array_equal(gpu.read_syn_state(v), cpu.read_syn_state(v), equal_nan=True), v
    assert gpu.counters()["nonfinite"] == cpu.counters()["nonfinite"]
    return len(ci)


@pytest.mark.parametrize("plan", ["graph", "resident"])
@pytest.mark.parametrize("n", [1, 31, 129, 300])
def test_ragged_sizes_with_random_synapses(n, plan):
    rng = np.random.default_rng(n)
    syn = random_synapses(n, 40 * n, rng, 3.0)
    syn["delay_ms"][: max(1, n // 2)] = 0.0                     # zero-delay synapses
    idc = 400 + 400 * rng.random(n)
    gpu = small(Engine, n, idc, method="rk2", accum="fixed", seed=5, plan=plan)
    cpu = small(OracleEngine, n, idc, method="rk2", accum="fixed", seed=5)
    for e in (gpu, cpu):
        e.set_synapses(**syn)
        e.run(4000)
    assert same(gpu, cpu) >= n // 2    # most cells fired (some run away to NaN without the depolarisation block)
    assert gpu.counters()["syn_deliveries"] == cpu.counters()["syn_deliveries"] > 0


@pytest.mark.parametrize("plan", ["graph", "resident"])
def test_no_synapses_and_silent_network(plan):
    gpu = small(Engine, 64, np.zeros(64), plan=plan)
    cpu = small(OracleEngine, 64, np.zeros(64))
    for e in (gpu, cpu):
        e.run(500)
    assert same(gpu, cpu, syn=False) == 0
    assert gpu.counters()["spikes"] == 0 and gpu.spike_count() == 0


def test_delays_longer_than_the_refractory_period():
    """Non-fused path: a second spike changes the amplitude of a delivery still in flight."""
    n = 200
    rng = np.random.default_rng(3)
    syn = random_synapses(n, 6000, rng, 14.0)                  # up to 280 steps > 100 refractory steps
    idc = 500 + 900 * rng.random(n)                            # fast cells: ISI shorter than the longest delays
    gpu = small(Engine, n, idc, method="euler", accum="fixed", seed=8, plan="graph")
    cpu = small(OracleEngine, n, idc, method="euler", accum="fixed", seed=8)
    for e in (gpu, cpu):
        e.set_synapses(**syn)
        e.run(6000)
    assert same(gpu, cpu) > 1000
    with pytest.raises(HhdError):                              # the resident plan must refuse this network
        bad = small(Engine, n, idc, plan="resident")
        bad.set_synapses(**syn)
        bad.run(1)


@pytest.mark.parametrize("plan", ["graph", "resident"])
def test_dt_001_as_the_stimulation_scripts_use(plan):
    """codes/Main_regular.py:30: dt = 0.01 ms -> 500 refractory steps, delays up to ~355 steps."""
    g = load_golden("net_n200_s0")
    gpu = make_engine("gpu", g, dt=0.01, method="rk4", accum="fixed", seed=2, plan=plan)
    cpu = make_engine("cpu", g, dt=0.01, method="rk4", accum="fixed", seed=2)
    assert cpu.read_syn_state("delay_steps").max() > 300
    for e in (gpu, cpu):
        e.run(8000)
    assert same(gpu, cpu) > 100


def test_monitor_buffers_grow_and_many_monitors():
    g = load_golden("net_n200_s0")
    res = {}
    for kind in ("gpu", "cpu"):
        e = make_engine(kind, g, method="euler", accum="fixed", seed=1, **({"plan": "graph"} if kind == "gpu" else {}))
        mons = [e.add_state_monitor(v, np.arange(0, 207, 7, dtype=np.int32)) for v in ("V", "w", "I_syn", "g_GABA", "I_exp", "last_spike")]
        mons += [e.add_state_monitor("I_tot", None, reduce=r) for r in ("sum", "mean")]
        for n in (10, 700, 1, 2400):                           # capacity is grown between runs
            e.run(n)
        res[kind] = [e.read_monitor(m) for m in mons]
    for a, b in zip(res["gpu"][:6], res["cpu"][:6]):
        assert a.shape == (3111, 30) and np.array_equal(a, b)
    for a, b in zip(res["gpu"][6:], res["cpu"][6:]):
        assert a.shape == (3111, 1) and np.allclose(a, b, rtol=1e-12, atol=1e-9)


def test_error_behaviour():
    e = Engine(10, 0.05)
    with pytest.raises(HhdError) as ei:
        e.run(1)
    assert ei.value.code == -5 and "set_neurons" in str(ei.value)
    e = small(Engine, 10, np.zeros(10))
    ok = dict(src=[0], tgt=[1], chan=[0], g_max=[1.0], U=[0.2], tau_rec=[100.0], tau_fac=[100.0], p_fail=[0.0], delay_ms=[1.0])
    for key, val, frag in (("src", [10], "source"), ("tgt", [-1], "target"), ("chan", [3], "channel"), ("delay_ms", [-1.0], "delay"),
                           ("tau_rec", [0.0], "tau"), ("delay_ms", [1e6], "65535")):
        bad = dict(ok)
        bad[key] = val
        with pytest.raises(HhdError) as ei:
            e.set_synapses(**bad)
        assert ei.value.code == -1 and frag in str(ei.value)
    e.set_synapses(**ok)
    with pytest.raises(HhdError):
        e.set_synapses(**ok)                                   # only once
    e.run(5)
    with pytest.raises(HhdError):
        e.add_spike_source(1, [0], [0.0], [0], [0], [1], [1.0], [0.0])   # spike in the past
    with pytest.raises(HhdError):
        e.add_state_monitor("V", [0, 0])                       # duplicate index
    with pytest.raises(HhdError):
        Engine(10, 0.05, method=7)
    with pytest.raises(HhdError):
        Engine(0, 0.05)
    with pytest.raises(HhdError):                              # instances must not be connected
        b = small(Engine, 20, np.zeros(20), plan="resident", instance_size=10)
        b.set_synapses(**dict(ok, src=[0], tgt=[15]))
        b.run(1)


def test_ring_of_columns_with_inter_column_synapses():
    """scaled.multicolumn_synapses_device (the bench's `multicolumn` workload) at small scale: 7 columns of the 207-neuron
    column, inter-column delays up to ~15 ms (3x the refractory period) -> the non-fused path, against the oracle."""
    from oracle_binding import golden_groups
    from rescience_hass_hertag_durstewitz_2016_b200.engine import channels_from_STypPar, neuron_rows_from_NeuPar
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import multicolumn_synapses_device, tile_columns
    g = load_golden("net_n200_s0")
    K, n1 = 7, 207
    groups = [grp[0] for grp in golden_groups(g)]
    syn = [x.numpy() for x in multicolumn_synapses_device(g["SynPar"], g["source_arr"], g["target_arr"], groups, K, n1, "cpu", seed=3)]
    assert len(syn[0]) > K * g["SynPar"].shape[1] and (syn[8] / 0.05).max() > 200
    net = tile_columns(g["NeuPar"], g["V0"], np.zeros((7, 0)), [], [], default_idc(g), K)
    out = {}
    for name, cls, kw in (("gpu", Engine, dict(plan="graph")), ("cpu", OracleEngine, {})):
        e = cls(K * n1, 0.05, method="rk2", accum="fixed", seed=6, **kw)
        e.set_neurons(neuron_rows_from_NeuPar(net["NeuPar"], net["I_DC"]), net["V0"][0], net["V0"][1])
        e.set_channels(**channels_from_STypPar(g["STypPar"]))
        e.set_synapses(*syn)
        e.run(8000)
        out[name] = e
    assert same(out["gpu"], out["cpu"]) > 2000
    assert out["gpu"].counters()["syn_deliveries"] == out["cpu"].counters()["syn_deliveries"]


@pytest.mark.parametrize("plan", ["graph", "auto"])
@pytest.mark.parametrize("method", ["euler", "rk4", "rk23"])
def test_timed_current(method, plan):
    """TimedArray I_AC: stages at t + dt read the next row; bit-identical to the oracle (the resident plan steps aside)."""
    g = load_golden("net_n200_s0")
    steps = 3000
    tt = np.arange(steps) * 0.05
    table = np.stack([150.0 * np.sin(2 * np.pi * tt / 37.0) + 100.0, np.where((tt > 20) & (tt < 60), 400.0, 0.0)], axis=1)
    col = np.full(207, -1, dtype=np.int32)
    col[:94] = 0
    col[150:] = 1
    res = {}
    for kind in ("gpu", "cpu"):
        e = make_engine(kind, g, method=method, accum="fixed", seed=12, **({"plan": plan} if kind == "gpu" else {}))
        e.set_timed_current(table, col)
        m = e.add_state_monitor("I_tot", np.arange(0, 207, 11, dtype=np.int32))
        e.run(steps + 500)
        res[kind] = (e, e.read_monitor(m))
    assert same(res["gpu"][0], res["cpu"][0]) > 200
    assert np.array_equal(res["gpu"][1], res["cpu"][1])


@pytest.mark.parametrize("plan", ["graph", "auto"])
def test_conductance_noise_bit_exact(plan):
    """`noise` variant: Euler-Maruyama term from Philox normals (hhd_normal2: polar method on hhd_log) — the same numbers on
    both sides, so the whole network run is bit-identical; the resident plan steps aside; other methods are refused."""
    g = load_golden("net_n200_s0")
    n = g["NeuPar"].shape[1]
    ncon = np.bincount(g["target_arr"], minlength=n) + np.bincount(g["source_arr"], minlength=n)
    sigma = np.sqrt(2 * 0.002 * ncon.astype(np.float64))
    res = {}
    for kind in ("gpu", "cpu"):
        e = make_engine(kind, g, method="euler", accum="fixed", seed=21, **({"plan": plan} if kind == "gpu" else {}))
        e.set_noise(sigma)
        m = e.add_state_monitor("g_NMDA_off", np.arange(0, n, 13, dtype=np.int32))
        e.run(4000)
        res[kind] = (e, e.read_monitor(m))
    assert same(res["gpu"][0], res["cpu"][0]) > 100
    assert np.array_equal(res["gpu"][1], res["cpu"][1])
    assert res["gpu"][1].std() > 0
    e = make_engine("gpu", g, method="rk4")
    with pytest.raises(HhdError):
        e.set_noise(sigma)
