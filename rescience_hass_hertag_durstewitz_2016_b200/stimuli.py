"""External spike-train stimuli as plain arrays: what the reference builds for `SpikeGeneratorGroup` + input `Synapses`
(codes/CortexNetwork.py:525-622 Poisson, :624-719 regular), drawn with np.random in the reference's order.  Used by the
`CortexNetwork` facade and by the batched-instance builder of BASELINE config 3 (scaled.instances_with_stimuli).

Every function returns dict(n_sources, spk_src, spk_t, syn_src, syn_tgt, syn_type, syn_gmax, pfail):
  spk_src[k], spk_t[k]  spike k of source spk_src[k] at spk_t[k] ms
  syn_*                 input synapses source -> target neuron, type 1 (AMPA + NMDA) or 2 (GABA), g_max in nS
  pfail                 release-failure probability of ALL sets = the LAST set's value (`poissonsynapse.failure_rate =
                        pfailinp`, codes/CortexNetwork.py:613, 710)
"""
import numpy as np


def _targets(targets, group_distr, inp_group, n_inp, tgt_all, src_all):
    for group, stripe, pcon in targets:
        target_group = np.asarray(group_distr[group][stripe]).astype(np.int64)
        n_target = int(round(pcon * len(target_group), 0))
        order = np.arange(len(target_group) * n_inp)      # the reference shuffles a list: same draws, same permutation
        np.random.shuffle(order)
        pick = order[:n_target].astype(np.int64)
        tgt_all.extend(target_group[pick // n_inp])
        src_all.extend(inp_group[pick % n_inp])


def _pack(n_tot, spk_src, spk_t, syn_src, syn_tgt, syn_type, syn_gmax, pfail):
    return dict(n_sources=int(n_tot), spk_src=np.asarray(spk_src, dtype=np.int32), spk_t=np.asarray(spk_t, dtype=np.float64),
                syn_src=np.asarray(syn_src, dtype=np.int32), syn_tgt=np.asarray(syn_tgt, dtype=np.int32),
                syn_type=np.asarray(syn_type, dtype=np.int32), syn_gmax=np.asarray(syn_gmax, dtype=np.float64), pfail=float(pfail))


def draw_poisson_input(PoissonInput, group_distr, time_step):
    """[[N, type, rate_Hz, g_max_nS, p_fail, start_ms, stop_ms, [[group, stripe, pCon], ...]], ...].  Spike times:
    cumulative sums of exponential intervals clipped below at one time step (codes/CortexNetwork.py:541-549)."""
    spk_src, spk_t, syn_src, syn_tgt, syn_type, syn_gmax = [], [], [], [], [], []
    n_tot, pfail = 0, 0.0
    for n_inp, type_inp, freq, gmax, pfail, start, stop, targets in PoissonInput:
        src_all, tgt_all = [], []
        inp_group = np.arange(n_tot, n_tot + n_inp)
        for i in range(n_inp):
            lamb = 1000 / freq
            n_draw = int(round(2 * (stop - start) / lamb, 0))
            # np.clip(x, time_step, np.NaN) in the reference (:547): with the numpy it pins (1.16) a NaN bound is ignored,
            # i.e. intervals are only bounded below by one time step
            t = start + np.cumsum(np.maximum(-lamb * np.log(1 - np.random.rand(n_draw)), time_step))
            t = t[t < stop]
            spk_t.extend(t)
            spk_src.extend([n_tot + i] * len(t))
            _targets(targets, group_distr, inp_group, n_inp, tgt_all, src_all)
        syn_src.extend(src_all)
        syn_tgt.extend(tgt_all)
        syn_type.extend([type_inp] * len(src_all))
        syn_gmax.extend([gmax] * len(src_all))
        n_tot += n_inp
    return _pack(n_tot, spk_src, spk_t, syn_src, syn_tgt, syn_type, syn_gmax, pfail)


def draw_regular_input(RegularInput, group_distr, time_step):
    """[[N, type, n_spikes, g_max_nS, p_fail, start_ms, stop_ms, targets], ...]   (codes/CortexNetwork.py:624-719)"""
    spk_src, spk_t, syn_src, syn_tgt, syn_type, syn_gmax = [], [], [], [], [], []
    n_tot, pfail = 0, 0.0
    for n_inp, type_inp, n_spikes, gmax, pfail, start, stop, targets in RegularInput:
        if (stop - start) / (n_spikes - 1) < time_step:
            print("WARNING: time step is larger than spikes intervals")
        src_all, tgt_all = [], []
        inp_group = np.arange(n_tot, n_tot + n_inp)
        for i in range(n_inp):
            spk_t.extend(np.linspace(start, stop, n_spikes))
            spk_src.extend([n_tot + i] * n_spikes)
            _targets(targets, group_distr, inp_group, n_inp, tgt_all, src_all)
        syn_src.extend(src_all)
        syn_tgt.extend(tgt_all)
        syn_type.extend([type_inp] * len(src_all))
        syn_gmax.extend([gmax] * len(src_all))
        n_tot += n_inp
    return _pack(n_tot, spk_src, spk_t, syn_src, syn_tgt, syn_type, syn_gmax, pfail)


def channel_mask(syn_type):
    """type 1 feeds AMPA and NMDA, type 2 GABA (codes/CortexNetwork.py:609-611) -> the engine's channel bit mask."""
    return np.where(np.asarray(syn_type) == 1, 0b101, 0b010).astype(np.uint8)


def check_one_spike_per_step(spk_src, spk_t, dt):
    """SpikeGeneratorGroup bins a spike to int((t + 1e-3 dt) / dt) and refuses two spikes of one source in one bin."""
    steps = np.floor((np.asarray(spk_t, dtype=np.float64) + 1e-3 * dt) / dt).astype(np.int64)
    if len(steps):
        pairs = np.stack([np.asarray(spk_src, dtype=np.int64), steps], axis=1)
        if len(np.unique(pairs, axis=0)) != len(pairs):
            raise ValueError("a source spikes more than once during one time step (SpikeGeneratorGroup refuses that as well)")


def attach(engine, s, dt, source_offset=0):
    """Hand one drawn stimulus set to an engine (hhd_add_spike_source)."""
    check_one_spike_per_step(s["spk_src"], s["spk_t"], dt)
    engine.add_spike_source(s["n_sources"], s["spk_src"], s["spk_t"], s["syn_src"], s["syn_tgt"], channel_mask(s["syn_type"]),
                            s["syn_gmax"], np.full(len(s["syn_src"]), s["pfail"]))
