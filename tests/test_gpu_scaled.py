"""The bench's network path on the GPU against the CPU oracle: distinct columns built by scaled.build_distinct_columns (each its
own draw of the reference's generator), inter-column synapses with delays beyond the refractory period, handed over as device
tensors (`set_synapses_dev`) — spike lists, membrane potentials, STP state and event counts bit-identical (accum = fixed)."""
import numpy as np
import pytest
import torch

from oracle_binding import OracleEngine
from rescience_hass_hertag_durstewitz_2016_b200 import scaled
from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine, channels_from_STypPar, neuron_rows_from_NeuPar
from oracle_binding import load_golden

pytestmark = pytest.mark.gpu


def test_ring_of_distinct_columns_bit_exact(tmp_path):
    ncol, steps = 5, 6000
    cols = scaled.build_distinct_columns(200, 0, ncol, seed=7, workers=2, cache_dir=str(tmp_path))
    net = scaled.distinct_columns_network(cols)
    n = net["NeuPar"].shape[1]
    STypPar = load_golden("net_n200_s0")["STypPar"]
    host = scaled.distinct_synapses_device(cols, 0, ncol, "cpu", inter_column=True, seed=1)
    dev = scaled.distinct_synapses_device(cols, 0, ncol, "cuda:0", inter_column=True, seed=1)
    assert float(host[8].max()) * 20 > 100                      # some delay outlasts the 100-step refractory period (LONG path)
    # the CPU and CUDA generators of the inter-column draws differ (torch.Generator per device): the oracle gets the GPU's arrays
    arrays = [t.cpu().numpy() for t in dev]
    out = {}
    for name, cls in (("gpu", Engine), ("cpu", OracleEngine)):
        e = cls(n, 0.05, method="euler", accum="fixed", seed=5, **({"plan": "graph"} if cls is Engine else {}))
        e.set_neurons(neuron_rows_from_NeuPar(net["NeuPar"], net["I_DC"]), net["V0"][0], net["V0"][1])
        e.set_channels(**channels_from_STypPar(STypPar))
        if cls is Engine:
            e.set_synapses_dev(*dev)
        else:
            e.set_synapses(*arrays)
        e.run(steps)
        out[name] = (e.spikes(), e.read_state("V"), e.read_state("w"), e.read_syn_state("u"), e.read_syn_state("R"), e.counters())
        e.close()
    (gi, gs), gV, gw, gu, gR, gc = out["gpu"]
    (ci, cs), cV, cw, cu, cR, cc = out["cpu"]
    assert len(ci) > 500 and np.array_equal(gi, ci) and np.array_equal(gs, cs)
    assert np.array_equal(gV, cV) and np.array_equal(gw, cw) and np.array_equal(gu, cu) and np.array_equal(gR, cR)
    assert gc["syn_events"] == cc["syn_events"] and gc["spikes"] == cc["spikes"]
    n_intra = sum(len(c["src"]) for c in cols)
    assert torch.equal(dev[1][:n_intra].cpu(), host[1][:n_intra]) and dev[0].numel() == host[0].numel() == n_intra + scaled.inter_column_count(net["groups"], ncol)
