"""The N>1 path on CPU: two gloo ranks each drive one partition of the CPU oracle, exchange the step's spike indices
with an all-gather (what libhhd does with NCCL on the GPUs) and must reproduce the single-process run bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle_binding import OracleEngine, default_idc, load_golden
from rescience_hass_hertag_durstewitz_2016_b200.engine import load_codes_network
from rescience_hass_hertag_durstewitz_2016_b200.partition import load_partition, order_by_target_rank, partition_network, split_bounds

STEPS = 3000
CAP = 512


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    _, locs = partition_network(g["NeuPar"], g["V0"], g["SynPar"], g["source_arr"], g["target_arr"], default_idc(g), world)
    loc = locs[rank]
    e = OracleEngine(loc["n_local"], 0.05, method="rk2", accum="fp64", seed=17, rank=rank, n_ranks=world,
                     n_global=loc["n_global"], global_offset=loc["global_offset"])
    load_partition(e, loc, g["STypPar"])
    send = torch.zeros(CAP + 1, dtype=torch.int32)
    recv = [torch.zeros(CAP + 1, dtype=torch.int32) for _ in range(world)]
    for _ in range(STEPS):
        mine = e.step_begin()
        assert len(mine) <= CAP
        send[0] = len(mine)
        send[1:1 + len(mine)] = torch.from_numpy(mine.astype(np.int32))
        dist.all_gather(recv, send)                      # fixed-size (count, ids...) buffers, as on the GPU
        allspk = np.concatenate([r[1:1 + int(r[0])].numpy() for r in recv])
        e.step_end(np.sort(allspk))
    i, s = e.spikes()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), i=i, s=s, V=e.read_state("V"), w=e.read_state("w"),
             u=e.read_syn_state("u"), ev=e.counters()["syn_events"])
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["net_n60x5_s1"])
def test_two_rank_partition_matches_single_process(tmp_path, name):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    g = load_golden(name)
    gnet, locs = partition_network(g["NeuPar"], g["V0"], g["SynPar"], g["source_arr"], g["target_arr"], default_idc(g), world)
    ref = OracleEngine(g["NeuPar"].shape[1], 0.05, method="rk2", accum="fp64", seed=17)
    load_codes_network(ref, gnet["NeuPar"], gnet["V0"], g["STypPar"], gnet["SynPar"], gnet["source_arr"], gnet["target_arr"], gnet["I_DC"])
    ref.run(STEPS)
    ri, rs = ref.spikes()
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    mi = np.concatenate([p["i"] for p in parts])
    ms = np.concatenate([p["s"] for p in parts])
    order = np.lexsort((mi, ms))
    assert len(ri) > 200
    assert np.array_equal(mi[order], ri) and np.array_equal(ms[order], rs)
    assert np.array_equal(np.concatenate([p["V"] for p in parts]), ref.read_state("V"))
    assert np.array_equal(np.concatenate([p["w"] for p in parts]), ref.read_state("w"))
    assert np.array_equal(np.concatenate([p["u"] for p in parts]), ref.read_syn_state("u"))
    assert sum(int(p["ev"]) for p in parts) == ref.counters()["syn_events"]
    cross = (np.searchsorted(split_bounds(g["NeuPar"].shape[1], world), g["source_arr"], side="right") !=
             np.searchsorted(split_bounds(g["NeuPar"].shape[1], world), g["target_arr"], side="right"))
    assert cross.sum() > 100           # the partition really cuts synapses


def test_partition_bookkeeping():
    g = load_golden("net_n200_s0")
    n = g["NeuPar"].shape[1]
    b = split_bounds(n, 3)
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) > 0)
    assert list(split_bounds(1003 * 8, 4, align=1003)) == [0, 2006, 4012, 6018, 8024]
    perm, sb = order_by_target_rank(g["target_arr"], b)
    assert sorted(perm) == list(range(len(perm))) and sb[-1] == len(perm)
    gnet, locs = partition_network(g["NeuPar"], g["V0"], g["SynPar"], g["source_arr"], g["target_arr"], default_idc(g), 3)
    assert sum(l["SynPar"].shape[1] for l in locs) == g["SynPar"].shape[1]
    for r, l in enumerate(locs):
        assert np.all((l["target_arr"] >= b[r]) & (l["target_arr"] < b[r + 1]))
        assert l["syn_id_base"] == sb[r] and l["global_offset"] == b[r]
        assert np.array_equal(l["SynPar"], gnet["SynPar"][:, sb[r]:sb[r + 1]])
    assert locs[0]["has_out"].sum() == len(np.unique(g["source_arr"]))
