"""Partitioned GPU run (NCCL spike all-gather inside libhhd) vs the single-GPU run of the same network: bit-identical.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import default_idc, load_golden  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200.engine import Engine, load_codes_network  # noqa: E402
from rescience_hass_hertag_durstewitz_2016_b200.partition import load_partition, partition_network  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    for name, steps in (("net_n60x5_s1", 6000), ("net_n200_s0", 6000)):
        g = load_golden(name)
        gnet, locs = partition_network(g["NeuPar"], g["V0"], g["SynPar"], g["source_arr"], g["target_arr"], default_idc(g), world)
        loc = locs[rank]
        e = Engine(loc["n_local"], 0.05, method="rk2", accum="fixed", seed=23, device=lr, rank=rank, n_ranks=world,
                   n_global=loc["n_global"], global_offset=loc["global_offset"], plan=os.environ.get("HHD_CHECK_PLAN", "graph"))
        load_partition(e, loc, g["STypPar"])
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(e.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        e.comm_init(bytes(uid.cpu().tolist()))
        print('rank', rank, name, 'comm ready', flush=True)
        e.run(steps)
        print('rank', rank, name, 'ran', flush=True)
        i, s = e.spikes()
        mine = dict(i=i, s=s, V=e.read_state("V"), ev=e.counters()["syn_events"])
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        print('rank', rank, name, 'gathered', flush=True)
        if rank == 0:
            ref = Engine(g["NeuPar"].shape[1], 0.05, method="rk2", accum="fixed", seed=23, device=lr, plan="graph")
            load_codes_network(ref, gnet["NeuPar"], gnet["V0"], g["STypPar"], gnet["SynPar"], gnet["source_arr"], gnet["target_arr"], gnet["I_DC"])
            ref.run(steps)
            ri, rs = ref.spikes()
            mi = np.concatenate([p["i"] for p in parts])
            ms = np.concatenate([p["s"] for p in parts])
            o = np.lexsort((mi, ms))
            same = (np.array_equal(mi[o], ri) and np.array_equal(ms[o], rs)
                    and np.array_equal(np.concatenate([p["V"] for p in parts]), ref.read_state("V"))
                    and sum(p["ev"] for p in parts) == ref.counters()["syn_events"])
            print("%s: %d ranks, %d steps, %d spikes, %d synaptic events: %s" %
                  (name, world, steps, len(ri), ref.counters()["syn_events"], "IDENTICAL to the single-GPU run" if same else "MISMATCH"), flush=True)
            ok = ok and same and len(ri) > 100
        e.close()
    ok = ring_case(rank, world, lr) and ok
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


def ring_case(rank, world, lr, ncol=40, steps=6000):
    """The bench's own network generator (scaled.multicolumn_synapses_device: tiled columns + inter-column synapses with
    delays beyond the refractory period), 40 columns of the 207-neuron template: every rank draws ITS target columns, the
    single-GPU reference draws all of them — the generator restarts its random stream per column chunk, so the union of the
    partitions is the same network.  Release-failure keys follow the position in the single-GPU arrays (syn_id_base)."""
    from oracle_binding import golden_groups
    from rescience_hass_hertag_durstewitz_2016_b200.scaled import multicolumn_synapses_device, tile_columns
    g = load_golden("net_n200_s0")
    n1 = g["NeuPar"].shape[1]
    groups = [x[0] for x in golden_groups(g)]
    dev = "cuda:%d" % lr
    per = ncol // world
    c0 = rank * per
    net = tile_columns(g["NeuPar"], g["V0"], np.zeros((7, 0)), [], [], default_idc(g), per)
    full = multicolumn_synapses_device(g["SynPar"], g["source_arr"], g["target_arr"], groups, ncol, n1, dev, 0, ncol, seed=1)
    # the single-GPU arrays ordered by owning rank (stable), so that a rank's synapses are one contiguous slice of them
    owner = (full[1].to(torch.int64) // (n1 * per))
    perm = torch.argsort(owner, stable=True)
    full = tuple(x[perm].contiguous() for x in full)
    counts = torch.bincount(owner, minlength=world).cpu().numpy()
    start = int(counts[:rank].sum())
    mine = tuple(x[start:start + int(counts[rank])].contiguous() for x in full)
    own = multicolumn_synapses_device(g["SynPar"], g["source_arr"], g["target_arr"], groups, ncol, n1, dev, c0, c0 + per, seed=1)
    same_net = all(torch.equal(torch.sort(a.to(torch.float64))[0], torch.sort(b.to(torch.float64))[0]) for a, b in zip(own, mine))
    e = Engine(n1 * per, 0.05, method="euler", accum="fixed", seed=5, device=lr, rank=rank, n_ranks=world, n_global=n1 * ncol,
               global_offset=c0 * n1, plan="graph")
    load_codes_network(e, net["NeuPar"], net["V0"], g["STypPar"], np.zeros((7, 0)), [], [], net["I_DC"])
    e.set_has_outgoing(np.ones(n1 * ncol, dtype=np.uint8))
    e.set_synapses_dev(*mine, syn_id_base=start)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(e.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    e.comm_init(bytes(uid.cpu().tolist()))
    e.run(steps)
    i, s = e.spikes()
    parts = [None] * world
    dist.all_gather_object(parts, dict(i=i, s=s, V=e.read_state("V"), ev=e.counters()["syn_events"], same_net=same_net))
    e.close()
    if rank != 0:
        return True
    netf = tile_columns(g["NeuPar"], g["V0"], np.zeros((7, 0)), [], [], default_idc(g), ncol)
    ref = Engine(n1 * ncol, 0.05, method="euler", accum="fixed", seed=5, device=lr, plan="graph")
    load_codes_network(ref, netf["NeuPar"], netf["V0"], g["STypPar"], np.zeros((7, 0)), [], [], netf["I_DC"])
    ref.set_has_outgoing(np.ones(n1 * ncol, dtype=np.uint8))
    ref.set_synapses_dev(*full)
    ref.run(steps)
    ri, rs = ref.spikes()
    mi = np.concatenate([p["i"] for p in parts])
    ms = np.concatenate([p["s"] for p in parts])
    o = np.lexsort((mi, ms))
    same = (np.array_equal(mi[o], ri) and np.array_equal(ms[o], rs) and np.array_equal(np.concatenate([p["V"] for p in parts]), ref.read_state("V"))
            and sum(p["ev"] for p in parts) == ref.counters()["syn_events"] and all(p["same_net"] for p in parts))
    print("ring of %d columns (bench generator, %d synapses, delays up to %.1f ms): %d ranks, %d steps, %d spikes, %d synaptic events: %s" %
          (ncol, full[0].numel(), float(full[8].max()), world, steps, len(ri), ref.counters()["syn_events"],
           "IDENTICAL to the single-GPU run" if same else "MISMATCH"), flush=True)
    ref.close()
    return same and len(ri) > 1000


if __name__ == "__main__":
    main()
