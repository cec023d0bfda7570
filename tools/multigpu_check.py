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
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
