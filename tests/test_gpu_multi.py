"""Partitioned runs (one process per GPU, spike exchange over peer memory / NCCL inside libhhd) against the single-GPU run of the
same network: spike lists, membrane potentials and event counts bit-identical — on the two golden networks and on a 40-column
ring made by the bench's own generator (scaled.multicolumn_synapses_device).  Needs >= 2 visible GPUs; tools/multigpu_check.py
is the program the ranks run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["peer_memory", "nccl_all_gather"])
def test_partitioned_run_is_bit_identical_to_the_single_gpu_run(exchange):
    """Both exchange paths: self-validating words over IPC-mapped peer memory (default) and the ncclAllGather + k_append fallback
    (HHD_NO_P2P=1)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (found %d)" % n)
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py")]
    env = dict(os.environ)
    if exchange == "nccl_all_gather":
        env["HHD_NO_P2P"] = "1"
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
    assert r.stdout.count("IDENTICAL to the single-GPU run") == 3 and "MISMATCH" not in r.stdout
