"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): column ownership partitions the Schur matrix,
the IPC-blob handshake gathers in rank order, and the reference arm of bench.py only runs on rank 0."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, m, nb, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from hdsdp_b200 import _lib
    lib = _lib.lib()
    mine = np.array([c for c in range(m) if lib.hdsdpcu_dist_owner(c, nb, world) == rank], dtype=np.int64)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    nbytes = lib.hdsdpcu_dist_blob_bytes()
    blob = bytes([rank + 1]) * nbytes                      # stand-in for the IPC handles (needs a GPU to export real ones)
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    ok = all(len(b) == nbytes and b[0] == r + 1 for r, b in enumerate(blobs))
    allc = np.concatenate(gathered)
    ok = ok and len(allc) == m and np.array_equal(np.sort(allc), np.arange(m))
    # every rank owns whole blocks and the load differs by at most one block
    sizes = [len(g) for g in gathered]
    ok = ok and max(sizes) - min(sizes) <= nb and all((g // nb % world == r).all() for r, g in enumerate(gathered))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t))
    dist.destroy_process_group()


@pytest.mark.parametrize("m,nb", [(5001, 256), (50000, 512)])
def test_ownership_and_handshake_world2(m, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + m) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1.0


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
