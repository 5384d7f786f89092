"""Multi-process check of the distributed Schur path (run under torchrun, one rank per GPU):
every rank assembles only its block columns, the factorisation exchanges panels over peer memory, and the result
(diag of the factor, both solutions, side vectors) must match a single-GPU KKT object built by the same process."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hdsdp_b200 import _lib, api, problem  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.require_gpu(local)
n, ne, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prob = problem.gen_theta(n, ne, seed=2)


def allgather(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out


sdp, lps, kkt = api.build_problem(prob)
kkt.dist_init(rank, world, nb, allgather)
ref = api.KKT(prob.m, sdp)  # same cone, single-GPU Schur object
cone = sdp[0]; cone.set_start(bench.RD)
worst = 0.0
for it in range(3):
    y = bench.theta_point(prob.m, n, it)
    cone.update(bench.TAU, y); assert cone.factorize()
    res = []
    for k in (ref, kkt):
        k.build_up(api.KKT_TYPE_INFEASIBLE); k.regularize(bench.KKT_REG)
        v = k.export()
        assert k.factorize() == 0
        res.append((v, k.solve(prob.rhs), k.solve(v["dASinvVec"]), np.tril(k.get_matrix())))
    (v0, a0, b0, M0), (v1, a1, b1, M1) = res
    own = (np.arange(prob.m) // nb) % world == rank
    e = [np.abs(v0[key] - v1[key]).max() / max(np.abs(v0[key]).max(), 1e-300) for key in ("dASinvVec", "dASinvRdSinvVec")]
    e.append(np.abs(M0[:, own] - M1[:, own]).max() / np.abs(M0).max())
    e.append(np.abs(a0 - a1).max() / np.abs(a0).max()); e.append(np.abs(b0 - b1).max() / np.abs(b0).max())
    worst = max(worst, max(e))
    print(json.dumps({"rank": rank, "it": it, "errs": [float(x) for x in e]}), flush=True)
# LDL^T fallback across processes: shift the diagonal so that a few eigenvalues go negative on BOTH objects
if os.environ.get("DIST_CHECK_LDL", "1") == "1":
    y = bench.theta_point(prob.m, n, 7)
    cone.update(bench.TAU, y); assert cone.factorize()
    sols = []
    for k in (ref, kkt):
        k.build_up(api.KKT_TYPE_INFEASIBLE)
        if k is ref:
            M = np.tril(k.get_matrix()); M = M + np.tril(M, -1).T
            lam = np.linalg.eigvalsh(M); shift = 0.5 * (lam[3] + lam[4])
        k.build_up_extra_bound(-shift * np.ones(prob.m), np.zeros(prob.m))
        assert k.factorize() == 0
        sols.append(k.solve(prob.rhs))
    e = float(np.abs(sols[0] - sols[1]).max() / np.abs(sols[0]).max())
    truth = np.linalg.solve(M - shift * np.eye(prob.m), prob.rhs)
    e2 = float(np.abs(sols[1] - truth).max() / np.abs(truth).max())
    print(json.dumps({"rank": rank, "ldl_dist_vs_single": e, "ldl_dist_vs_numpy": e2}), flush=True)
    worst = max(worst, e, e2 * 1e-2)
t = torch.tensor([worst], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"dist_check": "ok" if float(t) < 1e-9 else "FAIL", "world": world, "worst": float(t), "m": prob.m, "nb": nb}), flush=True)
dist.destroy_process_group()
sys.exit(0 if float(t) < 1e-9 else 1)
