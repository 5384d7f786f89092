"""[HDSDPCU_TRACE=1] python tools/trace_potrf.py N NB[,NB..] VARIANT[,VARIANT..] [THIN SCHED GRAPH PARTITION] -- time the Cholesky of an N x N SPD matrix
for every (block, gemm variant) combination; with HDSDPCU_TRACE=1 the library prints its per-step timing to stderr."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
n = int(sys.argv[1])
nbs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024]
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4]
thin = int(sys.argv[4]) if len(sys.argv) > 4 else -1
sched = int(sys.argv[5]) if len(sys.argv) > 5 else -1
graph = int(sys.argv[6]) if len(sys.argv) > 6 else -1
part = int(sys.argv[7]) if len(sys.argv) > 7 else -2
lib = _lib.require_gpu(0)
if thin >= 0:
    lib.hdsdpcu_set_option(b"gemm_thin", thin)
if sched >= 0:
    lib.hdsdpcu_set_option(b"chol_sched", sched)
if graph >= 0:
    lib.hdsdpcu_set_option(b"chol_graph", graph)
if part >= -1:
    lib.hdsdpcu_set_option(b"chol_partition", part)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())
h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n)); del G
torch.cuda.synchronize()
info = ctypes.c_int(0)
for v in variants:
    for nb in nbs:
        lib.hdsdpcu_set_option(b"gemm_variant", v); lib.hdsdpcu_set_option(b"chol_block", nb)
        lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info))
        best = 1e30
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info)); e1.record(st); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        print(json.dumps({"n": n, "chol_block": nb, "gemm_variant": v, "thin": thin, "sched": sched, "graph": graph, "partition": part, "info": info.value, "ms": best * 1e3, "tflops": n ** 3 / 3 / best / 1e12}), flush=True)
