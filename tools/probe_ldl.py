"""python tools/probe_ldl.py N -- time the LDL^T fallback factorisation (bounded Bunch-Kaufman leaves vs the unpivoted variant)
next to the Cholesky of the same N x N SPD matrix."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = _lib.require_gpu(0)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())
G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n)); del G
torch.cuda.synchronize()
info = ctypes.c_int(0)
for mode, piv in (("cholesky", 1), ("ldl_bunch_kaufman", 1), ("ldl_unpivoted", 0)):
    h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
    lib.hdsdpcu_set_option(b"ldl_pivot", piv)
    if mode != "cholesky":
        assert lib.hdsdpcu_linsys_set_indefinite(h, 1) == 0
    best = 1e30
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info)); e1.record(st); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"n": n, "mode": mode, "ms": best, "tflops": n ** 3 / 3 / best / 1e9, "info": info.value}), flush=True)
    lib.hdsdpcu_linsys_destroy(ctypes.byref(h))
lib.hdsdpcu_set_option(b"ldl_pivot", 1)
