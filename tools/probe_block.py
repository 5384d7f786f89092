"""potrf time vs look-ahead block size (chol_block) at mid sizes."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
lib = _lib.require_gpu(0)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())
for n in (1500, 4096, 8192, 16384):
    h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
    G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n)); del G
    info = ctypes.c_int(0)
    rec = {"n": n}
    for blk in (-1, 0, 256, 512):
        lib.hdsdpcu_set_option(b"chol_block", blk)
        best = 1e30
        for it in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info)); e1.record(st); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        rec[f"blk{blk}_ms"] = round(best, 3)
    rec["flops_ms_at_30TF"] = round(n ** 3 / 3 / 30e12 * 1e3, 3)
    print(json.dumps(rec), flush=True)
    lib.hdsdpcu_linsys_destroy(ctypes.byref(h)); del A
