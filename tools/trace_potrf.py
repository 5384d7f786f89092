"""HDSDPCU_TRACE=1 python tools/trace_potrf.py N [NB] -- per-step timing of the blocked look-ahead Cholesky."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
n = int(sys.argv[1]); nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lib = _lib.require_gpu(0)
lib.hdsdpcu_set_option(b"chol_block", nb)
h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n)); del G
torch.cuda.synchronize()
info = ctypes.c_int(0)
for _ in range(2):
    lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info))
print("info", info.value)
