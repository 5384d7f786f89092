import ctypes, os, sys
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from hdsdp_b200 import _lib
n = int(sys.argv[1]); nb = int(sys.argv[2])
lib = _lib.require_gpu(0)
lib.hdsdpcu_set_option(b"chol_block", nb); lib.hdsdpcu_set_option(b"chol_graph", 0)
h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n)); del G
torch.cuda.synchronize()
info = ctypes.c_int(0)
for _ in range(2):
    lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info))
lib.hdsdpcu_sync()
