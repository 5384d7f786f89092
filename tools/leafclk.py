import ctypes, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from hdsdp_b200 import _lib
lib = _lib.require_gpu(0)
lib.hdsdpcu_debug_leafclk.argtypes=[ctypes.c_void_p]
n=128
h = ctypes.c_void_p(); lib.hdsdpcu_linsys_create(ctypes.byref(h), n)
G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n))
info=ctypes.c_int(0)
for it in range(3):
    lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info)); lib.hdsdpcu_sync()
out=np.zeros(40,dtype=np.int64); lib.hdsdpcu_debug_leafclk(out.ctypes.data)
t=out[:31]-out[0]
print("load",t[1], "first diag block (a)", t[2]-t[1])
prev=t[2]
for p in range(7):
    print("panel",p,"(b) substitution",t[3+3*p]-prev,"(c)+next (a)",t[4+3*p]-t[3+3*p]); prev=t[4+3*p]
print("Lwriteback",t[27]-t[26],"W",t[28]-t[27],"inverse",t[29]-t[28],"dinv store",t[30]-t[29],"total",t[30])
