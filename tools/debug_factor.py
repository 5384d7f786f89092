"""Factor the bench's Schur matrix with several (gemm variant, block) combinations and compare."""
import ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hdsdp_b200 import _lib, api, problem
n, ne = int(sys.argv[1]), int(sys.argv[2])
lib = _lib.require_gpu(0)
prob = problem.gen_theta(n, ne, seed=2)
sdp, lps, kkt = api.build_problem(prob)
cone = sdp[0]; cone.set_start(bench.RD)
y = bench.theta_point(prob.m, n, 0)
cone.update(bench.TAU, y); assert cone.factorize()
kkt.build_up(0); kkt.regularize(bench.KKT_REG)
mp = lib.hdsdpcu_kkt_padded_dim(kkt.h)
Mptr = lib.hdsdpcu_kkt_matrix_dev(kkt.h)
h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), prob.m) == 0
info = ctypes.c_int(0)
ref = None
for v, nb in ((1, 0), (3, 0), (1, 1024), (3, 1024), (3, 2048), (3, 2048), (1, 2048)):
    lib.hdsdpcu_set_option(b"gemm_variant", v); lib.hdsdpcu_set_option(b"chol_block", nb)
    lib.hdsdpcu_linsys_numeric_dev(h, Mptr, mp, ctypes.byref(info))
    d = np.zeros(prob.m); lib.hdsdpcu_linsys_getdiag(h, d.ctypes.data_as(_lib.c_double_p))
    if ref is None and info.value == 0:
        ref = d.copy()
    rel = float(np.abs(d - ref).max() / np.abs(ref).max()) if ref is not None else None
    print(json.dumps({"variant": v, "block": nb, "info": info.value, "diag_min": float(d.min()), "diag_max": float(d.max()), "diag_rel_diff": rel,
                      "nan": bool(np.isnan(d).any())}), flush=True)
