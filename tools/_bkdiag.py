import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hdsdp_b200 import _lib
from hdsdp_b200.api import DenseLinsys
from test_gpu_linsys import sym_needs_pivoting
lib = _lib.require_gpu()
for n, kind in [(128, "saddle"), (1024, "saddle"), (2048, "saddle"), (1024, "zero_diag")]:
    A = np.asfortranarray(sym_needs_pivoting(n, kind, 3 * n + 1))
    B = np.random.RandomState(4).standard_normal((n, 2))
    normA = np.abs(A).sum(axis=1).max()
    Xl = np.linalg.solve(A, B)
    out = {}
    for piv in (1, 0):
        lib.hdsdpcu_set_option(b"ldl_pivot", piv)
        ls = DenseLinsys(n); lib.hdsdpcu_linsys_set_indefinite(ls.h, 1)
        rc = ls.numeric(A)
        neg, pert = ctypes.c_int(-1), ctypes.c_int(-1)
        lib.hdsdpcu_linsys_inertia(ls.h, ctypes.byref(neg), ctypes.byref(pert))
        X = ls.solve(B); ls.close()
        berr = np.abs(A @ X - B).max() / (normA * np.abs(X).max() + np.abs(B).max()) if np.isfinite(X).all() else np.inf
        ferr = np.abs(X - Xl).max() / np.abs(Xl).max() if np.isfinite(X).all() else np.inf
        out[piv] = (rc, neg.value, pert.value, berr, ferr)
    print(n, kind, "cond %.2e" % np.linalg.cond(A), "neg_true", int((np.linalg.eigvalsh(A) < 0).sum()), "BK", out[1], "static", out[0], flush=True)
