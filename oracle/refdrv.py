"""ctypes wrapper around oracle/_ref/libhdsdp_ref.so (the UNMODIFIED reference + oracle/ref_driver.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  Never by the product package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HDSDP_REFDRV_LIB selects another build exposing the same driver entry points (the integration build
# integration/_build/libhdsdp_integrated.so = unmodified reference host + CUDA hot path); one library per process.
REF_LIB = os.environ.get("HDSDP_REFDRV_LIB") or os.path.join(_HERE, "_ref", "libhdsdp_ref.so")
c_double_p = POINTER(c_double)
c_int_p = POINTER(c_int)

_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{REF_LIB} not built (run oracle/build_ref.sh where /root/reference exists)")
        l = ctypes.CDLL(REF_LIB)
        l.refdrv_create.restype = c_void_p
        l.refdrv_create.argtypes = [c_int]
        l.refdrv_add_cone.argtypes = [c_void_p, c_int, c_int, c_int_p, c_int_p, c_double_p]
        l.refdrv_finalize.argtypes = [c_void_p]
        l.refdrv_cone_type.argtypes = [c_void_p, c_int]
        l.refdrv_is_kkt_sparse.argtypes = [c_void_p]
        l.refdrv_set_point.argtypes = [c_void_p, c_double_p, c_double, c_double, c_double_p]
        l.refdrv_interior_check.argtypes = [c_void_p, c_double_p, c_double, c_int_p]
        if hasattr(l, "refdrv_build_xsx"):
            l.refdrv_build_xsx.argtypes = [c_void_p, c_int, c_double_p, c_double_p, c_int]
            l.refdrv_build_xsx.restype = None
        if hasattr(l, "refdrv_get_primal"):
            l.refdrv_get_primal.argtypes = [c_void_p, c_int, c_double, c_double_p, c_double_p, c_double_p, c_double_p]
            l.refdrv_get_primal.restype = None
        if hasattr(l, "refdrv_ratio_test"):
            l.refdrv_ratio_test.argtypes = [c_void_p, c_int, c_double, c_double_p, c_double, c_int, c_double_p]
        l.refdrv_build.argtypes = [c_void_p, c_int, c_int]
        l.refdrv_regularize.argtypes = [c_void_p, c_double]
        l.refdrv_get_M.argtypes = [c_void_p, c_double_p]
        l.refdrv_get_vectors.argtypes = [c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
        l.refdrv_get_sinv.argtypes = [c_void_p, c_int, c_double_p]
        l.refdrv_get_S.argtypes = [c_void_p, c_int, c_double_p]
        l.refdrv_get_Ldiag.argtypes = [c_void_p, c_int, c_double_p]
        l.refdrv_get_classification.argtypes = [c_void_p, c_int, c_int_p, c_int_p, c_int_p]
        l.refdrv_get_r1_sign.restype = c_double
        l.refdrv_get_r1_sign.argtypes = [c_void_p, c_int, c_int]
        l.refdrv_factorize.argtypes = [c_void_p]
        l.refdrv_solve_rhs.argtypes = [c_void_p, c_double_p, c_double_p]
        l.refdrv_time_iteration.argtypes = [c_void_p, c_int, c_double, c_int, c_int, c_double_p]
        if hasattr(l, "refdrv_cg_status"):
            l.refdrv_cg_status.argtypes = [c_void_p, c_int_p]
        l.refdrv_register_primal.argtypes = [c_void_p, POINTER(c_double_p)]
        l.refdrv_destroy.argtypes = [c_void_p]
        l.refdrv_optimize.argtypes = [c_int, c_int, c_int_p, c_int_p, POINTER(c_int_p), POINTER(c_int_p), POINTER(c_double_p),
                                      c_double_p, c_int, c_double_p, c_double_p]
        l.refdrv_read_sdpa.restype = c_void_p
        l.refdrv_read_sdpa.argtypes = [c_char_p]
        for name in ("nconstrs", "nblks", "nlp"):
            getattr(l, f"refdrv_sdpa_{name}").argtypes = [c_void_p]
        l.refdrv_sdpa_blkdim.argtypes = [c_void_p, c_int]
        l.refdrv_sdpa_rhs.restype = c_double_p
        l.refdrv_sdpa_rhs.argtypes = [c_void_p]
        l.refdrv_sdpa_beg.restype = c_int_p
        l.refdrv_sdpa_beg.argtypes = [c_void_p, c_int]
        l.refdrv_sdpa_idx.restype = c_int_p
        l.refdrv_sdpa_idx.argtypes = [c_void_p, c_int]
        l.refdrv_sdpa_elem.restype = c_double_p
        l.refdrv_sdpa_elem.argtypes = [c_void_p, c_int]
        _lib = l
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


class RefKKT:
    """Cone(s) + KKT of the reference at a chosen operating point."""

    def __init__(self, prob):
        self.l = lib()
        self.prob = prob
        self.m = prob.m
        self.h = self.l.refdrv_create(prob.m)
        for cone in prob.cones:
            beg = np.ascontiguousarray(cone.beg, dtype=np.int32)
            idx = np.ascontiguousarray(cone.idx, dtype=np.int32)
            elem = np.ascontiguousarray(cone.elem, dtype=np.float64)
            k = self.l.refdrv_add_cone(self.h, 0 if cone.kind == "sdp" else 1, cone.dim, _ip(beg), _ip(idx), _dp(elem))
            assert k >= 0
        rc = self.l.refdrv_finalize(self.h)
        if rc != 0:
            raise RuntimeError(f"reference finalize failed at stage {rc}")

    def set_point(self, y, tau, rd) -> float:
        y = np.ascontiguousarray(y, dtype=np.float64)
        ld = c_double(0.0)
        rc = self.l.refdrv_set_point(self.h, _dp(y), float(tau), float(rd), byref(ld))
        if rc != 0:
            raise RuntimeError(f"reference: S not positive definite / factorization failed (cone {rc - 1})")
        return ld.value

    def interior_check(self, y, tau) -> bool:
        y = np.ascontiguousarray(y, dtype=np.float64)
        f = c_int(0)
        self.l.refdrv_interior_check(self.h, _dp(y), float(tau), byref(f))
        return bool(f.value)

    def build_xsx(self, k, X, XSX, dual_mat=1) -> np.ndarray:
        """HConeBuildPrimalXSXDirection: XSX += X S X (in place on a Fortran-ordered copy, returned)."""
        X = np.asfortranarray(X, dtype=np.float64); out = np.asfortranarray(XSX, dtype=np.float64).copy(order="F")
        self.l.refdrv_build_xsx(self.h, int(k), _dp(X), _dp(out), int(dual_mat))
        return out

    def get_primal(self, k, dim, mu, y, dy) -> np.ndarray:
        """HConeGetPrimal on cone k (reference primal recovery)."""
        y = np.ascontiguousarray(y, dtype=np.float64); dy = np.ascontiguousarray(dy, dtype=np.float64)
        X = np.full((dim, dim), np.nan, order="F"); aux = np.zeros((dim, dim), order="F")
        self.l.refdrv_get_primal(self.h, int(k), float(mu), _dp(y), _dp(dy), _dp(X), _dp(aux))
        return X

    def ratio_test(self, k, dtau, dy, ada_ratio, which=0) -> float:
        """HConeRatioTest on cone k (reference Lanczos); set_point must have been called."""
        dy = np.ascontiguousarray(dy, dtype=np.float64)
        step = c_double(0.0)
        rc = self.l.refdrv_ratio_test(self.h, int(k), float(dtau), _dp(dy), float(ada_ratio), int(which), byref(step))
        if rc != 0:
            raise RuntimeError(f"reference HConeRatioTest failed rc={rc}")
        return step.value

    def build(self, type_kkt=0, strategy=-1):
        rc = self.l.refdrv_build(self.h, int(type_kkt), int(strategy))
        if rc != 0:
            raise RuntimeError(f"reference HKKTBuildUp failed rc={rc}")

    def regularize(self, reg):
        self.l.refdrv_regularize(self.h, float(reg))

    def get_M(self):
        M = np.zeros((self.m, self.m), order="F")
        self.l.refdrv_get_M(self.h, _dp(M))
        return M

    def get_vectors(self):
        a = np.zeros(self.m); ard = np.zeros(self.m); ac = np.zeros(self.m); s = np.zeros(4)
        self.l.refdrv_get_vectors(self.h, _dp(a), _dp(ard), _dp(ac), _dp(s))
        return {"dASinvVec": a, "dASinvRdSinvVec": ard, "dASinvCSinvVec": ac, "dCSinvCSinv": s[0], "dCSinv": s[1],
                "dCSinvRdSinv": s[2], "dTraceSinv": s[3]}

    def get_S(self, k):
        n = self.prob.cones[k].dim
        S = np.zeros((n, n), order="F")
        self.l.refdrv_get_S(self.h, k, _dp(S))
        return S

    def get_sinv(self, n):
        X = np.zeros((n, n), order="F")
        self.l.refdrv_get_sinv(self.h, n, _dp(X))
        return X

    def get_Ldiag(self, k):
        d = np.zeros(self.prob.cones[k].dim)
        self.l.refdrv_get_Ldiag(self.h, k, _dp(d))
        return d

    def classification(self, k):
        t = np.zeros(self.m + 1, dtype=np.int32); p = np.zeros(self.m, dtype=np.int32); s = np.zeros(self.m, dtype=np.int32)
        kind = self.l.refdrv_get_classification(self.h, k, _ip(t), _ip(p), _ip(s))
        return {"cone_kind": kind, "types": t, "perm": p, "strategies": s}

    def r1_sign(self, k, i):
        return self.l.refdrv_get_r1_sign(self.h, k, i)

    def factorize(self):
        return self.l.refdrv_factorize(self.h)

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        out = np.zeros(self.m)
        rc = self.l.refdrv_solve_rhs(self.h, _dp(rhs), _dp(out))
        if rc != 0:
            raise RuntimeError("reference HKKTSolve failed")
        return out

    def time_iteration(self, type_kkt=0, reg=0.0, nsolve=2, nrep=1):
        t = np.zeros(4)
        rc = self.l.refdrv_time_iteration(self.h, type_kkt, float(reg), nsolve, nrep, _dp(t))
        if rc != 0:
            raise RuntimeError(f"reference iteration failed rc={rc}")
        return {"build": t[0], "factorize": t[1], "solve": t[2], "total": t[3]}

    def register_primal(self, Xs):
        """HKKTRegisterPSDP: one n x n primal matrix per cone (used by HKKTBuildUp(KKT_TYPE_PRIMAL) in place of S^-1)."""
        self._primal_keep = [np.asfortranarray(X, dtype=np.float64) for X in Xs]
        self._primal_arr = (c_double_p * len(Xs))(*[_dp(X) for X in self._primal_keep])
        self.l.refdrv_register_primal(self.h, self._primal_arr)

    def cg_status(self):
        """The reference's PCG-on-M state after the last solve: did it fall back to a Cholesky preconditioner?"""
        if not hasattr(self.l, "refdrv_cg_status"):
            return None
        out = np.zeros(4, dtype=np.int32)
        if self.l.refdrv_cg_status(self.h, _ip(out)) != 0:
            return None
        return {"use_jacobi": int(out[0]), "n_iters": int(out[1]), "n_solves": int(out[2]), "status": int(out[3])}

    def close(self):
        if self.h:
            self.l.refdrv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def optimize(prob, max_iter=0):
    """Full HDSDPOptimize through the reference's public API."""
    l = lib()
    nc = len(prob.cones)
    kinds = np.array([0 if c.kind == "sdp" else 1 for c in prob.cones], dtype=np.int32)
    dims = np.array([c.dim for c in prob.cones], dtype=np.int32)
    begs = [np.ascontiguousarray(c.beg, dtype=np.int32) for c in prob.cones]
    idxs = [np.ascontiguousarray(c.idx, dtype=np.int32) for c in prob.cones]
    elems = [np.ascontiguousarray(c.elem, dtype=np.float64) for c in prob.cones]
    B = (c_int_p * nc)(*[_ip(b) for b in begs])
    I = (c_int_p * nc)(*[_ip(b) for b in idxs])
    E = (c_double_p * nc)(*[_dp(b) for b in elems])
    rhs = np.ascontiguousarray(prob.rhs, dtype=np.float64)
    out = np.zeros(11)
    y = np.zeros(prob.m)
    rc = l.refdrv_optimize(prob.m, nc, _ip(kinds), _ip(dims), B, I, E, _dp(rhs), int(max_iter), _dp(out), _dp(y))
    return {"retcode": rc, "pObj": out[0], "dObj": out[1], "iterations": int(out[2]), "status": int(out[3]), "seconds": out[4],
            "dimacs": out[5:11].copy(), "y": y}


def read_sdpa(path):
    """The reference's own SDPA reader (interface/hdsdp_file_io.c:34) -> Problem (arrays are copied)."""
    from hdsdp_b200.problem import ConeData, Problem
    l = lib()
    h = l.refdrv_read_sdpa(path.encode())
    if not h:
        raise RuntimeError(f"reference HReadSDPA failed on {path}")
    m = l.refdrv_sdpa_nconstrs(h); nb = l.refdrv_sdpa_nblks(h); nlp = l.refdrv_sdpa_nlp(h)
    rhs = np.ctypeslib.as_array(l.refdrv_sdpa_rhs(h), shape=(m,)).copy()
    cones = []
    for k in range(nb + (1 if nlp > 0 else 0)):
        beg = np.ctypeslib.as_array(l.refdrv_sdpa_beg(h, k), shape=(m + 2,)).copy()
        nnz = int(beg[m + 1])
        idx = np.ctypeslib.as_array(l.refdrv_sdpa_idx(h, k), shape=(max(nnz, 1),))[:nnz].copy()
        elem = np.ctypeslib.as_array(l.refdrv_sdpa_elem(h, k), shape=(max(nnz, 1),))[:nnz].copy()
        if k < nb:
            cones.append(ConeData("sdp", l.refdrv_sdpa_blkdim(h, k), beg.astype(np.int32), idx.astype(np.int32), elem))
        else:
            cones.append(ConeData("lp", nlp, beg.astype(np.int32), idx.astype(np.int32), elem))
    return Problem(m=m, cones=cones, rhs=rhs, name=os.path.basename(path))
