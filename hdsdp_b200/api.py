"""Host-side mirror of the reference's cone / KKT / linear-system interfaces, over the C ABI.

Names and argument meaning follow the reference so the parity tests read like its own harness
(reference tests/test_file_io.c:356-467):

    HFpLinsys*   linalg/hdsdp_linsolver.h:17-31      -> class DenseLinsys
    HCone*       interface/hdsdp_conic.h:27-63       -> class SDPCone / LPCone
    HKKT*        interface/hdsdp_schur.h:10-22       -> class KKT

Everything here only marshals numpy arrays into include/hdsdpcu.h calls; all arithmetic runs in
libhdsdp_cuda.so on the GPU.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_void_p
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import c_double_p, c_int_p, check
from .problem import ConeData, Problem

KKT_TYPE_INFEASIBLE, KKT_TYPE_CORRECTOR, KKT_TYPE_HOMOGENEOUS, KKT_TYPE_PRIMAL = 0, 1, 2, 3
BUFFER_DUALVAR, BUFFER_DUALCHECK, BUFFER_DUALSTEP = 0, 1, 2
SDP_COEFF_ZERO, SDP_COEFF_SPARSE, SDP_COEFF_DENSE, SDP_COEFF_SPR1, SDP_COEFF_DSR1 = 0, 1, 2, 3, 4


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32
    return a.ctypes.data_as(c_int_p)


class DenseLinsys:
    """hdsdp_linsys_fp with the CUDA dense back-end (HDSDP_LINSYS_DENSE_DIRECT twin)."""

    def __init__(self, n: int):
        self.lib = _lib.require_gpu()
        self.n = n
        self.h = c_void_p()
        check(self.lib.hdsdpcu_linsys_create(byref(self.h), n), "HFpLinsysCreate")

    def numeric(self, A: np.ndarray) -> int:
        """HFpLinsysNumeric: returns the retcode (1 = FAILED when A is not positive definite)."""
        A = np.asfortranarray(A, dtype=np.float64)
        return self.lib.hdsdpcu_linsys_numeric(self.h, None, None, _dp(A))

    def psd_check(self, A: np.ndarray) -> bool:
        A = np.asfortranarray(A, dtype=np.float64)
        flag = c_int(0)
        check(self.lib.hdsdpcu_linsys_psdcheck(self.h, None, None, _dp(A), byref(flag)), "HFpLinsysPsdCheck")
        return bool(flag.value)

    def _rhs(self, b):
        b = np.array(b, dtype=np.float64, order="F", copy=True)
        nrhs = 1 if b.ndim == 1 else b.shape[1]
        return b, nrhs

    def fsolve(self, b):
        x, nrhs = self._rhs(b)
        self.lib.hdsdpcu_linsys_fsolve(self.h, nrhs, _dp(x), None)
        return x

    def bsolve(self, b):
        x, nrhs = self._rhs(b)
        self.lib.hdsdpcu_linsys_bsolve(self.h, nrhs, _dp(x), None)
        return x

    def solve(self, b):
        x, nrhs = self._rhs(b)
        check(self.lib.hdsdpcu_linsys_solve(self.h, nrhs, _dp(x), None), "HFpLinsysSolve")
        return x

    def get_diag(self):
        d = np.zeros(self.n)
        check(self.lib.hdsdpcu_linsys_getdiag(self.h, _dp(d)), "HFpLinsysGetDiag")
        return d

    def invert(self):
        inv = np.zeros((self.n, self.n), order="F")
        self.lib.hdsdpcu_linsys_invert(self.h, _dp(inv), None)
        return inv

    def close(self):
        if self.h:
            self.lib.hdsdpcu_linsys_destroy(byref(self.h))
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SDPCone:
    """hdsdp_cone for an SDP block (dense or sparse cone type in the reference; one device image here)."""

    def __init__(self, data: ConeData, m: int):
        assert data.kind == "sdp"
        self.lib = _lib.require_gpu()
        self.m, self.n = m, data.dim
        self.h = c_void_p()
        self._keep = (np.ascontiguousarray(data.beg, dtype=np.int32), np.ascontiguousarray(data.idx, dtype=np.int32),
                      np.ascontiguousarray(data.elem, dtype=np.float64))
        check(self.lib.hdsdpcu_cone_create(byref(self.h), m, data.dim, _ip(self._keep[0]), _ip(self._keep[1]), _dp(self._keep[2])),
              "HConeProcData/HConePresolveData")

    def types(self) -> np.ndarray:
        t = np.zeros(self.m + 1, dtype=np.int32)
        check(self.lib.hdsdpcu_cone_gettypes(self.h, _ip(t)), "gettypes")
        return t

    def set_start(self, r: float):
        self.lib.hdsdpcu_cone_setstart(self.h, float(r))

    def reduce_resi(self, r: float):
        self.lib.hdsdpcu_cone_reduceresi(self.h, float(r))

    def set_perturb(self, p: float):
        self.lib.hdsdpcu_cone_setperturb(self.h, float(p))

    def scal(self, s: float):
        check(self.lib.hdsdpcu_cone_scal(self.h, float(s)), "HConeScalByConstant")

    def update(self, tau: float, y: np.ndarray):
        y = np.ascontiguousarray(y, dtype=np.float64)
        check(self.lib.hdsdpcu_cone_update(self.h, float(tau), _dp(y)), "HConeUpdate")

    def update_buffer(self, cC, aScal, a, eye, which):
        a = np.ascontiguousarray(a, dtype=np.float64)
        check(self.lib.hdsdpcu_cone_updatebuffer(self.h, float(cC), float(aScal), _dp(a), float(eye), int(which)), "UpdateBuffer")

    def interior_check(self, tau: float, y: np.ndarray) -> bool:
        y = np.ascontiguousarray(y, dtype=np.float64)
        flag = c_int(0)
        check(self.lib.hdsdpcu_cone_interiorcheck(self.h, float(tau), _dp(y), byref(flag)), "HConeCheckIsInterior")
        return bool(flag.value)

    def interior_check_expert(self, cC, aScal, a, eye, which) -> bool:
        a = np.ascontiguousarray(a, dtype=np.float64)
        flag = c_int(0)
        check(self.lib.hdsdpcu_cone_interiorcheckexpert(self.h, float(cC), float(aScal), _dp(a), float(eye), int(which), byref(flag)),
              "HConeCheckIsInteriorExpert")
        return bool(flag.value)

    def factorize(self, which=BUFFER_DUALVAR) -> bool:
        flag = c_int(0)
        check(self.lib.hdsdpcu_cone_factorize(self.h, int(which), byref(flag)), "HFpLinsysPsdCheck")
        return bool(flag.value)

    def ratio_test(self, dtau: float, dy: np.ndarray, ada_ratio: float, which=BUFFER_DUALVAR) -> float:
        """HConeRatioTest: largest alpha with S + alpha dS >= 0 (device Lanczos)."""
        dy = np.ascontiguousarray(dy, dtype=np.float64)
        step = c_double(0.0)
        check(self.lib.hdsdpcu_cone_ratiotest(self.h, float(dtau), _dp(dy), float(ada_ratio), int(which), byref(step)), "HConeRatioTest")
        return step.value

    def build_primal_xsx(self, X: np.ndarray, XSX: np.ndarray, dual_mat: int = 1) -> np.ndarray:
        """HConeBuildPrimalXSXDirection: returns XSX + X S X."""
        X = np.asfortranarray(X, dtype=np.float64); out = np.asfortranarray(XSX, dtype=np.float64).copy(order="F")
        check(self.lib.hdsdpcu_cone_buildprimalxsx(self.h, _dp(X), _dp(out), int(dual_mat)), "HConeBuildPrimalXSXDirection")
        return out

    def get_primal(self, mu: float, y: np.ndarray, dy: np.ndarray):
        """HConeGetPrimal: X = mu (S^-1 + S^-1 dS S^-1) with S = C - A'y, dS = A'dy; None if S is not positive definite."""
        y = np.ascontiguousarray(y, dtype=np.float64); dy = np.ascontiguousarray(dy, dtype=np.float64)
        X = np.zeros((self.n, self.n), order="F")
        ok = c_int(0)
        check(self.lib.hdsdpcu_cone_getprimal(self.h, float(mu), _dp(y), _dp(dy), _dp(X), byref(ok)), "HConeGetPrimal")
        return X if ok.value else None

    def lanczos_multiply(self, x: np.ndarray, which=BUFFER_DUALVAR) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        check(self.lib.hdsdpcu_cone_lanczosmultiply(self.h, int(which), _dp(x), _dp(y)), "LanczosMultiply")
        return y

    def get_log_barrier(self, tau: float, y: Optional[np.ndarray], which=BUFFER_DUALVAR) -> float:
        ld = c_double(0.0)
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        check(self.lib.hdsdpcu_cone_getbarrier(self.h, float(tau), _dp(yy), int(which), byref(ld)), "HConeGetLogBarrier")
        return ld.value

    def add_step_and_check(self, step: float, which: int) -> bool:
        flag = c_int(0)
        check(self.lib.hdsdpcu_cone_addstepandcheck(self.h, float(step), int(which), byref(flag)), "HConeAddStepToBufferAndCheck")
        return bool(flag.value)

    def get_buffer(self, which=BUFFER_DUALVAR) -> np.ndarray:
        out = np.zeros((self.n, self.n), order="F")
        check(self.lib.hdsdpcu_cone_getbuffer(self.h, int(which), _dp(out)), "getbuffer")
        return out

    def get_sinv(self) -> np.ndarray:
        out = np.zeros((self.n, self.n), order="F")
        check(self.lib.hdsdpcu_cone_getsinv(self.h, _dp(out)), "getsinv")
        return out

    def get_factor_diag(self, which=BUFFER_DUALVAR) -> np.ndarray:
        d = np.zeros(self.n)
        check(self.lib.hdsdpcu_cone_getfactordiag(self.h, int(which), _dp(d)), "HFpLinsysGetDiag")
        return d

    def close(self):
        if self.h:
            self.lib.hdsdpcu_cone_destroy(byref(self.h))
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LPCone:
    """Device twin of the LP cone's Schur contribution (reference interface/hdsdp_conic_lp.c:254-330)."""

    def __init__(self, data: ConeData, m: int):
        assert data.kind == "lp"
        self.lib = _lib.require_gpu()
        self.m, self.ncol = m, data.dim
        self.h = c_void_p()
        beg = np.ascontiguousarray(data.beg, dtype=np.int32)
        idx = np.ascontiguousarray(data.idx, dtype=np.int32)
        elem = np.ascontiguousarray(data.elem, dtype=np.float64)
        check(self.lib.hdsdpcu_lp_create(byref(self.h), m, data.dim, _ip(beg), _ip(idx), _dp(elem)), "LPConeProcData")
        # host-side slack s = c*tau - A^T y (O(nnz) work stays on the host as in the reference)
        self.obj = np.zeros(self.ncol)
        for e in range(beg[0], beg[1]):
            self.obj[idx[e]] = elem[e]
        self._beg, self._idx, self._elem = beg, idx, elem
        self.dual_residual = 0.0

    def slack(self, tau: float, y: np.ndarray) -> np.ndarray:
        """s = tau c - Rd - A^T y (reference LPConeUpdateImpl, hdsdp_conic_lp.c); O(nnz) host work, vectorised."""
        if not hasattr(self, "_con"):
            lo, hi = int(self._beg[1]), int(self._beg[self.m + 1])
            self._con = np.repeat(np.arange(self.m), np.diff(self._beg[1:self.m + 2]))
            self._cidx, self._cval = self._idx[lo:hi].astype(np.int64), self._elem[lo:hi]
        s = tau * self.obj - self.dual_residual
        s -= np.bincount(self._cidx, weights=self._cval * np.asarray(y)[self._con], minlength=self.ncol)
        return s

    def close(self):
        if self.h:
            self.lib.hdsdpcu_lp_destroy(byref(self.h))
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KKT:
    """hdsdp_kkt twin: M lives in HBM from HKKTBuildUp to the last HKKTSolve."""

    def __init__(self, m: int, cones: List[SDPCone]):
        self.lib = _lib.require_gpu()
        self.m = m
        self.cones = cones
        self.h = c_void_p()
        check(self.lib.hdsdpcu_kkt_create(byref(self.h), m), "HKKTCreate/HKKTInit")
        for c in cones:
            check(self.lib.hdsdpcu_kkt_addcone(self.h, c.h), "HKKTInit(cone)")
        self._primal_keep = None

    def set_shard(self, rank: int, nranks: int):
        check(self.lib.hdsdpcu_kkt_setshard(self.h, rank, nranks), "setshard")

    def dist_init(self, rank: int, nranks: int, block: int = 512, allgather=None):
        """Multi-GPU Schur matrix (one process per GPU): 1-D block-cyclic column ownership, peer-memory panel exchange.
        `allgather(bytes) -> list[bytes]` gathers the IPC blobs in rank order (torch.distributed, MPI, ...)."""
        check(self.lib.hdsdpcu_kkt_dist_init(self.h, rank, nranks, block), "kkt_dist_init")
        self.rank, self.nranks, self.block = rank, nranks, block
        if nranks == 1:
            return
        import ctypes
        nbytes = self.lib.hdsdpcu_dist_blob_bytes()
        blob = ctypes.create_string_buffer(nbytes)
        check(self.lib.hdsdpcu_kkt_dist_export(self.h, blob), "kkt_dist_export")
        blobs = allgather(blob.raw)
        assert len(blobs) == nranks and all(len(b) == nbytes for b in blobs)
        allb = ctypes.create_string_buffer(b"".join(blobs), nbytes * nranks)
        check(self.lib.hdsdpcu_kkt_dist_connect(self.h, allb), "kkt_dist_connect")

    def build_up(self, type_kkt: int = KKT_TYPE_INFEASIBLE):
        check(self.lib.hdsdpcu_kkt_buildup(self.h, int(type_kkt)), "HKKTBuildUp")

    def build_up_extra_bound(self, diag_add, asinv_add, asinvrd_add=None, type_kkt=KKT_TYPE_INFEASIBLE):
        d = None if diag_add is None else np.ascontiguousarray(diag_add, dtype=np.float64)
        a = None if asinv_add is None else np.ascontiguousarray(asinv_add, dtype=np.float64)
        r = None if asinvrd_add is None else np.ascontiguousarray(asinvrd_add, dtype=np.float64)
        check(self.lib.hdsdpcu_kkt_buildupextra_bound(self.h, _dp(d), _dp(a), _dp(r), int(type_kkt)), "HKKTBuildUpExtraCone(bound)")

    def build_up_extra_lp(self, lp: LPCone, col_dual_inverse: np.ndarray, dual_residual: float, type_kkt=KKT_TYPE_INFEASIBLE):
        s = np.ascontiguousarray(col_dual_inverse, dtype=np.float64)
        check(self.lib.hdsdpcu_kkt_buildupextra_lp(self.h, lp.h, _dp(s), float(dual_residual), int(type_kkt)), "HKKTBuildUpExtraCone(lp)")

    def regularize(self, reg: float):
        check(self.lib.hdsdpcu_kkt_regularize(self.h, float(reg)), "HKKTRegularize")

    def export(self):
        a = np.zeros(self.m); ard = np.zeros(self.m); ac = np.zeros(self.m)
        s = [c_double(0.0) for _ in range(4)]
        check(self.lib.hdsdpcu_kkt_export(self.h, _dp(a), _dp(ard), _dp(ac), byref(s[0]), byref(s[1]), byref(s[2]), byref(s[3])), "HKKTExport")
        return {"dASinvVec": a, "dASinvRdSinvVec": ard, "dASinvCSinvVec": ac, "dCSinvCSinv": s[0].value, "dCSinv": s[1].value,
                "dCSinvRdSinv": s[2].value, "dTraceSinv": s[3].value}

    def factorize(self) -> int:
        return self.lib.hdsdpcu_kkt_factorize(self.h)

    def solve(self, rhs: np.ndarray) -> np.ndarray:
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        if rhs.ndim == 1:
            out = np.zeros(self.m)
            check(self.lib.hdsdpcu_kkt_solve(self.h, _dp(rhs), _dp(out)), "HKKTSolve")
            return out
        r = np.asfortranarray(rhs)
        out = np.zeros_like(r, order="F")
        check(self.lib.hdsdpcu_kkt_solve_many(self.h, r.shape[1], _dp(r), _dp(out)), "HKKTSolve")
        return out

    def symv(self, x: np.ndarray) -> np.ndarray:
        """y = M x on the device-resident Schur matrix."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.m)
        check(self.lib.hdsdpcu_kkt_symv(self.h, _dp(x), _dp(y)), "symv")
        return y

    def set_solver(self, mode: int):
        """0: direct Cholesky (default); 1: the reference's policy, Jacobi-PCG first and Cholesky after its first failure."""
        check(self.lib.hdsdpcu_kkt_set_solver(self.h, int(mode)), "set_solver")

    def pcg_status(self):
        a, b, c, d = c_int(0), c_int(0), c_int(0), c_int(0)
        self.lib.hdsdpcu_kkt_pcg_status(self.h, byref(a), byref(b), byref(c), byref(d))
        return {"use_jacobi": a.value, "last_iterations": b.value, "n_solves": c.value, "n_fallbacks": d.value}

    def solve_status(self):
        r = c_double(0.0); s = c_int(0)
        self.lib.hdsdpcu_kkt_solve_status(self.h, byref(r), byref(s))
        return r.value, s.value

    def register_psdp(self, Xs: List[np.ndarray]):
        self._primal_keep = [np.asfortranarray(X, dtype=np.float64) for X in Xs]
        arr = (c_double_p * len(Xs))(*[_dp(X) for X in self._primal_keep])
        self._primal_arr = arr
        self.lib.hdsdpcu_kkt_registerpsdp(self.h, len(Xs), arr)

    def get_matrix(self) -> np.ndarray:
        M = np.zeros((self.m, self.m), order="F")
        check(self.lib.hdsdpcu_kkt_getmatrix(self.h, _dp(M)), "getmatrix")
        return M

    def close(self):
        if self.h:
            self.lib.hdsdpcu_kkt_destroy(byref(self.h))
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_problem(prob: Problem):
    """Create device cones + KKT for a Problem.  Returns (sdp_cones, lp_cones, kkt)."""
    sdp = [SDPCone(c, prob.m) for c in prob.cones if c.kind == "sdp"]
    lps = [LPCone(c, prob.m) for c in prob.cones if c.kind == "lp"]
    kkt = KKT(prob.m, sdp)
    return sdp, lps, kkt
