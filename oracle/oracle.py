"""ctypes wrapper of oracle/liboracle.so (oracle/hdsdp_oracle.c, the plain-C CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_double, c_int, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
SRC = os.path.join(_HERE, "hdsdp_oracle.c")
c_double_p = POINTER(c_double)
c_int_p = POINTER(c_int)
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        l = ctypes.CDLL(LIB)
        l.orc_cone_create.restype = c_void_p
        l.orc_cone_create.argtypes = [c_int, c_int, c_int_p, c_int_p, c_double_p]
        l.orc_cone_destroy.argtypes = [c_void_p]
        l.orc_cone_is_dense_type.argtypes = [c_void_p]
        l.orc_cone_types.argtypes = [c_void_p, c_int_p]
        l.orc_cone_strategies.argtypes = [c_void_p, c_int_p, c_int_p]
        l.orc_cone_r1sign.restype = c_double
        l.orc_cone_r1sign.argtypes = [c_void_p, c_int]
        l.orc_cone_set_resi.argtypes = [c_void_p, c_double]
        l.orc_cone_set_perturb.argtypes = [c_void_p, c_double]
        l.orc_cone_scal_obj.argtypes = [c_void_p, c_double]
        l.orc_cone_update_buffer.argtypes = [c_void_p, c_double, c_double, c_double_p, c_double, c_int, c_double_p]
        l.orc_potrf.argtypes = [c_int, c_double_p, c_double_p]
        l.orc_invert.argtypes = [c_int, c_double_p, c_double_p]
        l.orc_cone_set_point.argtypes = [c_void_p, c_double_p, c_double, c_double_p]
        for nm in ("S", "L", "Sinv"):
            getattr(l, f"orc_cone_get_{nm}").argtypes = [c_void_p, c_double_p]
        l.orc_cone_build_schur.argtypes = [c_void_p, c_int, c_int, c_double_p, c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        l.orc_lp_schur.argtypes = [c_int, c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_double, c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        l.orc_bound_schur.argtypes = [c_int, c_double_p, c_double_p, c_int, c_double_p, c_double_p]
        l.orc_regularize.argtypes = [c_int, c_double_p, c_double]
        l.orc_kkt_solve.argtypes = [c_int, c_double_p, c_double_p, c_double_p]
        l.orc_iteration.argtypes = [c_void_p, c_double_p, c_double, c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double, c_double_p, c_double_p]
        _lib = l
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


class OracleCone:
    def __init__(self, cone, m):
        self.l = lib()
        self.m, self.n = m, cone.dim
        self._keep = (np.ascontiguousarray(cone.beg, dtype=np.int32), np.ascontiguousarray(cone.idx, dtype=np.int32),
                      np.ascontiguousarray(cone.elem, dtype=np.float64))
        self.h = self.l.orc_cone_create(m, cone.dim, _ip(self._keep[0]), _ip(self._keep[1]), _dp(self._keep[2]))

    def types(self):
        t = np.zeros(self.m + 1, dtype=np.int32); self.l.orc_cone_types(self.h, _ip(t)); return t

    def strategies(self):
        p = np.zeros(self.m, dtype=np.int32); s = np.zeros(self.m, dtype=np.int32)
        self.l.orc_cone_strategies(self.h, _ip(p), _ip(s)); return p, s

    def is_dense_type(self):
        return bool(self.l.orc_cone_is_dense_type(self.h))

    def r1sign(self, i):
        return self.l.orc_cone_r1sign(self.h, i)

    def set_resi(self, rd):
        self.l.orc_cone_set_resi(self.h, float(rd))

    def set_point(self, y, tau):
        y = np.ascontiguousarray(y, dtype=np.float64); ld = c_double(0.0)
        ok = self.l.orc_cone_set_point(self.h, _dp(y), float(tau), byref(ld))
        return bool(ok), ld.value

    def get(self, which):
        out = np.zeros((self.n, self.n), order="F")
        getattr(self.l, f"orc_cone_get_{which}")(self.h, _dp(out)); return out

    def update_buffer(self, cC, aScal, a, eye, is_step=False):
        a = np.ascontiguousarray(a, dtype=np.float64); T = np.zeros((self.n, self.n), order="F")
        self.l.orc_cone_update_buffer(self.h, float(cC), float(aScal), _dp(a), float(eye), int(is_step), _dp(T)); return T

    def build_schur(self, kkt, type_kkt=0, strategy=-1, primal_x=None):
        X = None if primal_x is None else np.asfortranarray(primal_x, dtype=np.float64)
        rc = self.l.orc_cone_build_schur(self.h, int(type_kkt), int(strategy), _dp(X), kkt.m, _dp(kkt.M), _dp(kkt.asinv), _dp(kkt.asinvrd),
                                         _dp(kkt.asinvc), _dp(kkt.scal))
        if rc != 0:
            raise RuntimeError("oracle build_schur failed")

    def close(self):
        if self.h:
            self.l.orc_cone_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleKKT:
    """Host arrays with the layout of hdsdp_kkt (dense M, lower)."""

    def __init__(self, m):
        self.m = m
        self.M = np.zeros((m, m), order="F"); self.asinv = np.zeros(m); self.asinvrd = np.zeros(m); self.asinvc = np.zeros(m)
        self.scal = np.zeros(4)  # dCSinvCSinv, dCSinv, dCSinvRdSinv, dTraceSinv

    def clean(self, type_kkt=0):  # hdsdp_schur.c:141-165
        self.asinv[:] = 0; self.asinvrd[:] = 0
        if type_kkt == 2:
            self.asinvc[:] = 0; self.scal[:3] = 0
        self.scal[3] = 0
        if type_kkt in (0, 2, 3):
            self.M[:] = 0

    def add_lp(self, cone, s, rd, type_kkt=0):
        beg = np.ascontiguousarray(cone.beg, dtype=np.int32); idx = np.ascontiguousarray(cone.idx, dtype=np.int32)
        elem = np.ascontiguousarray(cone.elem, dtype=np.float64); s = np.ascontiguousarray(s, dtype=np.float64)
        lib().orc_lp_schur(self.m, cone.dim, _ip(beg), _ip(idx), _dp(elem), _dp(s), float(rd), int(type_kkt), _dp(self.M), _dp(self.asinv),
                           _dp(self.asinvrd), _dp(self.asinvc), _dp(self.scal))

    def regularize(self, reg):
        lib().orc_regularize(self.m, _dp(self.M), float(reg))

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64); x = np.zeros(self.m)
        info = lib().orc_kkt_solve(self.m, _dp(self.M), _dp(rhs), _dp(x))
        if info != 0:
            raise RuntimeError(f"oracle: M not positive definite (info {info})")
        return x

    def vectors(self):
        return {"dASinvVec": self.asinv, "dASinvRdSinvVec": self.asinvrd, "dASinvCSinvVec": self.asinvc, "dCSinvCSinv": self.scal[0],
                "dCSinv": self.scal[1], "dCSinvRdSinv": self.scal[2], "dTraceSinv": self.scal[3]}


def lp_slack(cone, tau, y, rd):
    """LP slack s = tau*c - A'y - Rd (reference interface/hdsdp_conic_lp.c:45-80)."""
    s = np.zeros(cone.dim)
    for e in range(cone.beg[0], cone.beg[1]):
        s[cone.idx[e]] += tau * cone.elem[e]
    for k in range(len(y)):
        lo, hi = cone.beg[k + 1], cone.beg[k + 2]
        if hi > lo:
            np.subtract.at(s, cone.idx[lo:hi], y[k] * cone.elem[lo:hi])
    return s - rd
