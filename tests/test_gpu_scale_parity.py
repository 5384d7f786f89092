"""Parity where the real kernels run: sizes far above one 128 x 128 leaf, against the UNMODIFIED reference
run live on the same box (oracle/_ref through oracle/refdrv.py) and against LAPACK on ill-conditioned inputs.

 (b) theta n = 1500, m = 8001 at bench.py's iterate: the device's S, log det, Schur matrix M (all 32 M lower entries), side vectors
     (INFEASIBLE and HOMOGENEOUS types) and solve against refdrv.RefKKT on the same y.   Gate: 1e-10 relative (north star).
 (c) ill-conditioned factorisations: synthetic SPD matrices with condition 1e10 / 1e12 / 1e14 at n = 1024 and 4096 (the
     recursion multiplies by explicit 128 x 128 leaf inverses, csrc/chol.cu) against LAPACK's dpotrf / dpotri / dpotrs:
     backward errors must stay within a small factor of LAPACK's own;  and S, M at the FINAL iterate of a real IPM solve
     (theta n = 200, m = 3001 through the drop-in build; cond(S) ~ 1e8..1e10), compared with the reference at the same y with the
     tolerance scaled by cond(S) * eps (two backward-stable algorithms agree no better than that).
"""
import os
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def entry_ok(got, ref, rtol, what):
    s = np.abs(ref).max()
    err = np.abs(got - ref)
    tol = rtol * np.maximum(np.abs(ref), 1e-3 * s)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} entries differ, max err {err.max():.3e} (scale {s:.3e}, rtol {rtol:.1e})"


def need_ref():
    from oracle import refdrv
    if not refdrv.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refdrv


def test_theta_n1500_m8001_against_live_reference():
    import bench
    from hdsdp_b200 import api, problem
    refdrv = need_ref()
    n, ne = 1500, 8000
    prob = problem.gen_theta(n, ne, seed=2)
    m = prob.m
    y = bench.theta_point(m, n, 0)
    ref = refdrv.RefKKT(prob)
    ld_ref = ref.set_point(y, bench.TAU, bench.RD)
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(bench.RD)
    cone.update(bench.TAU, y)
    assert cone.factorize()
    # S and its factor
    entry_ok(np.tril(cone.get_buffer(api.BUFFER_DUALVAR)), np.tril(ref.get_S(0)), 1e-13, "S")
    np.testing.assert_allclose(cone.get_factor_diag(), ref.get_Ldiag(0), rtol=1e-10)
    assert abs(cone.get_log_barrier(bench.TAU, None) - ld_ref) <= 1e-10 * abs(ld_ref)
    for type_kkt in (api.KKT_TYPE_INFEASIBLE, api.KKT_TYPE_HOMOGENEOUS):
        ref.build(type_kkt)
        kkt.build_up(type_kkt)
        M = np.tril(kkt.get_matrix())
        entry_ok(M, np.tril(ref.get_M()), 1e-10, f"Schur matrix (type {type_kkt}, m = {m})")
        v, vr = kkt.export(), ref.get_vectors()
        names = ["dASinvVec", "dASinvRdSinvVec"] + (["dASinvCSinvVec"] if type_kkt == api.KKT_TYPE_HOMOGENEOUS else [])
        for k in names:
            entry_ok(v[k], vr[k], 1e-10, k)
        scal = ["dTraceSinv"] + (["dCSinvCSinv", "dCSinv", "dCSinvRdSinv"] if type_kkt == api.KKT_TYPE_HOMOGENEOUS else [])
        for k in scal:
            assert abs(v[k] - vr[k]) <= 1e-10 * max(abs(vr[k]), 1e-300), (k, v[k], vr[k])
    # regularize + factorize + solve (the reference answers with its PCG: 1e-6 is its own accuracy)
    ref.regularize(bench.KKT_REG)
    kkt.regularize(bench.KKT_REG)
    entry_ok(np.diag(kkt.get_matrix()), np.diag(ref.get_M()), 1e-10, "diag(M) after HKKTRegularize")
    assert ref.factorize() == 0 and kkt.factorize() == 0
    xr, x = ref.solve(prob.rhs), kkt.solve(prob.rhs)
    assert np.abs(x - xr).max() <= 1e-6 * np.abs(xr).max()
    # the direct solve must satisfy the system the REFERENCE assembled to (much) better than the PCG answer does
    Mr = ref.get_M()
    Mr = np.tril(Mr) + np.tril(Mr, -1).T
    r_gpu = np.abs(Mr @ x - prob.rhs).max()
    r_ref = np.abs(Mr @ xr - prob.rhs).max()
    assert r_gpu <= max(r_ref, 1e-10 * np.abs(prob.rhs).max()), (r_gpu, r_ref)
    # ratio test at n = 1500 (12 leaves): device Lanczos against the reference's
    dy = np.random.RandomState(5).standard_normal(m) * 0.01
    a_ref = ref.ratio_test(0, 0.0, dy, 0.0, 0)
    a_gpu = cone.ratio_test(0.0, dy, 0.0, api.BUFFER_DUALVAR)
    assert abs(a_gpu - a_ref) <= 1e-3 * abs(a_ref), (a_gpu, a_ref)     # the method's own stopping accuracy
    ref.close(); kkt.close(); cone.close()


def device_factor(ls, n):
    """The Cholesky factor as it sits in HBM (lower triangle; the strict upper part of the diagonal leaves is scratch)."""
    import torch
    lib = ls.lib
    np_ = lib.hdsdpcu_linsys_padded_dim(ls.h)
    ptr = lib.hdsdpcu_linsys_factor_dev(ls.h)

    class _Holder:
        __cuda_array_interface__ = {"shape": (np_, np_), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    lib.hdsdpcu_sync()
    T = torch.as_tensor(_Holder(), device="cuda").cpu().numpy()      # T[j, i] = L[i, j] (column-major buffer)
    return np.tril(T.T[:n, :n])


def spd_with_cond(n, cond, seed):
    rs = np.random.RandomState(seed)
    Q, _ = np.linalg.qr(rs.standard_normal((n, n)))
    ev = np.logspace(0, -np.log10(cond), n)
    A = (Q * ev) @ Q.T
    return 0.5 * (A + A.T)


@pytest.mark.parametrize("n", [1024, 4096])
@pytest.mark.parametrize("cond", [1e10, 1e12, 1e14])
def test_ill_conditioned_factor_inverse_solve_against_lapack(n, cond):
    import scipy.linalg as sla
    from hdsdp_b200.api import DenseLinsys
    A = spd_with_cond(n, cond, seed=n + int(np.log10(cond)))
    nrmA = np.linalg.norm(A, 2)
    eps = np.finfo(float).eps
    try:
        Lr = sla.cholesky(A, lower=True)
    except np.linalg.LinAlgError:
        pytest.skip("not numerically positive definite for LAPACK either (cond * n * eps > 1)")
    margin = (np.diag(Lr).min() ** 2) / (n * eps * nrmA)      # how far LAPACK's smallest pivot is above the rounding level
    ls = DenseLinsys(n)
    rc = ls.numeric(np.asfortranarray(A))
    if rc != 0 and margin < 10.0:
        pytest.skip(f"smallest pivot within 10 n eps ||A|| of zero (margin {margin:.1f}): either outcome is legitimate")
    assert rc == 0, f"a positive definite matrix must factor (LAPACK does, pivot margin {margin:.1f})"
    # factor: L itself is only determined to cond * eps, so the diagonal is a sanity check; the gates are the backward errors
    L_diag = ls.get_diag()
    assert (L_diag > 0).all()
    np.testing.assert_allclose(L_diag, np.diag(Lr), rtol=min(0.5, 100 * cond * eps + 1e-10))
    rs = np.random.RandomState(3)
    b = rs.standard_normal(n)
    x = ls.solve(b)
    xr = sla.cho_solve((Lr, True), b)
    be_gpu = np.linalg.norm(A @ x - b) / (nrmA * np.linalg.norm(x))
    be_ref = np.linalg.norm(A @ xr - b) / (nrmA * np.linalg.norm(xr))
    assert be_gpu <= max(20.0 * be_ref, 100 * eps), f"solve backward error {be_gpu:.2e} vs LAPACK {be_ref:.2e}"
    # the factor itself: L is only determined to cond(A) eps, so compare BACKWARD errors ||L L^T - A|| / ||A|| (this is where
    # multiplying panels by explicit 128 x 128 leaf inverses, csrc/chol.cu trsm_rec, would show if it lost digits)
    Lg = device_factor(ls, n)
    fb_gpu = np.linalg.norm(Lg @ Lg.T - A) / nrmA
    fb_ref = np.linalg.norm(Lr @ Lr.T - A) / nrmA
    assert fb_gpu <= max(20.0 * fb_ref, 100 * eps), f"factorisation backward error {fb_gpu:.2e} vs LAPACK {fb_ref:.2e}"
    # triangular solves (Lanczos operator) with the device's OWN factor: forward error against a long-double substitution,
    # next to what LAPACK's dtrsv achieves on the same L
    nq = min(n, 2048)       # O(n^2) long-double work in Python: the leading 2048 x 2048 block is enough
    Lq = Lg[:nq, :nq].astype(np.longdouble)
    fq = np.zeros(nq, dtype=np.longdouble)
    bq = b[:nq].astype(np.longdouble)
    for i in range(nq):
        fq[i] = (bq[i] - Lq[i, :i] @ fq[:i]) / Lq[i, i]
    f = ls.fsolve(b)[:nq]                                   # forward substitution: the leading block is independent of the rest
    fr = sla.solve_triangular(Lg[:nq, :nq], b[:nq], lower=True)
    fe_gpu = float(np.abs(f - fq).max() / np.abs(fq).max())
    fe_ref = float(np.abs(fr - fq).max() / np.abs(fq).max())
    assert fe_gpu <= max(50.0 * fe_ref, 1e3 * eps), f"forward-substitution forward error {fe_gpu:.2e} vs LAPACK {fe_ref:.2e}"
    g = ls.bsolve(b)
    gr = sla.solve_triangular(Lg, b, lower=True, trans="T")
    assert np.abs(g - gr).max() <= max(1e3 * np.sqrt(cond) * eps, 1e-10) * np.abs(gr).max(), "backward substitution"
    # inverse (dpotri twin): residual ||A X - I|| against LAPACK's own
    X = ls.invert()
    Xr = np.linalg.inv(A)
    res_gpu = np.linalg.norm(A @ X - np.eye(n)) / (nrmA * np.linalg.norm(X))
    res_ref = np.linalg.norm(A @ Xr - np.eye(n)) / (nrmA * np.linalg.norm(Xr))
    assert res_gpu <= max(20.0 * res_ref, 100 * eps * np.sqrt(n)), f"inverse residual {res_gpu:.2e} vs LAPACK {res_ref:.2e}"
    assert np.abs(X - X.T).max() == 0.0
    ls.close()


def test_late_iterate_S_and_M_of_a_real_solve():
    """S and M at the final iterate of a full IPM solve (theta n = 200, m = 3001 through the drop-in build)."""
    sys.path.insert(0, ROOT)
    from tools import fullsolve
    from hdsdp_b200 import api, problem
    refdrv = need_ref()
    if not os.path.exists(fullsolve.INTEGRATED):
        pytest.skip("integration/_build/libhdsdp_integrated.so not built")
    with tempfile.TemporaryDirectory() as tmp:
        ypath = os.path.join(tmp, "y.npy")
        res, log, err = fullsolve.run(("theta", 200, 3000), True, 1, dump_y=ypath)
        assert res is not None and res["retcode"] == 0, log[-2000:] + err[-2000:]
        y = np.load(ypath)
    prob = problem.gen_theta(200, 3000, seed=2)
    m, n = prob.m, 200
    # feasible phase: S = C - A'y (tau = 1, Rd = 0); perturbation as the solver uses at the end (dPerturb ~ 1e-10 .. 0)
    ref = refdrv.RefKKT(prob)
    RD_LATE = -1e-7      # S = 1e-7 I + C - A'y: the final iterate sits on the boundary of the cone up to the solver's perturbation
    ref.set_point(y, 1.0, RD_LATE)
    Sr = ref.get_S(0)
    Sr = np.tril(Sr) + np.tril(Sr, -1).T
    w = np.linalg.eigvalsh(Sr)
    assert w[0] > 0
    cond = w[-1] / w[0]
    assert cond > 1e5, f"expected an ill-conditioned late-iterate S, cond = {cond:.2e}"
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(RD_LATE)
    cone.update(1.0, y)
    assert cone.factorize()
    eps = np.finfo(float).eps
    tol = max(1e-10, 50.0 * cond * eps)
    ref.build(api.KKT_TYPE_INFEASIBLE)
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    Mg, Mr = np.tril(kkt.get_matrix()), np.tril(ref.get_M())
    entry_ok(Mg, Mr, tol, f"late-iterate Schur matrix (cond(S) = {cond:.2e})")
    v, vr = kkt.export(), ref.get_vectors()
    entry_ok(v["dASinvVec"], vr["dASinvVec"], tol, "late-iterate dASinvVec")
    # S^-1 against extended-precision-free ground truth: residual of S S^-1 = I
    Si = cone.get_sinv()
    res = np.linalg.norm(Sr @ Si - np.eye(n)) / (np.linalg.norm(Sr, 2) * np.linalg.norm(Si, 2))
    assert res <= 100 * eps * np.sqrt(n), res
    # Cholesky of the late-iterate M (m = 3001, 24 leaves) and solve: backward error against LAPACK's
    import scipy.linalg as sla
    Mfull = Mr + np.tril(Mr, -1).T
    assert kkt.factorize() == 0
    b = prob.rhs + 0.1 * np.random.RandomState(9).standard_normal(m)
    x = kkt.solve(b)
    Mgf = Mg + np.tril(Mg, -1).T
    be_gpu = np.linalg.norm(Mgf @ x - b) / (np.linalg.norm(Mgf, 2) * np.linalg.norm(x))
    xr = sla.cho_solve(sla.cho_factor(Mgf, lower=True), b)
    be_ref = np.linalg.norm(Mgf @ xr - b) / (np.linalg.norm(Mgf, 2) * np.linalg.norm(xr))
    assert be_gpu <= max(20.0 * be_ref, 100 * eps), (be_gpu, be_ref, np.linalg.cond(Mfull))
    ref.close(); kkt.close(); cone.close()


def test_symv_and_refined_ldl_solve_on_a_late_iterate_M():
    """The Schur matrix of a real late iterate made indefinite (three eigenvalues pushed below zero, as the reference's
    "almost indefinite" case): HKKTFactorize switches to LDL^T, every solve is refined on b - M x with the device symv and reports
    its residual; a NaN right-hand side fails instead of returning garbage.  Also: symv against numpy at m = 3001 (24 x 24 tiles)."""
    import ctypes
    sys.path.insert(0, ROOT)
    from tools import fullsolve
    from hdsdp_b200 import _lib, api, problem
    if not os.path.exists(fullsolve.INTEGRATED):
        pytest.skip("integration/_build/libhdsdp_integrated.so not built")
    with tempfile.TemporaryDirectory() as tmp:
        ypath = os.path.join(tmp, "y.npy")
        res, log, err = fullsolve.run(("theta", 200, 3000), True, 1, max_iter=12, dump_y=ypath)   # a mid-solve iterate: mu ~ 1e-3
        assert res is not None, log[-2000:] + err[-2000:]
        y = np.load(ypath)
    prob = problem.gen_theta(200, 3000, seed=2)
    m = prob.m
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(-1e-3)
    cone.update(1.0, y)
    assert cone.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    M = np.tril(kkt.get_matrix()); M = M + np.tril(M, -1).T
    rs = np.random.RandomState(4)
    x = rs.standard_normal(m)
    yv = kkt.symv(x)
    assert np.abs(yv - M @ x).max() <= 1e-13 * np.abs(M).max() * np.abs(x).max() * m ** 0.5
    lam = np.linalg.eigvalsh(M)
    shift = 0.5 * (lam[2] + lam[3])
    kkt.build_up_extra_bound(-shift * np.ones(m), np.zeros(m))
    assert kkt.factorize() == 0
    isldl, neg, pert = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.lib().hdsdpcu_kkt_ldl_status(kkt.h, ctypes.byref(isldl), ctypes.byref(neg), ctypes.byref(pert))
    assert isldl.value == 1 and neg.value == 3
    b = rs.standard_normal(m)
    xs = kkt.solve(b)
    resid, steps = kkt.solve_status()
    Ms = M - shift * np.eye(m)
    true_res = np.abs(Ms @ xs - b).max() / np.abs(b).max()
    assert resid <= 1e-9 and true_res <= 1e-9, (resid, true_res, steps)
    xr = np.linalg.solve(Ms, b)
    assert np.abs(xs - xr).max() <= 1e-6 * np.abs(xr).max() * max(1.0, np.linalg.cond(Ms) * 1e-10)
    bad = b.copy(); bad[0] = np.nan
    out = np.zeros(m)
    rc = _lib.lib().hdsdpcu_kkt_solve(kkt.h, bad.ctypes.data_as(_lib.c_double_p), out.ctypes.data_as(_lib.c_double_p))
    assert rc == 1, "a NaN right-hand side must fail (HFpLinsysSolve, hdsdp_linsolver.c:2096-2098)"
    kkt.close(); cone.close()


def test_nan_solve_switches_cholesky_backend_to_ldl():
    """HFpLinsysSolve's rule (hdsdp_linsolver.c:2088-2103): a NaN in rhs[0] -> "KKT system is unstable. Switch to LDL." and the
    back-end of M stays indefinite from then on."""
    import ctypes
    from hdsdp_b200 import _lib, api, problem
    prob = problem.gen_theta(40, 300, seed=2)
    sdp, lps, kkt = api.build_problem(prob)
    c = sdp[0]
    c.set_start(-2.0)
    y = 0.05 * np.random.RandomState(0).uniform(-1, 1, prob.m); y[0] = -50.0
    c.update(1.0, y)
    assert c.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    assert kkt.factorize() == 0
    isldl = ctypes.c_int(-1)
    _lib.lib().hdsdpcu_kkt_ldl_status(kkt.h, ctypes.byref(isldl), None, None)
    assert isldl.value == 0
    b = np.ones(prob.m); b[0] = np.nan
    out = np.zeros(prob.m)
    rc = _lib.lib().hdsdpcu_kkt_solve(kkt.h, b.ctypes.data_as(_lib.c_double_p), out.ctypes.data_as(_lib.c_double_p))
    assert rc == 1
    _lib.lib().hdsdpcu_kkt_ldl_status(kkt.h, ctypes.byref(isldl), None, None)
    assert isldl.value == 1, "the switch to the indefinite back-end is permanent"
    M = np.tril(kkt.get_matrix()); M = M + np.tril(M, -1).T
    x = kkt.solve(np.ones(prob.m))
    assert np.abs(M @ x - 1.0).max() <= 1e-9
    kkt.close(); c.close()


def test_reference_pcg_policy_on_device():
    """hdsdpcu_kkt_set_solver(1): the reference's default solver of M (Jacobi-PCG, conjGradSolve) with the vectors in HBM must give
    the direct solve's answer to the reference's own tolerance, report its iteration count, and fall back to Cholesky (sticky) when
    CG cannot converge -- here on theta n = 200, m = 3001 at bench.py's iterate and at an ill-conditioned late iterate."""
    import bench
    from hdsdp_b200 import api, problem
    prob = problem.gen_theta(200, 3000, seed=2)
    m = prob.m
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(bench.RD)
    y = bench.theta_point(m, 200, 0)
    cone.update(bench.TAU, y)
    assert cone.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    kkt.regularize(bench.KKT_REG)
    assert kkt.factorize() == 0
    b = prob.rhs + 0.1 * np.random.RandomState(1).standard_normal(m)
    x_direct = kkt.solve(b)
    kkt.set_solver(1)
    assert kkt.factorize() == 0           # Numeric of the iterative back-end: nothing to factor
    x_cg = kkt.solve(b)
    st = kkt.pcg_status()
    assert st["use_jacobi"] == 1 and 1 <= st["last_iterations"] <= 120 and st["n_fallbacks"] == 0, st
    M = np.tril(kkt.get_matrix()); M = M + np.tril(M, -1).T
    assert np.linalg.norm(M @ x_cg - b) <= 1e-10, np.linalg.norm(M @ x_cg - b)      # cgTol = min(absTol, relTol ||b||) <= 1e-12
    assert np.abs(x_cg - x_direct).max() <= 1e-8 * np.abs(x_direct).max()
    # a system Jacobi-CG cannot solve within the reference's limits: pushed towards singularity
    lam = np.linalg.eigvalsh(M)
    kkt.build_up_extra_bound(-(lam[0] * (1.0 - 1e-9)) * np.ones(m), np.zeros(m))
    assert kkt.factorize() == 0
    x2 = kkt.solve(b)
    st = kkt.pcg_status()
    assert st["use_jacobi"] == 0 and st["n_fallbacks"] == 1, st                       # sticky switch to the Cholesky factor
    M2 = M - lam[0] * (1.0 - 1e-9) * np.eye(m)
    assert np.linalg.norm(M2 @ x2 - b) <= 1e-6 * np.linalg.norm(b) * np.linalg.cond(M2) * 1e-10 + 1e-6
    kkt.set_solver(0)
    kkt.close(); cone.close()


def test_primal_type_schur_matrix_against_live_reference():
    """KKT_TYPE_PRIMAL (PSDP refinement, hdsdp_conic_sdp.c:1745-1756: the registered primal X takes the place of S^-1) at a
    multi-leaf size against the REFERENCE's own HKKTBuildUp(KKT_TYPE_PRIMAL), not a numpy restatement."""
    from hdsdp_b200 import api, problem
    refdrv = need_ref()
    n, ne = 300, 2500
    prob = problem.gen_theta(n, ne, seed=2)
    m = prob.m
    rs = np.random.RandomState(21)
    G = rs.standard_normal((n, n))
    X = G @ G.T / n + np.eye(n)
    X = 0.5 * (X + X.T)
    ref = refdrv.RefKKT(prob)
    y = np.zeros(m); y[0] = -(n + 10.0)
    ref.set_point(y, 1.0, -1.0)
    ref.register_primal([X])
    ref.build(api.KKT_TYPE_PRIMAL)
    Mr = np.tril(ref.get_M())
    sdp, lps, kkt = api.build_problem(prob)
    sdp[0].set_start(-1.0)
    sdp[0].update(1.0, y)
    assert sdp[0].factorize()
    kkt.register_psdp([X])
    kkt.build_up(api.KKT_TYPE_PRIMAL)
    entry_ok(np.tril(kkt.get_matrix()), Mr, 1e-10, "PRIMAL-type Schur matrix")
    v, vr = kkt.export(), ref.get_vectors()
    entry_ok(v["dASinvVec"], vr["dASinvVec"], 1e-10, "PRIMAL-type dASinvVec")
    ref.close(); kkt.close(); sdp[0].close()
