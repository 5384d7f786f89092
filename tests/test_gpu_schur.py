"""GPU parity of the cone + Schur path (B2) through the C ABI against the reference's own outputs.

The expected values are the committed golden fixtures (tests/golden/*.npz, produced by the UNMODIFIED
reference through tests/golden/make_golden.py).  Tolerances (BASELINE.json north_star):
  * S entries: exact up to summation order (1e-13 relative),
  * Schur entries and side vectors: 1e-10 relative.  "Relative" is |a-b| <= 1e-10 * max(|b|, 1e-3 max|M|):
    entries that are signed sums (M5) cancel, a pure per-entry relative test is meaningless there
    ("parity metric"; the reference's own cross-check uses |a-b|/(|a|+1e-4) < 1e-8),
  * solves of M: 1e-6 relative to the reference's PCG solution (its own tolerance is ~1e-12 absolute residual).
"""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def close(got, ref, scale=None, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    s = np.abs(ref).max() if scale is None else scale
    tol = rtol * np.maximum(np.abs(ref), 1e-3 * s)
    err = np.abs(got - ref)
    bad = err > tol
    if bad.any():
        k = np.argmax(err / np.maximum(tol, 1e-300))
        return False, f"max violation at flat index {k}: got {got.flat[k]!r} ref {ref.flat[k]!r} (|ref|max {s:.3e}, {int(bad.sum())} bad)"
    return True, ""


def lp_terms(lp, tau, y, rd):
    lp.dual_residual = rd
    s = lp.slack(tau, y)
    return 1.0 / s


def run_point(prob, z, p, sdp, lps, kkt):
    from hdsdp_b200 import api
    y, tau, rd = z[p + "y"], float(z[p + "tau"]), float(z[p + "rd"])
    logdet = 0.0
    k_sdp = 0
    for k, cone in enumerate(prob.cones):
        if cone.kind != "sdp":
            continue
        c = sdp[k_sdp]; k_sdp += 1
        c.set_start(rd)
        c.update(tau, y)
        S = np.tril(c.get_buffer(api.BUFFER_DUALVAR))
        ok, msg = close(S, z[p + f"S{k}"], rtol=1e-13)
        assert ok, f"S cone {k}: {msg}"
        logdet += c.get_log_barrier(tau, y)
        ok, msg = close(c.get_factor_diag(), z[p + f"Ldiag{k}"])
        assert ok, f"diag(L) cone {k}: {msg}"
    for lp in lps:
        logdet += np.log(1.0 / lp_terms(lp, tau, y, rd)).sum()
    assert abs(logdet - float(z[p + "logdet"])) <= 1e-10 * abs(float(z[p + "logdet"]))

    for tname, t in (("inf", api.KKT_TYPE_INFEASIBLE), ("hsd", api.KKT_TYPE_HOMOGENEOUS), ("cor", api.KKT_TYPE_CORRECTOR)):
        if t == api.KKT_TYPE_CORRECTOR:
            kkt.build_up(api.KKT_TYPE_INFEASIBLE)
        kkt.build_up(t)
        for lp in lps:
            sinv = lp_terms(lp, tau, y, rd)
            kkt.build_up_extra_lp(lp, sinv, rd, t)
        got = kkt.export()
        if t != api.KKT_TYPE_CORRECTOR:
            M = np.tril(kkt.get_matrix())
            Mref = z[p + f"M_{tname}"]
            ok, msg = close(M, Mref)
            assert ok, f"{prob.name} {p}{tname} M: {msg}"
        for key in ("dASinvVec", "dASinvRdSinvVec"):
            ok, msg = close(got[key], z[p + f"{tname}_{key}"])
            assert ok, f"{prob.name} {p}{tname} {key}: {msg}"
        if t != api.KKT_TYPE_CORRECTOR:
            ref_tr = float(z[p + f"{tname}_dTraceSinv"])
            assert abs(got["dTraceSinv"] - ref_tr) <= RTOL * max(abs(ref_tr), 1e-300), f"{tname} dTraceSinv {got['dTraceSinv']} vs {ref_tr}"
        if t == api.KKT_TYPE_HOMOGENEOUS:   # with an LP cone too: its HSD terms are part of the device LP kernel
            ok, msg = close(got["dASinvCSinvVec"], z[p + "hsd_dASinvCSinvVec"])
            assert ok, f"{prob.name} {p} dASinvCSinvVec: {msg}"
            for key in ("dCSinv", "dCSinvCSinv", "dCSinvRdSinv"):
                r = float(z[p + f"hsd_{key}"])
                assert abs(got[key] - r) <= RTOL * max(abs(r), 1e-300), f"{prob.name} {p} {key}: {got[key]!r} vs {r!r}"
    if len(sdp) == 1 and (p + "Sinv") in z:
        ok, msg = close(sdp[0].get_sinv(), z[p + "Sinv"])
        assert ok, f"S^-1: {msg}"
    # factor + solve
    if (p + "sol_asinv") in z:
        kkt.build_up(api.KKT_TYPE_INFEASIBLE)
        for lp in lps:
            kkt.build_up_extra_lp(lp, lp_terms(lp, tau, y, rd), rd, api.KKT_TYPE_INFEASIBLE)
        rhs = kkt.export()["dASinvVec"]
        assert kkt.factorize() == 0
        x = kkt.solve(rhs)
        ref = z[p + "sol_asinv"]
        assert np.abs(x - ref).max() <= 1e-6 * np.abs(ref).max(), f"solve: {np.abs(x - ref).max()} vs scale {np.abs(ref).max()}"
        # and exactly: residual of our own solve against our own M
        Mfull = kkt.get_matrix(); Mfull = np.tril(Mfull) + np.tril(Mfull, -1).T
        assert np.abs(Mfull @ x - rhs).max() <= 1e-10 * max(np.abs(rhs).max(), np.abs(Mfull).max() * np.abs(x).max())


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_parity(name):
    from hdsdp_b200 import api
    prob, z = load_golden(name)
    sdp, lps, kkt = api.build_problem(prob)
    # classification must match the reference's presolve (decides storage class)
    ks = 0
    for k, cone in enumerate(prob.cones):
        if cone.kind != "sdp":
            continue
        assert np.array_equal(sdp[ks].types(), z[f"cone{k}_types"]), f"cone {k} classification differs"
        ks += 1
    assert int(z["npoints"]) >= 1
    for i in range(int(z["npoints"])):
        run_point(prob, z, f"pt{i}_", sdp, lps, kkt)
    kkt.close()
    for c in sdp:
        c.close()


def test_regularize_and_bound_terms():
    from hdsdp_b200 import api
    prob, z = load_golden("maxcut40")
    sdp, lps, kkt = api.build_problem(prob)
    y, tau, rd = z["pt0_y"], float(z["pt0_tau"]), float(z["pt0_rd"])
    sdp[0].set_start(rd); sdp[0].update(tau, y); assert sdp[0].factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    M0 = kkt.get_matrix(); v0 = kkt.export()
    d = np.linspace(0.1, 0.2, prob.m); a = np.linspace(-1, 1, prob.m)
    kkt.build_up_extra_bound(d, a, None, api.KKT_TYPE_INFEASIBLE)
    M1 = kkt.get_matrix(); v1 = kkt.export()
    np.testing.assert_allclose(np.diag(M1), np.diag(M0) + d, rtol=1e-15)
    np.testing.assert_allclose(v1["dASinvVec"], v0["dASinvVec"] + a, rtol=1e-15)
    # HKKTRegularize (hdsdp_schur.c:348-373)
    kkt.regularize(1e-6)
    M2 = kkt.get_matrix()
    reg = min(1e-6 * np.diag(M1).min(), 1e-5)
    reg = 0.0 if reg < 1e-14 else reg
    np.testing.assert_allclose(np.diag(M2), np.diag(M1) + reg, rtol=1e-15)
    assert np.array_equal(np.tril(M2, -1), np.tril(M1, -1))


def test_primal_type_uses_registered_x():
    """KKT_TYPE_PRIMAL: 'S^-1' is the registered primal matrix (hdsdp_conic_sdp.c:1745-1756)."""
    from hdsdp_b200 import api
    from hdsdp_b200.problem import cone_to_dense
    prob, z = load_golden("theta30")
    sdp, lps, kkt = api.build_problem(prob)
    n = prob.cones[0].dim
    rs = np.random.RandomState(0)
    G = rs.standard_normal((n, n)); X = G @ G.T + np.eye(n)
    sdp[0].set_start(0.0)
    kkt.register_psdp([X])
    kkt.build_up(api.KKT_TYPE_PRIMAL)
    M = np.tril(kkt.get_matrix())
    A = [cone_to_dense(prob.cones[0], i + 1) for i in range(prob.m)]
    XA = [X @ a for a in A]
    ref = np.array([[np.trace(XA[i] @ XA[j]) if i >= j else 0.0 for j in range(prob.m)] for i in range(prob.m)])
    ok, msg = close(M, ref)
    assert ok, msg


def test_interior_checks_and_steps():
    from hdsdp_b200 import api
    prob, z = load_golden("mcp100")
    sdp, lps, kkt = api.build_problem(prob)
    c = sdp[0]
    y = z["pt0_y"]
    c.set_start(-1e3)
    assert c.interior_check(1.0, y) is True
    c.set_start(+1e3)   # S = -1e3 I + C: not PSD
    assert c.interior_check(1.0, y) is False
    c.set_start(-1e3)
    c.update(1.0, y)
    # dS = I (via expert form on the step buffer), S + alpha dS
    c.update_buffer(0.0, 0.0, np.zeros(prob.m), 1.0, api.BUFFER_DUALSTEP)
    assert c.add_step_and_check(-2e3, api.BUFFER_DUALCHECK) is False
    assert c.add_step_and_check(+5.0, api.BUFFER_DUALCHECK) is True
    S0 = c.get_buffer(api.BUFFER_DUALVAR)
    Sc = c.get_buffer(api.BUFFER_DUALCHECK)
    np.testing.assert_allclose(np.diag(Sc), np.diag(S0) + 5.0, rtol=1e-15)


def test_kkt_switches_to_ldl_when_cholesky_fails():
    """Reference HFpLinsysNumeric (hdsdp_linsolver.c:2030-2039): dpotrf of M fails -> permanent switch to the indefinite
    back-end, HKKTFactorize still returns OK and HKKTSolve solves the (slightly indefinite) system."""
    import ctypes
    from hdsdp_b200 import _lib, api
    prob, z = load_golden("theta30")
    sdp, lps, kkt = api.build_problem(prob)
    p = "pt1_"
    c = sdp[0]
    c.set_start(float(z[p + "rd"])); c.update(float(z[p + "tau"]), z[p + "y"])
    assert c.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    M = np.tril(kkt.get_matrix()); M = M + np.tril(M, -1).T
    lam = np.linalg.eigvalsh(M)
    shift = 0.5 * (lam[2] + lam[3])                       # push three eigenvalues below zero
    assert lam[3] - lam[2] > 1e-6 * lam[-1], "test point must have a spectral gap"
    kkt.build_up_extra_bound(-shift * np.ones(prob.m), np.zeros(prob.m))
    assert kkt.factorize() == 0
    isldl, neg, pert = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.lib().hdsdpcu_kkt_ldl_status(kkt.h, ctypes.byref(isldl), ctypes.byref(neg), ctypes.byref(pert))
    assert isldl.value == 1 and neg.value == 3
    b = np.random.RandomState(3).standard_normal(prob.m)
    x = kkt.solve(b)
    Ms = M - shift * np.eye(prob.m)
    assert np.abs(x - np.linalg.solve(Ms, b)).max() <= 1e-8 * np.abs(x).max()
    # the switch is permanent (HFpLinsysSwitchToIndefinite): a positive definite M is now solved by LDL as well
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    assert kkt.factorize() == 0
    x2 = kkt.solve(b)
    assert np.abs(x2 - np.linalg.solve(M, b)).max() <= 1e-9 * np.abs(x2).max()


def test_multiblock_midsize_against_oracle():
    """Config-E shape at a size where the many-row code paths run (split-K S assembly from 5000 dense rank-one rows, chunked packed
    axpy over 200 dense rows, batched dense x dense Gram block with chunked traces, LP cone), against the pinned plain-C oracle:
    S per cone 1e-13, Schur matrix / side vectors 1e-10 for the INFEASIBLE and HOMOGENEOUS types."""
    from hdsdp_b200 import api, problem
    from oracle import oracle
    m = 5000
    prob = problem.gen_multiblock(m, n1=40, n2=30, ndense=200, nlp=300, seed=5)
    rs = np.random.RandomState(11)
    y = 1e-3 * rs.uniform(-1, 1, m)
    tau, rd = 1.0, -2e3
    sdp, lps, kkt = api.build_problem(prob)
    ocones = [oracle.OracleCone(c, m) for c in prob.cones if c.kind == "sdp"]
    for c, oc in zip(sdp, ocones):
        c.set_start(rd); c.update(tau, y)
        assert c.factorize()
        oc.set_resi(rd)
        ok, _ = oc.set_point(y, tau)
        assert ok
        S, So = np.tril(c.get_buffer(api.BUFFER_DUALVAR)), np.tril(oc.get("S"))
        assert np.abs(S - So).max() <= 1e-13 * np.abs(So).max()
    lp = lps[0]
    lp.dual_residual = rd
    s_lp = lp.slack(tau, y)
    lpc = [c for c in prob.cones if c.kind == "lp"][0]
    assert np.abs(s_lp - oracle.lp_slack(lpc, tau, y, rd)).max() <= 1e-12 * np.abs(s_lp).max()
    for type_kkt in (api.KKT_TYPE_INFEASIBLE, api.KKT_TYPE_HOMOGENEOUS):
        kkt.build_up(type_kkt)
        kkt.build_up_extra_lp(lp, 1.0 / s_lp, rd, type_kkt)
        ok_ = oracle.OracleKKT(m)
        ok_.clean(type_kkt)
        for oc in ocones:
            oc.build_schur(ok_, type_kkt)
        ok_.add_lp(lpc, s_lp, rd, type_kkt)
        M, Mo = np.tril(kkt.get_matrix()), np.tril(ok_.M)
        scale = np.abs(Mo).max()
        assert (np.abs(M - Mo) <= 1e-10 * np.maximum(np.abs(Mo), 1e-3 * scale)).all(), np.abs(M - Mo).max() / scale
        v, vo = kkt.export(), ok_.vectors()
        names = ["dASinvVec", "dASinvRdSinvVec"] + (["dASinvCSinvVec"] if type_kkt == api.KKT_TYPE_HOMOGENEOUS else [])
        for k in names:
            sc = np.abs(vo[k]).max()
            assert (np.abs(v[k] - vo[k]) <= 1e-10 * np.maximum(np.abs(vo[k]), 1e-3 * sc)).all(), (k, np.abs(v[k] - vo[k]).max() / sc)
    kkt.close()
    for c in sdp:
        c.close()
