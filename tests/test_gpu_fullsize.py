"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run these in seconds).

Operator identity.  M_kl = <A_k, S^-1 A_l S^-1> (+ LP term), so for ANY x
    (M x)_k = sum_cones <A_k, S^-1 X S^-1>  with  X = sum_l x_l A_l     (+ A D^2 A^T x for the LP cone),
which costs O(n^3 + nnz) on the host and never forms M.  With x = M^-1 b from the GPU (assembly -> Cholesky -> solve)
the identity must give back b: one check of the whole hot path at m = 50 000 (config D), n = m = 8000 (config C) and
the multi-block config E.  Config C additionally has the closed form M = S^-1 o S^-1 (reference M2 with unit vectors,
SURVEY appendix A), config D the closed form of E_ij x E_kl pairs, both checked entry-wise at 1e-10.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def packed_rc(n):
    cols = np.repeat(np.arange(n), n - np.arange(n))
    rows = np.concatenate([np.arange(c, n) for c in range(n)])
    return rows, cols


def apply_schur_operator(prob, sinvs, x, lp_d2=None):
    """(M x) through the operator form; sinvs: one dense S^-1 per SDP cone, lp_d2: s^-2 per LP column."""
    m = prob.m
    y = np.zeros(m)
    ks = 0
    for cone in prob.cones:
        beg, idx, elem = cone.beg.astype(np.int64), cone.idx.astype(np.int64), cone.elem
        lo, hi = beg[1], beg[m + 1]
        con = np.repeat(np.arange(m), np.diff(beg[1:m + 2]))
        ii, vv = idx[lo:hi], elem[lo:hi]
        if cone.kind == "sdp":
            n = cone.dim
            R, C = packed_rc(n)
            r, c = R[ii], C[ii]
            X = np.zeros((n, n))
            np.add.at(X, (r, c), x[con] * vv)
            X = X + np.tril(X, -1).T
            Si = sinvs[ks]; ks += 1
            B = Si @ X @ Si
            w = np.where(r == c, 1.0, 2.0)
            y += np.bincount(con, weights=w * vv * B[r, c], minlength=m)
        else:
            t = np.bincount(ii, weights=vv * x[con], minlength=cone.dim)        # A^T x per LP column
            y += np.bincount(con, weights=vv * (lp_d2 * t)[ii], minlength=m)
    return y


def device_matrix(lib, kkt):
    """torch view of the device-resident M (as its transpose: T[j, i] = M[i, j])."""
    import torch
    mp = lib.hdsdpcu_kkt_padded_dim(kkt.h)
    ptr = lib.hdsdpcu_kkt_matrix_dev(kkt.h)

    class _Holder:
        __cuda_array_interface__ = {"shape": (mp, mp), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Holder(), device="cuda"), mp


def test_config_D_theta_m50000_operator_identity_and_entries():
    import torch
    import bench
    from hdsdp_b200 import _lib, api, problem
    lib = _lib.require_gpu()
    n, ne = bench.THETA_N, bench.THETA_EDGES
    prob = problem.gen_theta(n, ne, seed=2)
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(bench.RD)
    y = bench.theta_point(prob.m, n, 0)
    cone.update(bench.TAU, y)
    assert cone.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    v = kkt.export()
    Si = cone.get_sinv()
    # entries: constraint 0 is the identity, constraint k >= 1 is E_ij (one off-diagonal 1): M = 2 (S_ik S_jl + S_il S_jk)
    c = prob.cones[0]
    R, C = packed_rc(n)
    e_r = R[c.idx[c.beg[2]:c.beg[prob.m + 1]]]; e_c = C[c.idx[c.beg[2]:c.beg[prob.m + 1]]]
    assert len(e_r) == prob.m - 1 and (e_r != e_c).all()
    T, mp = device_matrix(lib, kkt)
    rs = np.random.RandomState(0)
    p = rs.randint(1, prob.m, size=200000); q = rs.randint(1, prob.m, size=200000)
    p, q = np.maximum(p, q), np.minimum(p, q)
    got = T[torch.as_tensor(q, device="cuda"), torch.as_tensor(p, device="cuda")].cpu().numpy()   # M[p, q] lower
    i, j, k, l = e_r[p - 1], e_c[p - 1], e_r[q - 1], e_c[q - 1]
    ref = 2.0 * (Si[i, k] * Si[j, l] + Si[i, l] * Si[j, k])
    scale = np.abs(ref).max()
    assert (np.abs(got - ref) <= 1e-10 * np.maximum(np.abs(ref), 1e-3 * scale)).all()
    # first column (identity row, reference M4): M[k, 0] = <E_ij, S^-2> * 2 ; M[0,0] = tr(S^-2)
    S2 = Si @ Si
    col0 = T[0, :prob.m].cpu().numpy()
    ref0 = np.concatenate([[np.trace(S2)], 2.0 * S2[e_r, e_c]])
    assert (np.abs(col0 - ref0) <= 1e-10 * np.maximum(np.abs(ref0), 1e-3 * np.abs(ref0).max())).all()
    # side vector: tr(A_k S^-1)
    refv = np.concatenate([[np.trace(Si)], 2.0 * Si[e_r, e_c]])
    assert (np.abs(v["dASinvVec"] - refv) <= 1e-10 * np.maximum(np.abs(refv), 1e-3 * np.abs(refv).max())).all()
    # whole path: assembly -> Cholesky (m = 50 000) -> solve, checked through the operator form
    assert kkt.factorize() == 0
    b = prob.rhs + 0.3 * rs.standard_normal(prob.m)
    x = kkt.solve(b)
    back = apply_schur_operator(prob, [Si], x)
    assert np.abs(back - b).max() <= 1e-8 * np.abs(b).max(), np.abs(back - b).max() / np.abs(b).max()
    # linearity of the solve (two right-hand sides in one call)
    X2 = kkt.solve(np.stack([b, 2.0 * b - 1.0], axis=1))
    assert np.abs(X2[:, 0] - x).max() <= 1e-12 * np.abs(x).max()
    one = kkt.solve(np.ones(prob.m))
    assert np.abs(X2[:, 1] - (2.0 * x - one)).max() <= 1e-9 * np.abs(x).max()


def test_config_C_maxcut_n8000_hadamard_square():
    from hdsdp_b200 import api, problem
    n = 8000
    prob = problem.gen_maxcut(n, degree=6, seed=1)
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(-10.0)
    rs = np.random.RandomState(1)
    y = -(8.0 + rs.uniform(0, 1, n))          # S = 10 I - Diag(y) + tau C is diagonally dominant
    cone.update(1.0, y)
    assert cone.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    v = kkt.export()
    Si = cone.get_sinv()
    S = cone.get_buffer(api.BUFFER_DUALVAR); S = np.tril(S) + np.tril(S, -1).T
    z = rs.standard_normal(n)
    assert np.abs(Si @ (S @ z) - z).max() <= 1e-10 * np.abs(z).max()              # S^-1 really is the inverse
    M = np.tril(kkt.get_matrix())
    ref = np.tril(Si * Si)                                                          # M = S^-1 o S^-1 (A_i = e_i e_i^T)
    assert np.abs(M - ref).max() <= 1e-10 * np.abs(ref).max()
    assert np.abs(v["dASinvVec"] - np.diag(Si)).max() <= 1e-10 * np.abs(np.diag(Si)).max()
    assert np.abs(v["dASinvRdSinvVec"] - (-10.0) * (Si * Si).sum(axis=0)).max() <= 1e-9 * np.abs((Si * Si).sum(axis=0)).max() * 10.0
    assert kkt.factorize() == 0
    b = rs.standard_normal(n)
    x = kkt.solve(b)
    Mfull = ref + np.tril(ref, -1).T
    assert np.abs(Mfull @ x - b).max() <= 1e-9 * np.abs(b).max()


def test_config_E_multiblock_m20000_operator_identity():
    from hdsdp_b200 import api, problem
    m = 20000
    prob = problem.gen_multiblock(m)
    sdp, lps, kkt = api.build_problem(prob)
    rs = np.random.RandomState(2)
    y = np.zeros(m)
    tau, rd = 1.0, -1e4          # y = 0, tau = 1 as in the reference harness (tests/test_file_io.c:421-446); R_d large enough for S > 0
    sinvs = []
    for c in sdp:
        c.set_start(rd)
        c.update(tau, y)
        assert c.factorize()
    kkt.build_up(api.KKT_TYPE_INFEASIBLE)
    lp = lps[0]
    lp.dual_residual = rd
    sinv_lp = 1.0 / lp.slack(tau, y)
    assert (sinv_lp > 0).all()
    kkt.build_up_extra_lp(lp, sinv_lp, rd)
    # bound cone l <= y <= u (reference hdsdp_conic_bound.c:201-249): diag(M) += 1/(y-l)^2 + 1/(u-y)^2.  Without it M
    # is rank deficient here (rank <= 5050 + 3000 + 5000 < m), exactly as in the reference, which always adds this cone.
    bound = 1e3
    diag_add = 1.0 / (y + bound) ** 2 + 1.0 / (bound - y) ** 2
    kkt.build_up_extra_bound(diag_add, 1.0 / (bound - y) - 1.0 / (y + bound))
    for c in sdp:
        sinvs.append(c.get_sinv())
    assert kkt.factorize() == 0
    b = rs.standard_normal(m)
    x = kkt.solve(b)
    back = apply_schur_operator(prob, sinvs, x, lp_d2=sinv_lp ** 2) + diag_add * x
    assert np.abs(back - b).max() <= 1e-8 * np.abs(b).max(), np.abs(back - b).max() / np.abs(b).max()
