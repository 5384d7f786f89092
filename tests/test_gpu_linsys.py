"""GPU parity of the dense linear-system back-end (B1) against numpy/LAPACK on the same inputs.

Floating point: tolerance 1e-10 relative (BASELINE.json north_star) on factor diagonal, inverse and solves
for well-conditioned SPD inputs; PSD detection must agree exactly.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def spd(n, seed, cond=1e3):
    rs = np.random.RandomState(seed)
    Q, _ = np.linalg.qr(rs.standard_normal((n, n)))
    ev = np.logspace(0, np.log10(cond), n)
    return (Q * ev) @ Q.T


@pytest.mark.parametrize("n", [1, 7, 100, 128, 129, 300, 1000])
def test_factor_diag_inverse_solve(n):
    from hdsdp_b200.api import DenseLinsys
    A = spd(n, n)
    A = 0.5 * (A + A.T)
    ls = DenseLinsys(n)
    Ain = np.asfortranarray(np.tril(A) + np.triu(np.full((n, n), 7.5), 1))  # garbage strict upper: must be ignored
    keep = Ain.copy()
    assert ls.numeric(Ain) == 0
    assert np.array_equal(Ain, keep), "HFpLinsysNumeric must not modify the caller's matrix"
    L = np.linalg.cholesky(A)
    np.testing.assert_allclose(ls.get_diag(), np.diag(L), rtol=1e-10)
    inv = ls.invert()
    ref = np.linalg.inv(A)
    assert np.abs(inv - ref).max() <= 1e-10 * np.abs(ref).max()
    assert np.abs(inv - inv.T).max() == 0.0, "inverse must be exactly symmetric (mirrored lower)"
    rs = np.random.RandomState(1)
    b = rs.standard_normal(n)
    x = ls.solve(b)
    assert np.abs(x - np.linalg.solve(A, b)).max() <= 1e-9 * np.abs(x).max()
    f = ls.fsolve(b)
    assert np.abs(f - np.linalg.solve(L, b)).max() <= 1e-10 * max(1.0, np.abs(f).max())
    g = ls.bsolve(b)
    assert np.abs(g - np.linalg.solve(L.T, b)).max() <= 1e-10 * max(1.0, np.abs(g).max())
    B = rs.standard_normal((n, 3))
    X = ls.solve(B)
    assert np.abs(X - np.linalg.solve(A, B)).max() <= 1e-9 * np.abs(X).max()
    ls.close()


@pytest.mark.parametrize("n,bad", [(5, 2), (200, 150), (300, 299), (260, 0)])
def test_not_positive_definite(n, bad):
    from hdsdp_b200.api import DenseLinsys
    A = spd(n, 3)
    A[bad, bad] = -1.0
    ls = DenseLinsys(n)
    assert ls.psd_check(A) is False
    assert ls.numeric(A) == 1  # HDSDP_RETCODE_FAILED where dpotrf reports info > 0
    assert ls.psd_check(spd(n, 4)) is True
    ls.close()


def test_gemm_nt_matches_numpy():
    import torch
    from hdsdp_b200 import _lib
    lib = _lib.require_gpu()
    M, N, K = 256, 384, 160
    rs = np.random.RandomState(0)
    A = rs.standard_normal((M, K)); B = rs.standard_normal((N, K)); C = rs.standard_normal((M, N))
    dA = torch.tensor(A.T.copy(), device="cuda")  # column-major M x K == row-major K x M
    dB = torch.tensor(B.T.copy(), device="cuda")
    dC = torch.tensor(C.T.copy(), device="cuda")
    torch.cuda.synchronize()
    rc = lib.hdsdpcu_dgemm_nt_dev(M, N, K, -1.5, dA.data_ptr(), M, dB.data_ptr(), N, 0.5, dC.data_ptr(), M, 0)
    assert rc == 0
    lib.hdsdpcu_sync()
    got = dC.cpu().numpy().T
    ref = -1.5 * A @ B.T + 0.5 * C
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max() * K


@pytest.mark.parametrize("M,N,K,lower", [(1024, 1024, 256, 0), (1024, 1024, 256, 1), (512, 512, 128, 1), (384, 128, 128, 0), (2048, 256, 512, 0)])
def test_gemm_nt_tile_paths(M, N, K, lower):
    """Both tile paths of the DMMA GEMM (thin 32 x 128 tiles below 96 big tiles, 128 x 64 tiles above), full and lower-only."""
    import torch
    from hdsdp_b200 import _lib
    lib = _lib.require_gpu()
    rs = np.random.RandomState(M + K + lower)
    A = rs.standard_normal((M, K)); B = rs.standard_normal((N, K)); C = rs.standard_normal((M, N))
    dA = torch.tensor(A.T.copy(), device="cuda"); dB = torch.tensor(B.T.copy(), device="cuda"); dC = torch.tensor(C.T.copy(), device="cuda")
    torch.cuda.synchronize()
    assert lib.hdsdpcu_dgemm_nt_dev(M, N, K, -1.0, dA.data_ptr(), M, dB.data_ptr(), N, 1.0, dC.data_ptr(), M, lower) == 0
    lib.hdsdpcu_sync()
    got = dC.cpu().numpy().T
    ref = C - A @ B.T
    if lower:
        iu = np.triu_indices(M, 1)
        assert np.array_equal(got[iu], C[iu]), "entries strictly above the diagonal must stay untouched"
        got, ref = np.tril(got), np.tril(ref)
    assert np.abs(got - ref).max() <= 1e-13 * K * np.abs(ref).max()


def test_large_factor_residual():
    """n = 4096: relative residual |A - L L^T| through solves, and log det against numpy."""
    from hdsdp_b200.api import DenseLinsys
    n = 4096
    rs = np.random.RandomState(5)
    G = rs.standard_normal((n, n // 2))
    A = G @ G.T + n * np.eye(n)
    ls = DenseLinsys(n)
    assert ls.numeric(A) == 0
    b = rs.standard_normal(n)
    x = ls.solve(b)
    assert np.abs(A @ x - b).max() <= 1e-10 * np.abs(b).max() * 10
    sign, ld = np.linalg.slogdet(A)
    assert abs(2 * np.log(ls.get_diag()).sum() - ld) <= 1e-10 * abs(ld)
    ls.close()


def sym_indefinite(n, seed, nneg):
    """Symmetric, strongly diagonally dominant by blocks so that unpivoted LDL^T is stable; nneg negative eigen-directions."""
    rs = np.random.RandomState(seed)
    d = np.concatenate([-rs.uniform(1.0, 3.0, nneg), rs.uniform(1.0, 3.0, n - nneg)])
    rs.shuffle(d)
    E = rs.standard_normal((n, n)) * (0.2 / np.sqrt(n))
    return np.diag(d) + 0.5 * (E + E.T), d


@pytest.mark.parametrize("n,nneg", [(60, 7), (128, 1), (300, 40), (1000, 333), (2500, 5)])
def test_indefinite_backend_ldl(n, nneg):
    """Reference a15 (dsytrf/dsytrs fallback, hdsdp_linsolver.c:1662-1825): symmetric indefinite solve + inertia."""
    import ctypes
    from hdsdp_b200 import _lib
    from hdsdp_b200.api import DenseLinsys
    A, d = sym_indefinite(n, n + nneg, nneg)
    ls = DenseLinsys(n)
    lib = _lib.lib()
    assert lib.hdsdpcu_linsys_set_indefinite(ls.h, 1) == 0
    assert ls.numeric(np.asfortranarray(A)) == 0
    neg, pert = ctypes.c_int(-1), ctypes.c_int(-1)
    assert lib.hdsdpcu_linsys_inertia(ls.h, ctypes.byref(neg), ctypes.byref(pert)) == 0
    assert neg.value == int((np.linalg.eigvalsh(A) < 0).sum()) and pert.value == 0      # Sylvester: inertia of J
    rs = np.random.RandomState(0)
    B = rs.standard_normal((n, 3))
    X = ls.solve(B)
    assert np.abs(A @ X - B).max() <= 1e-10 * np.abs(B).max() * n
    assert np.abs(X - np.linalg.solve(A, B)).max() <= 1e-9 * np.abs(X).max()
    ls.close()


def test_indefinite_backend_static_pivot():
    """An exactly singular leading pivot is replaced by the floor instead of producing NaN."""
    import ctypes
    from hdsdp_b200 import _lib
    from hdsdp_b200.api import DenseLinsys
    n = 200
    A, _ = sym_indefinite(n, 9, 3)
    A[0, :] = 0.0; A[:, 0] = 0.0
    ls = DenseLinsys(n)
    lib = _lib.lib()
    lib.hdsdpcu_linsys_set_indefinite(ls.h, 1)
    assert ls.numeric(np.asfortranarray(A)) == 0
    neg, pert = ctypes.c_int(-1), ctypes.c_int(-1)
    lib.hdsdpcu_linsys_inertia(ls.h, ctypes.byref(neg), ctypes.byref(pert))
    assert pert.value == 1
    b = np.random.RandomState(1).standard_normal(n); b[0] = 0.0
    x = ls.solve(b)
    assert np.isfinite(x).all() and np.abs((A @ x - b)[1:]).max() <= 1e-9 * np.abs(b).max()
    ls.close()


def sym_needs_pivoting(n, kind, seed):
    """Symmetric indefinite matrices on which an UNPIVOTED L D L^T meets zero / tiny pivots at once: a scaled GOE matrix with a
    zero diagonal, and a saddle-point matrix [[H, B^T], [B, 0]] with the two kinds of rows interleaved (every second diagonal
    entry is 0; every 128 x 128 diagonal leaf is itself a nonsingular saddle-point matrix)."""
    rs = np.random.RandomState(seed)
    if kind == "zero_diag":
        B = rs.standard_normal((n, n))
        A = (B + B.T) / np.sqrt(2.0 * n)
        np.fill_diagonal(A, 0.0)
        return A
    h = n // 2
    G = rs.standard_normal((h, h))
    H = G @ G.T / h + np.eye(h)
    Bm = rs.standard_normal((n - h, h)) / np.sqrt(h)
    A = np.block([[H, Bm.T], [Bm, np.zeros((n - h, n - h))]])
    p = np.empty(n, dtype=int)
    p[0::2] = np.arange(h)
    p[1::2] = h + np.arange(n - h)
    return A[np.ix_(p, p)]


@pytest.mark.parametrize("n,kind", [(100, "zero_diag"), (128, "saddle"), (300, "zero_diag"), (1024, "saddle"), (1024, "zero_diag"),
                                    (2048, "saddle"), (2500, "zero_diag")])
def test_indefinite_backend_bounded_bunch_kaufman(n, kind):
    """Reference a15: the fallback is dsytrf, i.e. SYMMETRIC PIVOTING with 1 x 1 and 2 x 2 pivots (hdsdp_linsolver.c:1662-1825).
    The device searches its pivots inside every 128 x 128 leaf (bounded Bunch-Kaufman, chol.cu ldl_bk_leaf_kernel): matrices
    whose diagonal is (half) zero factor without a single perturbed pivot and with the inertia of the matrix; without the
    pivoting (option ldl_pivot = 0, the round-1 behaviour) the same matrices need perturbed pivots and lose the solution, or
    (saddle point) lose 2 - 4 digits.
    The gate on the backward error is 1e-8, not LAPACK's 1e-15: pivots never cross a leaf, so an ill-conditioned leaf costs
    digits (measured 1e-16 .. 2e-10 on these cases); on the product path the KKT solve refines on b - M x and fails above 1e-9."""
    import ctypes
    from hdsdp_b200 import _lib
    from hdsdp_b200.api import DenseLinsys
    A = np.asfortranarray(sym_needs_pivoting(n, kind, 3 * n + 1))
    lib = _lib.require_gpu()
    rs = np.random.RandomState(4)
    B = rs.standard_normal((n, 2))
    normA = np.abs(A).sum(axis=1).max()

    def run():
        ls = DenseLinsys(n)
        assert lib.hdsdpcu_linsys_set_indefinite(ls.h, 1) == 0
        assert ls.numeric(A) == 0
        neg, pert = ctypes.c_int(-1), ctypes.c_int(-1)
        assert lib.hdsdpcu_linsys_inertia(ls.h, ctypes.byref(neg), ctypes.byref(pert)) == 0
        X = ls.solve(B)
        ls.close()
        berr = np.abs(A @ X - B).max() / (normA * np.abs(X).max() + np.abs(B).max()) if np.isfinite(X).all() else np.inf
        return neg.value, pert.value, X, berr

    neg, pert, X, berr = run()
    assert pert == 0
    assert neg == int((np.linalg.eigvalsh(A) < 0).sum())                                  # Sylvester: inertia of J
    assert berr <= 1e-8, berr
    Xl = np.linalg.solve(A, B)                                                            # LAPACK dgesv
    assert np.abs(X - Xl).max() <= 1e-8 * np.linalg.cond(A) * np.abs(Xl).max()
    try:                                                                                  # contrast: static pivoting alone
        assert lib.hdsdpcu_set_option(b"ldl_pivot", 0) == 0
        _, pert0, _, berr0 = run()
        assert pert0 > 0 or berr0 > 10.0 * berr       # zero diagonal: perturbed pivots; saddle: 3e2 .. 4e4 times the backward error
    finally:
        lib.hdsdpcu_set_option(b"ldl_pivot", 1)


def test_bunch_kaufman_leaf_agrees_with_static_ldl_where_no_pivoting_is_needed():
    """On a block-diagonally dominant indefinite matrix both variants are stable: same inertia, solutions equal to 1e-12."""
    from hdsdp_b200 import _lib
    from hdsdp_b200.api import DenseLinsys
    lib = _lib.require_gpu()
    n = 700
    A, _ = sym_indefinite(n, 21, 50)
    A = np.asfortranarray(A)
    b = np.random.RandomState(2).standard_normal(n)
    xs = []
    try:
        for piv in (1, 0):
            assert lib.hdsdpcu_set_option(b"ldl_pivot", piv) == 0
            ls = DenseLinsys(n)
            lib.hdsdpcu_linsys_set_indefinite(ls.h, 1)
            assert ls.numeric(A) == 0
            xs.append(ls.solve(b))
            ls.close()
    finally:
        lib.hdsdpcu_set_option(b"ldl_pivot", 1)
    assert np.abs(xs[0] - xs[1]).max() <= 1e-12 * np.abs(xs[0]).max()
    assert np.abs(A @ xs[0] - b).max() <= 1e-12 * np.abs(b).max() * n


@pytest.mark.parametrize("leaf,trsv", [(1, 1), (1, 2), (2, 1), (2, 3)])
def test_alternative_kernel_versions(leaf, trsv):
    """The measurement knobs (include/hdsdpcu.h hdsdpcu_set_option) select older kernel generations: they must stay correct."""
    from hdsdp_b200 import _lib
    from hdsdp_b200.api import DenseLinsys
    lib = _lib.require_gpu()
    try:
        assert lib.hdsdpcu_set_option(b"chol_leaf", leaf) == 0 and lib.hdsdpcu_set_option(b"trsv_version", trsv) == 0
        n = 700
        A = spd(n, 77)
        A = 0.5 * (A + A.T)
        ls = DenseLinsys(n)
        assert ls.numeric(np.asfortranarray(A)) == 0
        np.testing.assert_allclose(ls.get_diag(), np.diag(np.linalg.cholesky(A)), rtol=1e-10)
        B = np.random.RandomState(2).standard_normal((n, 4))
        X = ls.solve(B)
        assert np.abs(X - np.linalg.solve(A, B)).max() <= 1e-9 * np.abs(X).max()
        ls.close()
    finally:
        lib.hdsdpcu_set_option(b"chol_leaf", 2); lib.hdsdpcu_set_option(b"trsv_version", 2)


@pytest.mark.parametrize("n", [130, 1000, 3001])
def test_tile_ticket_triangular_solves(n):
    """trsv_version 3 (one work item per 128 x 128 tile, ordered accumulation): forward, backward and full solves for 1 .. 5
    right-hand sides against the default kernels (1e-12) and against numpy; twice, to see the run-to-run determinism of the
    ordered sums."""
    from hdsdp_b200 import _lib
    from hdsdp_b200.api import DenseLinsys
    lib = _lib.require_gpu()
    A = spd(n, 3 * n)
    A = np.asfortranarray(0.5 * (A + A.T))
    Lref = np.linalg.cholesky(A)
    ls = DenseLinsys(n)
    assert ls.numeric(A) == 0
    rs = np.random.RandomState(n)
    try:
        for nrhs in (1, 2, 3, 5):
            B = rs.standard_normal((n, nrhs)) if nrhs > 1 else rs.standard_normal(n)
            out = {}
            for ver in (2, 3, 3):
                assert lib.hdsdpcu_set_option(b"trsv_version", ver) == 0
                out.setdefault(ver, []).append((ls.fsolve(B), ls.bsolve(B), ls.solve(B)))
            (f2, b2, s2), (f3, b3, s3), (f3b, b3b, s3b) = out[2][0], out[3][0], out[3][1]
            for got, ref in ((f3, f2), (b3, b2), (s3, s2)):
                assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
            assert np.array_equal(f3, f3b) and np.array_equal(b3, b3b) and np.array_equal(s3, s3b)
            Bm = B.reshape(n, -1)
            assert np.abs(Lref @ f3.reshape(n, -1) - Bm).max() <= 1e-11 * np.abs(Bm).max() * n
            assert np.abs(Lref.T @ b3.reshape(n, -1) - Bm).max() <= 1e-11 * np.abs(Bm).max() * n
            assert np.abs(A @ s3.reshape(n, -1) - Bm).max() <= 1e-10 * np.abs(Bm).max() * n
    finally:
        lib.hdsdpcu_set_option(b"trsv_version", 2)
        ls.close()
