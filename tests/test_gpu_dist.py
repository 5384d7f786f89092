"""Multi-GPU Cholesky of M (hdsdp_b200/csrc/dist.cu): the complete distributed schedule run with P ranks inside one
process on one GPU (CUDA events in place of the peer-memory flags) must give every rank LAPACK's factor."""
import ctypes

import numpy as np
import pytest

from hdsdp_b200 import _lib

pytestmark = pytest.mark.gpu


def spd(n, seed):
    rs = np.random.RandomState(seed)
    B = rs.standard_normal((n, n))
    return np.asfortranarray(B @ B.T / n + np.eye(n))


@pytest.mark.parametrize("n,nb,P", [(700, 128, 1), (1000, 128, 2), (1500, 256, 3), (2100, 256, 4), (4000, 512, 8), (1300, 512, 2)])
def test_distributed_schedule_matches_lapack(n, nb, P):
    lib = _lib.require_gpu()
    A = spd(n, n + P)
    L0 = np.zeros((n, n), order="F")
    L1 = np.zeros((n, n), order="F")
    info = ctypes.c_int(-1)
    rc = lib.hdsdpcu_distchol_selftest(n, nb, P, A.ctypes.data_as(_lib.c_double_p), L0.ctypes.data_as(_lib.c_double_p),
                                       L1.ctypes.data_as(_lib.c_double_p), ctypes.byref(info))
    assert rc == 0 and info.value == 0
    ref = np.linalg.cholesky(A)
    for L in (L0, L1):
        got = np.tril(L)
        assert np.isfinite(got).all(), "a rank read a block column it never received"
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()


def test_distributed_schedule_reports_indefinite_matrix():
    lib = _lib.require_gpu()
    n, nb, P = 900, 128, 3
    A = spd(n, 5)
    A[650, 650] = -1.0
    info = ctypes.c_int(0)
    L0 = np.zeros((n, n), order="F")
    rc = lib.hdsdpcu_distchol_selftest(n, nb, P, A.ctypes.data_as(_lib.c_double_p), L0.ctypes.data_as(_lib.c_double_p), None, ctypes.byref(info))
    assert rc == 0 and info.value == 651   # LAPACK dpotrf: 1-based index of the first non-positive pivot
