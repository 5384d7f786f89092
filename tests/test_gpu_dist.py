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


@pytest.mark.parametrize("delay", [1, 0])
@pytest.mark.parametrize("n,nb,P", [(700, 128, 1), (1000, 128, 2), (1500, 256, 3), (2100, 256, 4), (4000, 512, 8), (1300, 512, 2),
                                    (3300, 128, 8), (2900, 256, 2)])
def test_distributed_schedule_matches_lapack(n, nb, P, delay):
    """delay = 1: panels applied in pairs (K = 2 nb) to the block columns that are not next in line; 0: every panel at once."""
    lib = _lib.require_gpu()
    assert lib.hdsdpcu_set_option(b"dist_delay", delay) == 0
    A = spd(n, n + P)
    L0 = np.zeros((n, n), order="F")
    L1 = np.zeros((n, n), order="F")
    info = ctypes.c_int(-1)
    rc = lib.hdsdpcu_distchol_selftest(n, nb, P, A.ctypes.data_as(_lib.c_double_p), L0.ctypes.data_as(_lib.c_double_p),
                                       L1.ctypes.data_as(_lib.c_double_p), ctypes.byref(info), 0, None)
    assert rc == 0 and info.value == 0
    ref = np.linalg.cholesky(A)
    for L in (L0, L1):
        got = np.tril(L)
        assert np.isfinite(got).all(), "a rank read a block column it never received"
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
    lib.hdsdpcu_set_option(b"dist_delay", 0)


def test_distributed_schedule_reports_indefinite_matrix():
    lib = _lib.require_gpu()
    n, nb, P = 900, 128, 3
    A = spd(n, 5)
    A[650, 650] = -1.0
    info = ctypes.c_int(0)
    L0 = np.zeros((n, n), order="F")
    rc = lib.hdsdpcu_distchol_selftest(n, nb, P, A.ctypes.data_as(_lib.c_double_p), L0.ctypes.data_as(_lib.c_double_p), None, ctypes.byref(info), 0, None)
    assert rc == 0 and info.value == 651   # LAPACK dpotrf: 1-based index of the first non-positive pivot


@pytest.mark.parametrize("n,nb,P,nneg", [(900, 128, 3, 11), (2100, 256, 4, 300), (1300, 512, 2, 1)])
def test_distributed_ldl_mode(n, nb, P, nneg):
    """The LDL^T fallback through the distributed schedule: A = L J L^T on every rank, signs travel with the panels."""
    lib = _lib.require_gpu()
    rs = np.random.RandomState(n)
    d = np.concatenate([-rs.uniform(1.0, 3.0, nneg), rs.uniform(1.0, 3.0, n - nneg)]); rs.shuffle(d)
    E = rs.standard_normal((n, n)) * (0.2 / np.sqrt(n))
    A = np.asfortranarray(np.diag(d) + 0.5 * (E + E.T))
    L0 = np.zeros((n, n), order="F"); L1 = np.zeros((n, n), order="F"); sg = np.zeros(n)
    info = ctypes.c_int(-1)
    rc = lib.hdsdpcu_distchol_selftest(n, nb, P, A.ctypes.data_as(_lib.c_double_p), L0.ctypes.data_as(_lib.c_double_p),
                                       L1.ctypes.data_as(_lib.c_double_p), ctypes.byref(info), 1, sg.ctypes.data_as(_lib.c_double_p))
    assert rc == 0 and info.value == 0
    assert set(np.unique(sg)) <= {-1.0, 1.0} and int((sg < 0).sum()) == int((np.linalg.eigvalsh(A) < 0).sum())
    for L in (L0, L1):
        T = np.tril(L)
        for s0 in range(0, n, 128):     # diagonal leaves are G = P L Q |Lambda|^(1/2) of the bounded Bunch-Kaufman leaf: all entries count
            T[s0:s0 + 128, s0:s0 + 128] = L[s0:s0 + 128, s0:s0 + 128]
        assert np.isfinite(T).all()
        assert np.abs((T * sg) @ T.T - A).max() <= 1e-11 * np.abs(A).max() * n


def test_multi_process_distributed_schur_on_two_gpus():
    """One process per GPU over CUDA-IPC peer memory (tools/dist_check.py): distributed assembly + Cholesky + solves must equal
    the single-GPU KKT object of the same process.  Needs >= 2 GPUs; skipped on the single-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29655", os.path.join(root, "tools", "dist_check.py"), "200", "3000", "128"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and '"dist_check": "ok"' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
