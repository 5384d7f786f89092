"""Small driver for `ncu --set full` captures of the dominant kernel (dgemm_nt_kernel): the Cholesky trailing
update shape (lower-triangular SYRK, K = 2048) and a square NT GEMM.  Run plain first, then under ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib  # noqa: E402

lib = _lib.require_gpu(0)
for (M, N, K, lower) in ((32768, 32768, 2048, 1), (8192, 8192, 8192, 0)):
    A = torch.randn(K, M, dtype=torch.float64, device="cuda")
    C = torch.zeros(N, M, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        assert lib.hdsdpcu_dgemm_nt_dev(M, N, K, -1.0, A.data_ptr(), M, A.data_ptr(), N, 1.0, C.data_ptr(), M, lower) == 0
    lib.hdsdpcu_sync()
print("ok")
