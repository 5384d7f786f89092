"""DMMA GEMM variants on lower-triangular trailing updates with small K (the look-ahead Cholesky's shapes)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
lib = _lib.require_gpu(0)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())
variants = [int(a) for a in sys.argv[1:]] or [1, 3, 4, 5]
for (M, K) in ((8192, 128), (8192, 256), (8192, 512), (20096, 256), (20096, 512), (20096, 1024), (32768, 256), (32768, 2048)):
    A = torch.randn(K, M, dtype=torch.float64, device="cuda")
    C = torch.zeros(M, M, dtype=torch.float64, device="cuda")
    row = {"M": M, "K": K}
    for v in variants:
        lib.hdsdpcu_set_option(b"gemm_variant", v)
        torch.cuda.synchronize()
        best = 1e30
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            assert lib.hdsdpcu_dgemm_nt_dev(M, M, K, -1.0, A.data_ptr(), M, A.data_ptr(), M, 1.0, C.data_ptr(), M, 1) == 0
            e1.record(st); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row[f"v{v}_tf"] = round(float(M) * M * K / best / 1e9, 2)
    print(json.dumps(row), flush=True)
    del A, C
