"""DMMA GEMM variants on the Cholesky trailing-update shape and a square NT GEMM."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
lib = _lib.require_gpu(0)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())
variants = [int(a) for a in sys.argv[1:]] or [3, 4, 5]
shapes = ((32768, 32768, 2048, 1), (8192, 8192, 8192, 0), (16384, 16384, 512, 1), (49152, 512, 512, 0))
if os.environ.get('PROBE_SHAPES') == 'syrk':
    shapes = shapes[:1] + shapes[2:3]
for (M, N, K, lower) in shapes:
    A = torch.randn(K, M, dtype=torch.float64, device="cuda"); B = torch.randn(K, N, dtype=torch.float64, device="cuda")
    C = torch.zeros(N, M, dtype=torch.float64, device="cuda")
    ref = None
    for v in variants:
        lib.hdsdpcu_set_option(b"gemm_variant", v)
        C.zero_(); torch.cuda.synchronize()
        best = 1e30
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            assert lib.hdsdpcu_dgemm_nt_dev(M, N, K, -1.0, A.data_ptr(), M, B.data_ptr(), N, 1.0 if it else 0.0, C.data_ptr(), M, lower) == 0
            e1.record(st); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        flops = 2.0 * M * N * K * (0.5 if lower else 1.0)
        chk = float(C[:64, :64].abs().sum())
        if ref is None: ref = chk
        print(json.dumps({"M": M, "N": N, "K": K, "lower": lower, "variant": v, "ms": round(best, 3), "tflops": round(flops / best / 1e9, 2), "checksum_rel": abs(chk - ref) / ref}), flush=True)
    del A, B, C
