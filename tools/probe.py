"""One-shot performance probe on a B200: FP64 peaks (cuBLAS DGEMM as the roofline denominator), the DMMA
GEMM kernel, recursive Cholesky, inverse and triangular solves.  Prints one JSON object per line.

    python tools/probe.py [--big]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib  # noqa: E402


def ev_time(stream, fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    stream.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--variant", type=int, default=1)
    args = ap.parse_args()
    lib = _lib.require_gpu(0)
    st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())
    cur = torch.cuda.current_stream()
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "sms": torch.cuda.get_device_properties(0).multi_processor_count}))

    # 1. cuBLAS DGEMM peak (roofline denominator for FP64 tensor work)
    for n in (4096, 8192) + ((16384,) if args.big else ()):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        c = torch.empty_like(a)
        best, mean = ev_time(cur, lambda: torch.matmul(a, b, out=c), reps=5, warm=2)
        print(json.dumps({"probe": "cublas_dgemm", "n": n, "tflops_best": 2 * n ** 3 / best / 1e12, "tflops_mean": 2 * n ** 3 / mean / 1e12}))
        del a, b, c
    # sustained: 3 seconds back to back
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda"); c = torch.empty_like(a)
    torch.cuda.synchronize(); t0 = time.time(); k = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    print(json.dumps({"probe": "cublas_dgemm_sustained", "n": n, "tflops": 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12}))
    del a, b, c

    # 2. our DMMA GEMM
    shapes = [(4096, 4096, 4096, 0), (8192, 8192, 8192, 0), (16384, 16384, 2048, 1), (16384, 16384, 1024, 1), (16384, 16384, 512, 1), (16384, 128, 128, 0)]
    if args.big:
        shapes += [(32768, 32768, 1024, 1)]
    for variant in (0, 1, 2, 3):
      lib.hdsdpcu_set_option(b"gemm_variant", variant)
      for (M, N, K, lower) in shapes:
        A = torch.randn(K, M, dtype=torch.float64, device="cuda"); B = torch.randn(K, N, dtype=torch.float64, device="cuda")
        C = torch.zeros(N, M, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        beta = 1.0 if lower else 0.0   # lower shapes are the Cholesky trailing update: C -= A B^T (read-modify-write)
        fn = lambda: lib.hdsdpcu_dgemm_nt_dev(M, N, K, -1.0 if lower else 1.0, A.data_ptr(), M, B.data_ptr(), N, beta, C.data_ptr(), M, lower)
        best, mean = ev_time(st, fn, reps=3, warm=1)
        flops = 2.0 * M * N * K * (0.5 if lower else 1.0)
        rec = {"probe": "dmma_gemm_nt", "variant": variant, "M": M, "N": N, "K": K, "lower": lower, "ms": best * 1e3, "tflops_best": flops / best / 1e12, "tflops_mean": flops / mean / 1e12}
        if M <= 4096:
            ref = (A.T @ B).T  # C^T layout: C is N x M row-major == M x N column-major
            rec["maxerr"] = float((C - ref).abs().max())
        print(json.dumps(rec))
        del A, B, C

    # 3. Cholesky / inverse / solves through the linsys device entry points
    import ctypes
    lib.hdsdpcu_set_option(b"gemm_variant", args.variant)
    for n, blk in ((2048, 0), (8192, 1024), (16384, 1024)) + (((32768, 1024), (32768, 2048)) if args.big else ()):
        lib.hdsdpcu_set_option(b"chol_block", blk)
        h = ctypes.c_void_p()
        assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
        npad = lib.hdsdpcu_linsys_padded_dim(h)
        G = torch.randn(n, 64, dtype=torch.float64, device="cuda")
        A = G @ G.T
        A.diagonal().add_(float(n))
        del G
        torch.cuda.synchronize()
        info = ctypes.c_int(0)
        t0 = time.time()
        fn = lambda: lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info))
        best, mean = ev_time(st, fn, reps=2, warm=1)
        rec = {"probe": "potrf", "n": n, "chol_block": blk, "info": info.value, "ms": best * 1e3, "tflops": n ** 3 / 3.0 / best / 1e12}
        # residual through one solve
        x = torch.zeros(npad, dtype=torch.float64, device="cuda"); b = torch.randn(n, dtype=torch.float64, device="cuda")
        x[:n] = b
        torch.cuda.synchronize()
        fs = lambda: lib.hdsdpcu_linsys_solve_dev(h, 1, x.data_ptr(), npad)
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        a0.record(st); fs(); a1.record(st); a1.synchronize()
        rec["solve_ms"] = a0.elapsed_time(a1)
        rec["solve_resid"] = float((A @ x[:n] - b).abs().max() / b.abs().max())
        if n <= 16384 and blk == 0:
            inv = torch.empty(npad * npad, dtype=torch.float64, device="cuda")
            fi = lambda: lib.hdsdpcu_linsys_invert_dev(h, inv.data_ptr())
            bi, _ = ev_time(st, fi, reps=2, warm=1)
            rec["invert_ms"] = bi * 1e3
            rec["invert_tflops_alg"] = 2.0 * n ** 3 / 3.0 / bi / 1e12
            del inv
        print(json.dumps(rec))
        lib.hdsdpcu_linsys_destroy(ctypes.byref(h))
        del A, x, b
        torch.cuda.empty_cache()
    print(json.dumps({"launches": lib.hdsdpcu_launch_count(0)}))


if __name__ == "__main__":
    main()
