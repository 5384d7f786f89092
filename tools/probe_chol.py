"""Times Cholesky / inverse / solves through the linsys device entry points (leaf kernel versions, solve bandwidth)."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib
lib = _lib.require_gpu(0)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    lib.hdsdpcu_sync()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


sizes = [int(a) for a in sys.argv[1:]] or [1500, 2048, 8192]
for n in sizes:
    h = ctypes.c_void_p(); assert lib.hdsdpcu_linsys_create(ctypes.byref(h), n) == 0
    npad = lib.hdsdpcu_linsys_padded_dim(h)
    G = torch.randn(n, 64, dtype=torch.float64, device="cuda"); A = G @ G.T; A.diagonal().add_(float(n)); del G
    info = ctypes.c_int(0)
    rec = {"n": n}
    for leaf in (1, 2):
        lib.hdsdpcu_set_option(b"chol_leaf", leaf)
        rec[f"potrf_ms_leaf{leaf}"] = ev_time(lambda: lib.hdsdpcu_linsys_numeric_dev(h, A.data_ptr(), n, ctypes.byref(info)))
        assert info.value == 0
    for nrhs in (1, 2, 4):
        x = torch.zeros(npad * nrhs, dtype=torch.float64, device="cuda"); b = torch.randn(n, dtype=torch.float64, device="cuda")
        def solve():
            for r in range(nrhs):
                x[r * npad: r * npad + n] = b
            lib.hdsdpcu_linsys_solve_dev(h, nrhs, x.data_ptr(), npad)
        # time only the solve: fill first
        for r in range(nrhs):
            x[r * npad: r * npad + n] = b
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); lib.hdsdpcu_linsys_solve_dev(h, nrhs, x.data_ptr(), npad); e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1)
        rec[f"solve{nrhs}_ms"] = ms
        rec[f"solve{nrhs}_GBs"] = 2 * 4.0 * n * n / (ms * 1e-3) / 1e9   # two passes over the lower triangle
        rec[f"solve{nrhs}_resid"] = float((A @ x[(nrhs - 1) * npad:(nrhs - 1) * npad + n] - b).abs().max() / b.abs().max())
    if n <= 8192:
        inv = torch.empty(npad * npad, dtype=torch.float64, device="cuda")
        rec["invert_ms"] = ev_time(lambda: lib.hdsdpcu_linsys_invert_dev(h, inv.data_ptr()))
        del inv
    print(json.dumps(rec), flush=True)
    lib.hdsdpcu_linsys_destroy(ctypes.byref(h)); del A
    torch.cuda.empty_cache()
