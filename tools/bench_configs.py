"""Stage times (CUDA events, warm, best of 3) of the hot path on the other BASELINE.json configs:
C = max-cut n = m = 8000 (rank-one unit-vector Schur), E = multi-block m = 20000 (dense rank-one + dense rows + LP + bound).
One JSON line per config: ms per stage, algorithmic TFLOP/s or GB/s where SURVEY 8(d) gives a figure."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdsdp_b200 import _lib, api, problem  # noqa: E402

lib = _lib.require_gpu(0)
st = torch.cuda.ExternalStream(lib.hdsdpcu_stream())


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(best, 3)


def run(name, prob, y, tau, rd, extras=None):
    t0 = time.time()
    sdp, lps, kkt = api.build_problem(prob)
    setup = time.time() - t0
    for c in sdp:
        c.set_start(rd)
    rec = {"config": name, "m": prob.m, "cones": [(c.kind, c.dim) for c in prob.cones], "setup_s": round(setup, 2)}

    def upd():
        for c in sdp:
            c.update(tau, y)
    def fac():
        for c in sdp:
            assert c.factorize()
    def build():
        for c in sdp:
            c.update(tau, y); assert c.factorize()    # invalidates S^-1 so that the build includes the inverse
        kkt.build_up(api.KKT_TYPE_INFEASIBLE)
        if extras:
            extras(kkt, lps)
    upd(); fac(); build()
    rec["S_assembly_ms"] = timed(upd)
    rec["S_cholesky_ms"] = timed(fac)
    t_all = timed(build)
    rec["S_inverse+schur_build_ms"] = round(t_all - rec["S_assembly_ms"] - rec["S_cholesky_ms"], 3)
    rec["M_cholesky_ms"] = timed(lambda: kkt.factorize())
    b = np.random.RandomState(0).standard_normal((prob.m, 2))
    kkt.solve(b)
    t0 = time.time(); kkt.solve(b); rec["M_solve_2rhs_host_ms"] = round((time.time() - t0) * 1e3, 3)
    n_max = max(c.dim for c in prob.cones if c.kind == "sdp")
    rec["M_cholesky_tflops"] = round(prob.m ** 3 / 3.0 / (rec["M_cholesky_ms"] * 1e-3) / 1e12, 2)
    rec["iteration_ms"] = round(rec["S_assembly_ms"] + rec["S_cholesky_ms"] + rec["S_inverse+schur_build_ms"] + rec["M_cholesky_ms"]
                                + rec["M_solve_2rhs_host_ms"], 3)
    print(json.dumps(rec), flush=True)


which = sys.argv[1:] or ["C", "E"]
if "C" in which:
    n = 8000
    prob = problem.gen_maxcut(n, degree=6, seed=1)
    y = -(8.0 + np.random.RandomState(1).uniform(0, 1, n))
    run("C maxcut n=m=8000", prob, y, 1.0, -10.0)
if "E" in which:
    m = 20000
    prob = problem.gen_multiblock(m)
    y = np.zeros(m)

    def extras(kkt, lps):
        lp = lps[0]; lp.dual_residual = -1e4
        kkt.build_up_extra_lp(lp, 1.0 / lp.slack(1.0, y), -1e4)
        kkt.build_up_extra_bound(np.full(m, 2e-6), np.zeros(m))
    run("E multiblock m=20000", prob, y, 1.0, -1e4, extras)
