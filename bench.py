#!/usr/bin/env python
"""bench.py -- headline benchmark of the HDSDP Newton-system hot path on B200.

Metric (BASELINE.json): seconds per IPM iteration spent in the hot path (dual slack S + Cholesky + inverse,
Schur complement assembly, dense FP64 Cholesky of M, solves) at m = 50k, and FP64 TFLOP/s vs peak.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, through libhdsdp_cuda.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation (oracle/_ref)

Workload (config.workload): BASELINE.json configs[3] -- synthetic Lovasz theta, n = 1500 graph, 49 999 edge
constraints + the trace row, m = 50 000 (dense 50k x 50k FP64 M = 20 GB; fits one B200).
One "step" = one IPM iteration's pass over the path at a fresh dual iterate y:
    S = -Rd I - A'y + tau C  ->  Cholesky(S) (PSD check)  ->  S^-1  ->  M_ij = <A_i, S^-1 A_j S^-1> and the side vectors
    ->  HKKTRegularize  ->  Cholesky(M)  ->  two solves (M^-1 b and M^-1 A S^-1, hdsdp_algo.c:1750-1751).
`value`  : inputs already resident in HBM (y and the right-hand sides on the device), no host copies in the timed region.
`e2e`    : the same step through the reference-facing C ABI with HOST buffers (y in, side vectors + two solutions out).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

THETA_N = 1500
THETA_EDGES = 49999
RD = -1.0
TAU = 1.0
KKT_REG = 1e-6


def theta_point(m, n, step):
    """A dual iterate with S = -Rd I - A'y + tau C positive definite: C = -J, so y_1 = -(n + 10) dominates."""
    rs = np.random.RandomState(1000 + step)
    y = 0.1 * rs.uniform(-1.0, 1.0, m)
    y[0] = -(n + 10.0 + 0.01 * step)
    return y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cublas_dgemm_peak(torch, n=8192, reps=5):
    """FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so cuBLAS DGEMM is measured live."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e12


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference (oracle/_ref) on host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def reference_sample(sizes=(8000, 16000), nsolve=2):
    """Times the reference's HKKTBuildUp + HKKTRegularize + HKKTFactorize + 2 x HKKTSolve (its PCG path) on theta
    problems of the SAME graph family (n = 1500) at reduced constraint counts and extrapolates to m = 50 000:
    the reference itself cannot run m > 46 340 (int overflow of nRow*nRow, interface/hdsdp_schur.c:16)."""
    from hdsdp_b200 import problem
    from oracle import refdrv
    cores = os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
    rows = []
    for ne in sizes:
        prob = problem.gen_theta(THETA_N, ne, seed=2)
        ref = refdrv.RefKKT(prob)
        y = theta_point(prob.m, THETA_N, 0)
        t0 = time.time()
        ref.set_point(y, TAU, RD)        # S assembly + Cholesky(S) (+ log det)
        t_s = time.time() - t0
        t = ref.time_iteration(0, KKT_REG, nsolve, 1)
        rows.append({"m": prob.m, "s_update_factor": t_s, "build": t["build"], "factorize": t["factorize"], "solve": t["solve"],
                     "total": t_s + t["total"]})
        ref.close()
    # model: S work independent of m; build, PCG solves ~ m^2 (the reference's default KKT solve is Jacobi-PCG on M,
    # linalg/hdsdp_linsolver.c:1446; it factors M (m^3) only after a PCG failure, which did not happen in the sample)
    big = rows[-1]
    scale = (THETA_EDGES + 1.0) / big["m"]
    extrap = big["s_update_factor"] + (big["build"] + big["solve"]) * scale ** 2 + big["factorize"] * scale ** 3
    return {"rows": rows, "extrapolated_m50000": extrap, "cores": cores}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refdrv
    if not refdrv.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhdsdp_ref.so not built (no /root/reference at build time)"}))
        return
    vals = []
    info = None
    for it in range(args.warmup + args.steps):
        info = reference_sample(sizes=(8000,) if it < args.warmup else (8000, 16000))
        if it >= args.warmup:
            vals.append(info["extrapolated_m50000"])
    v = float(np.mean(vals))
    sample = ("reference HKKTBuildUp+Regularize+Factorize+2xHKKTSolve (Jacobi-PCG) and S update+Cholesky on theta n=1500 at "
              "m=8001 and m=16001, extrapolated to m=50000 with build,solve~m^2 (the reference overflows int at m>46340); measured: "
              + "; ".join(f"m={r['m']}: {r['total']:.2f}s" for r in info["rows"]))
    line = {"metric": "sec/IPM iteration (Schur build + Cholesky) at m=50k", "value": v, "unit": "s/iteration", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "theta n=1500 m=50000 (BASELINE.json configs[3])", "sample_sizes": [r["m"] for r in info["rows"]]},
            "cpu_baseline": {"value": v, "unit": "s/iteration", "cores": info["cores"], "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "s/iteration", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hdsdp_b200 import _lib, api, problem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_gpu(local)
    st = torch.cuda.ExternalStream(lib.hdsdpcu_stream(), device=torch.device("cuda", local))

    n, ne = args.n, args.edges
    prob = problem.gen_theta(n, ne, seed=2)
    m = prob.m
    t0 = time.time()
    sdp, lps, kkt = api.build_problem(prob)
    cone = sdp[0]
    cone.set_start(RD)
    setup_s = time.time() - t0
    mp = lib.hdsdpcu_kkt_padded_dim(kkt.h)
    if world > 1:
        # strong scaling: the SAME m = 50k problem; M is assembled and factored 1-D block-cyclic over the ranks, factor
        # panels travel over NVLink peer memory (hdsdp_b200/csrc/dist.cu); NCCL is used for this handshake and the barriers only
        def allgather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        if args.dist_block <= 0:
            args.dist_block = 512 if world <= 4 else 256   # measured: the panel chain dominates from 8 ranks on
        kkt.dist_init(rank, world, args.dist_block, allgather)

    peak = cublas_dgemm_peak(torch) if (rank == 0 and not args.no_peak) else 0.0

    # device-resident inputs for `value`: y per step and the two right-hand sides (b and A S^-1)
    nsteps = args.warmup + args.steps
    y_host = [theta_point(m, n, s) for s in range(nsteps + 1)]
    y_dev = [torch.tensor(y, device="cuda") for y in y_host]
    rhs_dev = torch.zeros(2 * mp, dtype=torch.float64, device="cuda")
    b_dev = torch.zeros(mp, dtype=torch.float64, device="cuda")
    b_dev[:m] = torch.tensor(prob.rhs, device="cuda")
    torch.cuda.synchronize()
    asinv_ptr = lib.hdsdpcu_kkt_asinv_dev(kkt.h)
    from ctypes import byref, c_int

    def ev():
        return torch.cuda.Event(enable_timing=True)

    fact_ms = []

    def step_device(s, timed):
        flag = c_int(0)
        _lib.check(lib.hdsdpcu_cone_update_dev(cone.h, TAU, y_dev[s].data_ptr()), "update")
        _lib.check(lib.hdsdpcu_cone_factorize(cone.h, 0, byref(flag)), "factorize S")
        assert flag.value == 1, "S not positive definite at the bench iterate"
        _lib.check(lib.hdsdpcu_kkt_buildup(kkt.h, 0), "HKKTBuildUp")
        _lib.check(lib.hdsdpcu_kkt_regularize(kkt.h, KKT_REG), "HKKTRegularize")
        # rhs = [b, A S^-1] (device to device, on the library stream)
        _lib.check(lib.hdsdpcu_copy_dev(rhs_dev.data_ptr(), b_dev.data_ptr(), 8 * mp), "copy b")
        _lib.check(lib.hdsdpcu_copy_dev(rhs_dev.data_ptr() + 8 * mp, asinv_ptr, 8 * mp), "copy ASinv")
        e0, e1 = ev(), ev()
        e0.record(st)
        rc = lib.hdsdpcu_kkt_factorize(kkt.h)
        e1.record(st)
        assert rc == 0, "M not positive definite at the bench iterate"
        _lib.check(lib.hdsdpcu_kkt_solve_dev(kkt.h, 2, rhs_dev.data_ptr()), "solve")
        if timed:
            e1.synchronize()
            fact_ms.append(e0.elapsed_time(e1))

    def step_host(s):
        y = y_host[s]
        cone.update(TAU, y)
        assert cone.factorize()
        kkt.build_up(api.KKT_TYPE_INFEASIBLE)
        kkt.regularize(KKT_REG)
        v = kkt.export()
        assert kkt.factorize() == 0
        d1 = kkt.solve(prob.rhs)
        d2 = kkt.solve(v["dASinvVec"])
        return d1, d2, v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.hdsdpcu_sync()

    # ---- value: HBM-resident -------------------------------------------------------------------
    for s in range(args.warmup):
        step_device(s, False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lib.hdsdpcu_launch_count(1)
    e_start, e_stop = ev(), ev()
    e_start.record(st)
    for s in range(args.warmup, nsteps):
        step_device(s, True)
    e_stop.record(st)
    e_stop.synchronize()
    barrier()
    launches = lib.hdsdpcu_launch_count(0)
    dev_s = e_start.elapsed_time(e_stop) * 1e-3
    clocks = sampler.stop()

    # ---- e2e: host buffers through the C ABI ----------------------------------------------------
    step_host(nsteps)  # warm
    barrier()
    e0, e1 = ev(), ev()
    e0.record(st)
    for s in range(args.steps):
        d1, d2, v = step_host(args.warmup + s)
    e1.record(st)
    e1.synchronize()
    e2e_s = e0.elapsed_time(e1) * 1e-3
    h2d = 8 * (m + 2 * m)            # y, two right-hand sides
    d2h = 8 * (3 * m + 2 * m) + 64   # three side vectors + scalars, two solutions

    # ---- per-stage device times of one more HBM-resident step (not part of `value`) --------------
    def stage_times(s_idx):
        flag = c_int(0)
        marks = []

        def mark(name):
            e = ev(); e.record(st); marks.append((name, e))
        mark("start")
        _lib.check(lib.hdsdpcu_cone_update_dev(cone.h, TAU, y_dev[s_idx].data_ptr()), "update"); mark("S_assembly")
        _lib.check(lib.hdsdpcu_cone_factorize(cone.h, 0, byref(flag)), "factorize S"); mark("S_cholesky")
        _lib.check(lib.hdsdpcu_kkt_buildup(kkt.h, 0), "HKKTBuildUp"); mark("S_inverse+schur_build")
        _lib.check(lib.hdsdpcu_kkt_regularize(kkt.h, KKT_REG), "HKKTRegularize"); mark("regularize")
        _lib.check(lib.hdsdpcu_copy_dev(rhs_dev.data_ptr(), b_dev.data_ptr(), 8 * mp), "copy b")
        _lib.check(lib.hdsdpcu_copy_dev(rhs_dev.data_ptr() + 8 * mp, asinv_ptr, 8 * mp), "copy ASinv")
        assert lib.hdsdpcu_kkt_factorize(kkt.h) == 0; mark("M_cholesky")
        _lib.check(lib.hdsdpcu_kkt_solve_dev(kkt.h, 2, rhs_dev.data_ptr()), "solve"); mark("M_solve_2rhs")
        marks[-1][1].synchronize()
        return {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 3) for i in range(1, len(marks))}

    barrier()
    stages = stage_times(nsteps)
    barrier()

    t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        per_step = dev_s / args.steps
        chol_flops = float(m) ** 3 / 3.0
        fact_s = float(np.mean(fact_ms)) * 1e-3
        achieved = chol_flops / fact_s / 1e12 / world   # per GPU, against the single-GPU peak
        line = {
            "metric": "sec/IPM iteration (Schur build + Cholesky) at m=50k", "value": per_step, "unit": "s/iteration", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"theta n={n} m={m} (BASELINE.json configs[3])", "cone_dim": n, "constraints": m,
                       "schur_bytes": 8 * mp * mp, "l2_policy": "inputs larger than L2 (M is 20 GB; every step uses a new y)",
                       "parallelism": (f"M 1-D block-cyclic (nb={args.dist_block}) over {world} GPUs, peer-memory panel exchange; S-side work replicated" if world > 1 else "single GPU"),
                       "step": "S update + Cholesky(S) + S^-1 + Schur M + regularize + Cholesky(M) + 2 solves", "setup_s": setup_s},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": 20.93e9,
                         "traffic_note": "dram__bytes_read+write of ONE representative dgemm_nt launch (32768^2 lower, K=2048: 2.2e12 of the "
                                         "factorisation's 4.17e13 flop), ncu --set full, profiles/README.md; algorithmic 9.1e9 B for that launch",
                         "kernel": "dgemm_nt_kernel (DMMA) inside Cholesky(M): m^3/3 flop per factorisation" + (f", split over {world} GPUs (achieved is per GPU)" if world > 1 else ""),
                         "peak_source": "cuBLAS DGEMM 8192^3 measured live (MEASURED_PEAKS.json has no FP64 entry)",
                         "factorize_ms": fact_s * 1e3, "share_of_step": fact_s / per_step},
            "e2e": {"value": e2e_s / args.steps, "unit": "s/iteration", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "stages_ms": stages,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                from oracle import refdrv
                if refdrv.available():
                    info = reference_sample()
                    line["cpu_baseline"] = {
                        "value": info["extrapolated_m50000"], "unit": "s/iteration", "cores": info["cores"], "kind": "reference",
                        "sample": "unmodified reference (oracle/_ref) HKKTBuildUp+Regularize+Factorize+2xHKKTSolve and S update+Cholesky on theta "
                                  "n=1500 at m=8001 and m=16001, build/solve extrapolated ~m^2 to m=50000: "
                                  + "; ".join(f"m={r['m']}: {r['total']:.2f}s" for r in info["rows"])}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "s/iteration", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
            except Exception as exc:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "s/iteration", "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=THETA_N)
    ap.add_argument("--edges", type=int, default=THETA_EDGES)
    ap.add_argument("--dist-block", type=int, default=0, help="block-column width of the multi-GPU distribution of M (0 = 512 up to 4 GPUs, 256 beyond)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="skip the live cuBLAS DGEMM peak measurement (ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
