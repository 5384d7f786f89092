#!/usr/bin/env python
"""bench.py -- headline benchmark of the HDSDP Newton-system hot path on B200.

Metric (BASELINE.json): seconds per IPM iteration spent in the hot path (dual slack S + Cholesky + inverse,
Schur complement assembly, dense FP64 Cholesky of M, solves) at m = 50k, and FP64 TFLOP/s vs peak.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, through libhdsdp_cuda.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation (oracle/_ref)
    python bench.py --workload C|D|E                          # headline on another BASELINE.json config (default D)

Workloads (BASELINE.json configs):
    D  configs[3]  synthetic Lovasz theta, n = 1500, 49 999 edge constraints + the trace row, m = 50 000 (M = 20 GB)  [headline]
    C  configs[2]  synthetic max-cut n = m = 8000 (rank-one unit-vector constraints, M = S^-1 o S^-1)
    E  configs[4]  multi-block: SDP n=100 with 20 000 dense rank-one rows, SDP n=120 with 3000 dense rows, LP 5000 columns, bound cone
One "step" = one IPM iteration's pass over the path at a fresh dual iterate y:
    per SDP cone  S = -Rd I - A'y + tau C -> Cholesky(S) (PSD check)  ->  S^-1, M_ij = <A_i, S^-1 A_j S^-1>, side vectors
    (+ LP / bound contributions)  ->  HKKTRegularize  ->  Cholesky(M)  ->  two solves (M^-1 b and M^-1 A S^-1, hdsdp_algo.c:1750-1751).
`value`  : inputs already resident in HBM (y and the right-hand sides on the device), no host copies in the timed region.
`e2e`    : the same step through the reference-facing C ABI with HOST buffers (y in, side vectors + two solutions out).
The default run reports D as the headline line and carries C and E as compact records in config.other_workloads; every
workload is followed, outside the timed region, by a correctness check of the whole path (config.check): the operator identity
(M x)_k = <A_k, S^-1 (sum_l x_l A_l) S^-1> evaluated with numpy from the problem data and a numpy inverse of S, applied to the
device's x = M^-1 b, plus (N > 1) the spread of x across ranks.  A failed check makes the run exit non-zero.
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
CHECK_TOL = 1e-8
HBM_PEAK_FALLBACK = 6456.8   # GB/s, MEASURED_PEAKS.json of this pool (used when the driver-written file is absent)


def theta_point(m, n, step):
    """A dual iterate with S = -Rd I - A'y + tau C positive definite: C = -J, so y_1 = -(n + 10) dominates."""
    rs = np.random.RandomState(1000 + step)
    y = 0.1 * rs.uniform(-1.0, 1.0, m)
    y[0] = -(n + 10.0 + 0.01 * step)
    return y


def maxcut_point(m, n, step):
    """S = 10 I - Diag(y) + tau C diagonally dominant (degree-6 graph, |C_ij| = 1/4)."""
    rs = np.random.RandomState(2000 + step)
    return -(8.0 + rs.uniform(0.0, 1.0, m))


def multiblock_point(m, n, step):
    """y = 0, tau = 1 as in the reference harness (tests/test_file_io.c:421-446) plus a small step-dependent perturbation;
    R_d = -1e4 keeps every block interior."""
    rs = np.random.RandomState(3000 + step)
    return 1e-3 * rs.uniform(-1.0, 1.0, m)


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


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return HBM_PEAK_FALLBACK, "fallback: this pool's MEASURED_PEAKS.json value (file absent on the box)"


def ncu_traffic(kernel):
    """dram bytes (read + write) per launch of a kernel, from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t.get(kernel)
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# host-side checker (numpy, problem data only): the operator form of M
# --------------------------------------------------------------------------------------------------
def packed_rc(n):
    cols = np.repeat(np.arange(n), n - np.arange(n))
    rows = np.concatenate([np.arange(c, n) for c in range(n)])
    return rows, cols


def apply_schur_operator(prob, sinvs, x, lp_d2=None):
    """(M x)_k = sum_cones <A_k, S^-1 X S^-1>, X = sum_l x_l A_l (+ A D^2 A^T x for an LP cone): O(n^3 + nnz), never forms M."""
    m = prob.m
    y = np.zeros(m)
    ks = 0
    for cone in prob.cones:
        beg, idx, elem = cone.beg.astype(np.int64), cone.idx.astype(np.int64), cone.elem
        lo, hi = beg[1], beg[m + 1]
        con = np.repeat(np.arange(m), np.diff(beg[1:m + 2]))
        ii, vv = idx[lo:hi], elem[lo:hi]
        if cone.kind == "sdp":
            n = cone.dim
            R, C = packed_rc(n)
            r, c = R[ii], C[ii]
            Si = sinvs[ks]; ks += 1
            if len(ii) == m and (r == c).all() and (np.diff(beg[1:m + 2]) == 1).all():
                # unit rank-one rows A_k = v_k e_k e_k^T (max-cut): M = (v v^T) o S^-1[r, r] o S^-1[r, r], no n^3 products
                G = Si[np.ix_(r, r)]
                y += vv * ((G * G) @ (vv * x))
                continue
            X = np.zeros((n, n))
            np.add.at(X, (r, c), x[con] * vv)
            X = X + np.tril(X, -1).T
            B = Si @ X @ Si
            w = np.where(r == c, 1.0, 2.0)
            y += np.bincount(con, weights=w * vv * B[r, c], minlength=m)
        else:
            t = np.bincount(ii, weights=vv * x[con], minlength=cone.dim)        # A^T x per LP column
            y += np.bincount(con, weights=vv * (lp_d2 * t)[ii], minlength=m)
    return y


# --------------------------------------------------------------------------------------------------
# one workload = problem + iterate generator + the generic hot-path step
# --------------------------------------------------------------------------------------------------
class HotPath:
    def __init__(self, key, torch, lib, api, problem, rank, world, args):
        self.key, self.torch, self.lib, self.api = key, torch, lib, api
        self.rank, self.world, self.args = rank, world, args
        from hdsdp_b200 import _lib
        self._lib = _lib
        t0 = time.time()
        if key == "D":
            self.prob = problem.gen_theta(args.n, args.edges, seed=2)
            self.rd, self.tau, self.point = RD, TAU, theta_point
            self.label = f"theta n={args.n} m={self.prob.m} (BASELINE.json configs[3])"
            self.bound = None
        elif key == "C":
            self.prob = problem.gen_maxcut(args.maxcut_n, degree=6, seed=1)
            self.rd, self.tau, self.point = -10.0, 1.0, maxcut_point
            self.label = f"max-cut n=m={args.maxcut_n} (BASELINE.json configs[2])"
            self.bound = None
        elif key == "E":
            self.prob = problem.gen_multiblock(args.multiblock_m)
            self.rd, self.tau, self.point = -1e4, 1.0, multiblock_point
            self.label = f"multi-block dense+low-rank+LP m={args.multiblock_m} (BASELINE.json configs[4])"
            self.bound = 1e3     # dual box of the bound cone (hdsdp_conic_bound.c:201-249)
        else:
            raise ValueError(key)
        self.gen_s = time.time() - t0
        self.m = self.prob.m
        self.dims = [c.dim for c in self.prob.cones if c.kind == "sdp"]
        self.nmax = max(self.dims)
        t0 = time.time()
        self.sdp, self.lps, self.kkt = api.build_problem(self.prob)
        for c in self.sdp:
            c.set_start(self.rd)
        for lp in self.lps:
            lp.dual_residual = self.rd
        self.setup_s = time.time() - t0
        self.mp = lib.hdsdpcu_kkt_padded_dim(self.kkt.h)
        self.dist_block = 0
        self.st = torch.cuda.ExternalStream(lib.hdsdpcu_stream(), device=torch.device("cuda", torch.cuda.current_device()))

    # -- multi-GPU: M assembled and factored 1-D block-cyclic over the ranks (csrc/dist.cu); S-side work replicated ---------
    def distribute(self, dist, block):
        if self.world <= 1:
            return

        def allgather(b):
            out = [None] * self.world
            dist.all_gather_object(out, b)
            return out
        self.dist_block = block
        self.kkt.dist_init(self.rank, self.world, block, allgather)

    def prepare(self, nsteps):
        torch, m, mp = self.torch, self.m, self.mp
        self.y_host = [self.point(m, self.nmax, s) for s in range(nsteps + 2)]
        self.y_dev = [torch.tensor(y, device="cuda") for y in self.y_host]
        self.rhs_dev = torch.zeros(2 * mp, dtype=torch.float64, device="cuda")
        self.b_dev = torch.zeros(mp, dtype=torch.float64, device="cuda")
        self.b_dev[:m] = torch.tensor(self.prob.rhs, device="cuda")
        torch.cuda.synchronize()
        self.asinv_ptr = self.lib.hdsdpcu_kkt_asinv_dev(self.kkt.h)
        self.fact_ms = []

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def _extras(self, y):
        """LP cone (host slack inversion, O(nnz), as the integration hook does) and bound cone contributions."""
        for lp in self.lps:
            s = lp.slack(self.tau, y)
            assert (s > 0).all(), "LP slack not positive at the bench iterate"
            self.kkt.build_up_extra_lp(lp, 1.0 / s, self.rd)
        if self.bound is not None:
            lo, up = 1.0 / (y + self.bound), 1.0 / (self.bound - y)
            self.kkt.build_up_extra_bound(lo * lo + up * up, up - lo)

    def step_device(self, s, timed, regularize=True):
        s = s % len(self.y_host)  # fewer prepared points than steps (tiny --steps/--warmup): reuse them
        from ctypes import byref, c_int
        lib, check = self.lib, self._lib.check
        flag = c_int(0)
        for c in self.sdp:
            check(lib.hdsdpcu_cone_update_dev(c.h, self.tau, self.y_dev[s].data_ptr()), "update")
            check(lib.hdsdpcu_cone_factorize(c.h, 0, byref(flag)), "factorize S")
            assert flag.value == 1, "S not positive definite at the bench iterate"
        check(lib.hdsdpcu_kkt_buildup(self.kkt.h, 0), "HKKTBuildUp")
        self._extras(self.y_host[s])
        if regularize:
            check(lib.hdsdpcu_kkt_regularize(self.kkt.h, KKT_REG), "HKKTRegularize")
        # rhs = [b, A S^-1] (device to device, on the library stream)
        check(lib.hdsdpcu_copy_dev(self.rhs_dev.data_ptr(), self.b_dev.data_ptr(), 8 * self.mp), "copy b")
        check(lib.hdsdpcu_copy_dev(self.rhs_dev.data_ptr() + 8 * self.mp, self.asinv_ptr, 8 * self.mp), "copy ASinv")
        e0, e1 = self.ev(), self.ev()
        e0.record(self.st)
        rc = lib.hdsdpcu_kkt_factorize(self.kkt.h)
        e1.record(self.st)
        assert rc == 0, "M not positive definite at the bench iterate"
        check(lib.hdsdpcu_kkt_solve_dev(self.kkt.h, 2, self.rhs_dev.data_ptr()), "solve")
        if timed:
            e1.synchronize()
            self.fact_ms.append(e0.elapsed_time(e1))

    def step_host(self, s, regularize=True):
        s = s % len(self.y_host)  # fewer prepared points than steps (tiny --steps/--warmup): reuse them
        y = self.y_host[s]
        for c in self.sdp:
            c.update(self.tau, y)
            assert c.factorize()
        self.kkt.build_up(self.api.KKT_TYPE_INFEASIBLE)
        self._extras(y)
        if regularize:
            self.kkt.regularize(KKT_REG)
        v = self.kkt.export()
        assert self.kkt.factorize() == 0
        d1 = self.kkt.solve(self.prob.rhs)
        d2 = self.kkt.solve(v["dASinvVec"])
        return d1, d2, v

    def io_bytes(self):
        m = self.m
        nlp = sum(lp.ncol for lp in self.lps)
        h2d = 8 * (m * len(self.sdp) + 2 * m + nlp + (2 * m if self.bound is not None else 0))   # y per cone, two rhs, LP / bound vectors
        d2h = 8 * (3 * m + 2 * m) + 64                                                          # three side vectors + scalars, two solutions
        return h2d, d2h

    def stage_times(self, s):
        from ctypes import byref, c_int
        lib, check, st = self.lib, self._lib.check, self.st
        flag = c_int(0)
        marks = []

        def mark(name):
            e = self.ev(); e.record(st); marks.append((name, e))
        mark("start")
        for c in self.sdp:
            check(lib.hdsdpcu_cone_update_dev(c.h, self.tau, self.y_dev[s].data_ptr()), "update")
        mark("S_assembly")
        for c in self.sdp:
            check(lib.hdsdpcu_cone_factorize(c.h, 0, byref(flag)), "factorize S")
        mark("S_cholesky")
        check(lib.hdsdpcu_kkt_buildup(self.kkt.h, 0), "HKKTBuildUp")
        self._extras(self.y_host[s])
        mark("S_inverse+schur_build")
        check(lib.hdsdpcu_kkt_regularize(self.kkt.h, KKT_REG), "HKKTRegularize"); mark("regularize")
        check(lib.hdsdpcu_copy_dev(self.rhs_dev.data_ptr(), self.b_dev.data_ptr(), 8 * self.mp), "copy b")
        check(lib.hdsdpcu_copy_dev(self.rhs_dev.data_ptr() + 8 * self.mp, self.asinv_ptr, 8 * self.mp), "copy ASinv")
        assert lib.hdsdpcu_kkt_factorize(self.kkt.h) == 0; mark("M_cholesky")
        check(lib.hdsdpcu_kkt_solve_dev(self.kkt.h, 2, self.rhs_dev.data_ptr()), "solve"); mark("M_solve_2rhs")
        marks[-1][1].synchronize()
        return {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 3) for i in range(1, len(marks))}

    def stage_rooflines(self, stages, p64, hbm):
        """algorithmic flops / bytes of each stage over its measured time, as a fraction of the measured peak."""
        m = float(self.m)
        n3 = float(sum(d ** 3 for d in self.dims))
        n2 = float(sum(d ** 2 for d in self.dims))
        out = {}

        def put(stage, flops=None, bytes_=None):
            ms = stages.get(stage)
            if not ms or ms <= 0:
                return
            rec = {"ms": ms}
            if flops:
                rec["tflops"] = round(flops / (ms * 1e-3) / 1e12, 3)
                if p64:
                    rec["frac_fp64_peak"] = round(rec["tflops"] / p64, 4)
            if bytes_:
                rec["gbs"] = round(bytes_ / (ms * 1e-3) / 1e9, 1)
                rec["frac_hbm_peak"] = round(rec["gbs"] / hbm, 4)
            out[stage] = rec
        put("S_cholesky", flops=n3 / 3.0)
        if self.key == "D":      # S^-1 (2n^3/3) + M5 sparse pairs: bound by writing the lower triangle of M
            put("S_inverse+schur_build", flops=2.0 * n3 / 3.0, bytes_=4.0 * m * m + 8.0 * n2)
        elif self.key == "C":    # S^-1 (2n^3/3, tensor) + M2 with unit vectors (8n^2 read + 4m^2 written)
            put("S_inverse+schur_build", flops=2.0 * n3 / 3.0, bytes_=8.0 * n2 + 4.0 * m * m)
        else:                    # E: DSR1 block 2 n1^2 m + n1 m^2 ; dense rows: U_i = A_i S^-1 (2 n2^3 each) + Gram (nd^2 n2^2 lower)
            n1, n2d = float(self.dims[0]), float(self.dims[1])
            nd = float(self.prob.meta.get("ndense", 0))
            put("S_inverse+schur_build", flops=2.0 * n1 * n1 * m + n1 * m * m + nd * 2.0 * n2d ** 3 + nd * nd * n2d * n2d)
        put("M_cholesky", flops=m ** 3 / 3.0 / max(self.world, 1))
        put("M_solve_2rhs", bytes_=8.0 * m * m)
        return out

    def check(self, dist, s):
        """Whole path at this workload's full size against numpy: x = M^-1 b from the device (unregularised build at iterate s),
        residual of the operator identity on every rank (rank 0 only for n > 4000: the numpy inverse is n^3 host work), and the
        spread of x across ranks."""
        torch = self.torch
        d1, d2, v = self.step_host(s, regularize=False)
        rs = np.random.RandomState(77)
        b = self.prob.rhs + 0.3 * rs.standard_normal(self.m)
        x = self.kkt.solve(b)
        rec = {"tol": CHECK_TOL}
        do_host = self.rank == 0 or self.nmax <= 4000
        resid = 0.0
        if do_host:
            sinvs = []
            for c in self.sdp:
                S = c.get_buffer(self.api.BUFFER_DUALVAR)
                S = np.tril(S) + np.tril(S, -1).T
                sinvs.append(np.linalg.inv(S))
            lp_d2 = None
            if self.lps:
                lp_d2 = (1.0 / self.lps[0].slack(self.tau, self.y_host[s])) ** 2
            back = apply_schur_operator(self.prob, sinvs, x, lp_d2=lp_d2)
            if self.bound is not None:
                y = self.y_host[s]
                back += (1.0 / (y + self.bound) ** 2 + 1.0 / (self.bound - y) ** 2) * x
            resid = float(np.abs(back - b).max() / np.abs(b).max())
            # the first solve of the step is M^-1 rhs as well: check it the same way through linearity of the operator
        t = torch.tensor([resid], dtype=torch.float64, device="cuda")
        xs = torch.tensor(x, device="cuda")
        spread = 0.0
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            hi, lo = xs.clone(), xs.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            spread = float((hi - lo).abs().max() / xs.abs().max())
        rec["operator_residual"] = float(t[0])
        rec["ranks_checked"] = "all" if self.nmax <= 4000 else "rank 0"
        rec["cross_rank_max_rel_diff"] = spread
        rec["ok"] = bool(rec["operator_residual"] <= CHECK_TOL and spread <= 1e-12)
        return rec

    def close(self):
        self.kkt.close()
        for c in self.sdp:
            c.close()
        for lp in self.lps:
            lp.close()
        self.y_dev = self.rhs_dev = self.b_dev = None
        self.torch.cuda.empty_cache()


def measure(hp, dist, args, steps, warmup, with_clocks):
    """W warm-up steps, K timed HBM-resident steps, K timed host-buffer steps, stage times, check.  Device times are CUDA events on
    the library stream, max over ranks."""
    torch, lib = hp.torch, hp.lib
    nsteps = warmup + steps
    hp.prepare(nsteps)

    def barrier():
        if hp.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.hdsdpcu_sync()

    for s in range(warmup):
        hp.step_device(s, False)
    barrier()
    sampler = None
    if with_clocks:
        sampler = ClockSampler(torch.cuda.current_device())
        sampler.start()
    lib.hdsdpcu_launch_count(1)
    e_start, e_stop = hp.ev(), hp.ev()
    e_start.record(hp.st)
    for s in range(warmup, nsteps):
        hp.step_device(s, True)
    e_stop.record(hp.st)
    e_stop.synchronize()
    barrier()
    launches = lib.hdsdpcu_launch_count(0)
    dev_s = e_start.elapsed_time(e_stop) * 1e-3
    clocks = sampler.stop() if sampler else None

    hp.step_host(nsteps)  # warm
    barrier()
    e0, e1 = hp.ev(), hp.ev()
    e0.record(hp.st)
    for s in range(steps):
        hp.step_host(warmup + s)
    e1.record(hp.st)
    e1.synchronize()
    e2e_s = e0.elapsed_time(e1) * 1e-3
    barrier()
    stages = hp.stage_times(nsteps)
    barrier()
    chk = hp.check(dist, nsteps + 1)
    barrier()
    t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    if hp.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"dev_s": float(t[0]), "e2e_s": float(t[1]), "launches": launches, "clocks": clocks, "stages": stages, "check": chk,
            "fact_s": float(np.mean(hp.fact_ms)) * 1e-3}


def measure_pcg(hp, steps):
    """The same step with the reference's own solver policy for M (Jacobi-PCG first, hdsdp_linsolver.c:1446-1588) on the device:
    no m^3 factorisation while CG converges.  Reported next to the headline (which is the direct Cholesky the metric names)."""
    torch, lib = hp.torch, hp.lib
    hp.kkt.set_solver(1)
    try:
        for s in range(2):
            hp.step_device(s, False)
        lib.hdsdpcu_sync()
        e0, e1 = hp.ev(), hp.ev()
        e0.record(hp.st)
        for s in range(steps):
            hp.step_device(2 + s, False)
        e1.record(hp.st)
        e1.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / steps
        st = hp.kkt.pcg_status()
        stages = hp.stage_times(1)
        chk = hp.check(None, 0)
        st_after = hp.kkt.pcg_status()
    finally:
        hp.kkt.set_solver(0)
    return {"value": t, "unit": "s/iteration", "steps": steps, "solver": "Jacobi-PCG on M (reference default policy), device symv",
            "cg_iterations_last_solve": st["last_iterations"], "cg_solves": st_after["n_solves"], "fallbacks_to_cholesky": st_after["n_fallbacks"],
            "stages_ms": stages, "check": chk}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference (oracle/_ref) on host cores
# --------------------------------------------------------------------------------------------------
def reference_point(ne, nsolve=2):
    """One measurement of the reference's own S update + Cholesky(S), HKKTBuildUp + HKKTRegularize + HKKTFactorize +
    2 x HKKTSolve (its PCG path; dpotrf only after a PCG failure, hdsdp_linsolver.c:1558-1567) on theta n = 1500 with ne edges."""
    from hdsdp_b200 import problem
    from oracle import refdrv
    prob = problem.gen_theta(THETA_N, ne, seed=2)
    t0 = time.time()
    ref = refdrv.RefKKT(prob)
    t_setup = time.time() - t0
    y = theta_point(prob.m, THETA_N, 0)
    t0 = time.time()
    ref.set_point(y, TAU, RD)        # S assembly + Cholesky(S) (+ log det)
    t_s = time.time() - t0
    t = ref.time_iteration(0, KKT_REG, nsolve, 1)
    cg = ref.cg_status()
    ref.close()
    return {"m": prob.m, "s_update_factor": t_s, "build": t["build"], "factorize": t["factorize"], "solve": t["solve"],
            "total": t_s + t["total"], "setup": t_setup, "cg": cg}


def extrapolate(row, m_to):
    """m -> m_to: Schur assembly and PCG mat-vecs ~ m^2; a Cholesky of M (only if PCG fell back to it) ~ m^3."""
    scale = float(m_to) / row["m"]
    fell_back = bool(row["cg"] and row["cg"].get("use_jacobi") == 0)
    if fell_back:   # the fallback factorisation happens inside the first solve after the failed PCG
        return row["s_update_factor"] + row["build"] * scale ** 2 + (row["factorize"] + row["solve"]) * scale ** 3, fell_back
    return row["s_update_factor"] + (row["build"] + row["solve"]) * scale ** 2 + row["factorize"] * scale ** 2, fell_back


def reference_sample(sizes=(8000, 16000)):
    cores = os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
    rows = [reference_point(ne) for ne in sizes]
    ext, fb = extrapolate(rows[-1], THETA_EDGES + 1)
    return {"rows": rows, "extrapolated_m50000": ext, "cores": cores, "fell_back_to_cholesky": fb}


def mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return float(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refdrv
    if not refdrv.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhdsdp_ref.so not built (no /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
    # every step: one bounded sample (m = 8001 and m = 16001, a few seconds)
    step_s, info = [], None
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        info = reference_sample(sizes=(8000,) if it < args.warmup else (8000, 16000))
        if it >= args.warmup:
            step_s.append(time.time() - t0)
    # once: the largest size the reference can run (m = 46 001 < 46 340, the int overflow of nRow*nRow at hdsdp_schur.c:16);
    # M alone is 17 GB, its Cholesky fallback copies it, so 45 GB of free host memory are required
    big = None
    big_note = "skipped (--ref-big-edges 0)"
    if args.ref_big_edges > 0:
        if mem_available_gb() >= 45.0:
            t0 = time.time()
            try:
                big = reference_point(args.ref_big_edges)
                big["wall"] = time.time() - t0
                big_note = "measured once in this run"
            except Exception as exc:
                big_note = f"failed: {exc}"
        else:
            big_note = f"skipped: {mem_available_gb():.0f} GB host memory available, 45 GB needed"
    rows = info["rows"] + ([big] if big else [])
    base = big if big else info["rows"][-1]
    v, fell_back = extrapolate(base, THETA_EDGES + 1)
    algo = ("Jacobi-preconditioned CG on M (no m^3 factorisation)" if not fell_back else
            "PCG failed -> dpotrf of a copy of M as preconditioner (m^3/3), hdsdp_linsolver.c:1558-1567")
    sample = (f"unmodified reference (oracle/_ref): S update + Cholesky(S), HKKTBuildUp + HKKTRegularize + HKKTFactorize + 2 x HKKTSolve "
              f"on theta n=1500; every step times m=8001 and m=16001 ({np.mean(step_s):.1f} s per step); m={base['m']} {big_note}: "
              f"{base['total']:.2f} s; value = that iteration extrapolated {base['m']} -> 50000 (x{(50000.0 / base['m']) ** 2:.2f} on the m^2 parts"
              + (f", x{(50000.0 / base['m']) ** 3:.2f} on the factorisation" if fell_back else "") + "; the reference overflows int at m > 46340); "
              + "; ".join(f"m={r['m']}: build {r['build']:.2f} factorize {r['factorize']:.2f} solves {r['solve']:.2f} total {r['total']:.2f} s "
                          f"(CG iterations of the last solve {r['cg'].get('n_iters') if r['cg'] else '?'}, Jacobi {r['cg'].get('use_jacobi') if r['cg'] else '?'})"
                          for r in rows))
    line = {"metric": "sec/IPM iteration (Schur build + Cholesky) at m=50k", "value": v, "unit": "s/iteration", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(step_s)) * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "theta n=1500 m=50000 (BASELINE.json configs[3])", "sample_sizes": [r["m"] for r in rows],
                       "extrapolated": True, "extrapolated_from_m": base["m"], "algorithm": algo,
                       "ms_per_step_is": "wall time of one bounded sample step (m=8001 + m=16001), NOT the value; the value is the "
                                         "measured iteration at the largest size scaled to m=50000",
                       "rows": rows},
            "cpu_baseline": {"value": v, "unit": "s/iteration", "cores": cores, "kind": "reference", "sample": sample,
                             "extrapolated": True, "algorithm": algo},
            "e2e": {"value": v, "unit": "s/iteration", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
METRIC = {"D": "sec/IPM iteration (Schur build + Cholesky) at m=50k",
          "C": "sec/IPM iteration (Schur build + Cholesky) at max-cut n=m=8000",
          "E": "sec/IPM iteration (Schur build + Cholesky) at multi-block m=20000"}


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
    peak = cublas_dgemm_peak(torch) if (rank == 0 and not args.no_peak) else 0.0
    hbm, hbm_src = hbm_peak()

    def block_for(m):
        if args.dist_block > 0:
            return args.dist_block
        return 512 if world <= 4 else 256   # measured: the panel chain dominates from 8 ranks on

    # ---- headline workload ---------------------------------------------------------------------
    hp = HotPath(args.workload, torch, lib, api, problem, rank, world, args)
    if world > 1 and hp.m >= args.dist_min_m:
        hp.distribute(dist, block_for(hp.m))
    r = measure(hp, dist, args, args.steps, args.warmup, with_clocks=True)
    m, n = hp.m, hp.nmax
    per_step = r["dev_s"] / args.steps
    h2d, d2h = hp.io_bytes()
    chol_flops = float(m) ** 3 / 3.0
    nshare = world if hp.dist_block else 1
    achieved = chol_flops / r["fact_s"] / 1e12 / nshare     # per GPU, against the single-GPU peak
    parallelism = "single GPU"
    if world > 1:
        parallelism = (f"M 1-D block-cyclic (nb={hp.dist_block}) over {world} GPUs, peer-memory panel exchange; S-side work replicated"
                       if hp.dist_block else f"replicated on {world} GPUs (m < {args.dist_min_m}: distributing M does not pay)")
    ok = r["check"]["ok"]
    setup_s, gen_s = hp.setup_s, hp.gen_s
    stage_roof = hp.stage_rooflines(r["stages"], peak, hbm)
    pcg = None
    if world == 1 and not args.no_pcg:
        pcg = measure_pcg(hp, args.other_steps)
        ok = ok and pcg["check"]["ok"]
    hp.close()

    # ---- the other BASELINE.json configs as compact records ------------------------------------------
    others = {}
    if not args.no_other_workloads:
        for key in [k for k in ("C", "D", "E") if k != args.workload]:
            try:
                o = HotPath(key, torch, lib, api, problem, rank, world, args)
                if world > 1 and o.m >= args.dist_min_m:
                    o.distribute(dist, block_for(o.m))
                ro = measure(o, dist, args, args.other_steps, 3, with_clocks=False)
                oh2d, od2h = o.io_bytes()
                others[key] = {
                    "workload": o.label, "value": ro["dev_s"] / args.other_steps, "unit": "s/iteration", "steps": args.other_steps, "warmup": 3,
                    "e2e": {"value": ro["e2e_s"] / args.other_steps, "unit": "s/iteration", "h2d_bytes_per_step": oh2d, "d2h_bytes_per_step": od2h},
                    "scaling": "strong" if o.dist_block else ("single GPU" if world == 1 else "replicated (every rank runs the whole iteration)"),
                    "dist_block": o.dist_block, "stages_ms": ro["stages"], "stage_roofline": o.stage_rooflines(ro["stages"], peak, hbm),
                    "gpu_launches": ro["launches"], "check": ro["check"], "setup_s": round(o.setup_s, 2), "generate_s": round(o.gen_s, 2)}
                ok = ok and ro["check"]["ok"]
                o.close()
            except Exception as exc:   # an extra record must never take the headline down silently: report and fail the run
                others[key] = {"error": repr(exc)}
                ok = False

    if rank == 0:
        line = {
            "metric": METRIC[args.workload], "value": per_step, "unit": "s/iteration", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": hp.label, "cone_dim": n, "constraints": m,
                       "schur_bytes": 8 * hp.mp * hp.mp, "l2_policy": "inputs larger than L2 (M alone exceeds the 126 MB L2; every step uses a new y)",
                       "parallelism": parallelism,
                       "step": "S update + Cholesky(S) + S^-1 + Schur M + regularize + Cholesky(M) + 2 solves", "setup_s": setup_s,
                       "generate_s": gen_s, "check": r["check"], "stages_ms": r["stages"], "stage_roofline": stage_roof,
                       "reference_policy_pcg": pcg, "other_workloads": others},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": ncu_traffic("dgemm_nt_kernel"),
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one dgemm_nt launch of the trailing update "
                                         "(32768^2 lower, K=2048), ncu --set full, profiles/ncu_traffic.json; null if that file is absent",
                         "kernel": "dgemm_nt_kernel (DMMA) inside Cholesky(M): m^3/3 flop per factorisation"
                                   + (f", split over {world} GPUs (achieved is per GPU)" if nshare > 1 else ""),
                         "peak_source": "cuBLAS DGEMM 8192^3 measured live (MEASURED_PEAKS.json has no FP64 entry)",
                         "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src,
                         "factorize_ms": r["fact_s"] * 1e3, "share_of_step": r["fact_s"] / per_step, "stages_ms": r["stages"]},
            "e2e": {"value": r["e2e_s"] / args.steps, "unit": "s/iteration", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": r["launches"], "clocks": r["clocks"],
        }
        if not args.no_cpu_baseline and world == 1 and args.workload == "D":
            try:
                from oracle import refdrv
                if refdrv.available():
                    info = reference_sample()
                    algo = "PCG fell back to dpotrf" if info["fell_back_to_cholesky"] else "Jacobi-PCG (no m^3 factorisation)"
                    line["cpu_baseline"] = {
                        "value": info["extrapolated_m50000"], "unit": "s/iteration", "cores": info["cores"], "kind": "reference",
                        "extrapolated": True, "algorithm": algo,
                        "sample": "unmodified reference (oracle/_ref) S update+Cholesky, HKKTBuildUp+Regularize+Factorize+2xHKKTSolve on theta "
                                  "n=1500 at m=8001 and m=16001 (bounded sample; `bench.py --impl reference` measures m=46001), "
                                  f"EXTRAPOLATED ~m^2 from m=16001 to m=50000, {algo}: "
                                  + "; ".join(f"m={r_['m']}: {r_['total']:.2f}s" for r_ in info["rows"])}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "s/iteration", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
            except Exception as exc:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "s/iteration", "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: correctness check FAILED (config.check / config.other_workloads[*].check)\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="D", choices=["C", "D", "E"])
    ap.add_argument("--n", type=int, default=THETA_N)
    ap.add_argument("--edges", type=int, default=THETA_EDGES)
    ap.add_argument("--maxcut-n", type=int, default=8000)
    ap.add_argument("--multiblock-m", type=int, default=20000)
    ap.add_argument("--other-steps", type=int, default=5, help="timed steps of the non-headline workloads")
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--no-pcg", action="store_true", help="skip the reference-policy (Jacobi-PCG) record of the headline workload")
    ap.add_argument("--dist-block", type=int, default=0, help="block-column width of the multi-GPU distribution of M (0 = 512 up to 4 GPUs, 256 beyond)")
    ap.add_argument("--dist-min-m", type=int, default=16000, help="below this m the Schur matrix is replicated instead of distributed over the GPUs")
    ap.add_argument("--ref-big-edges", type=int, default=46000, help="reference arm: theta edges of the one large measurement (m = edges + 1; 0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="skip the live cuBLAS DGEMM peak measurement (ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
