"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/build_ref.sh):

    OPENBLAS_NUM_THREADS=1 python tests/golden/make_golden.py

For every example shipped with the reference (examples/{mcp100,theta1,truss1,gpp100}.dat-s) and a few
small synthetic problems this stores
  * the input in the reference's user_data CSC layout (read by the reference's own HReadSDPA),
  * per operating point (y, tau, R_d): the dual slack S, diag(L), log det S, S^-1 (single-cone problems),
    the Schur matrix M and the side vectors / scalars for KKT types INFEASIBLE, HOMOGENEOUS, CORRECTOR
    (HKKTBuildUp, interface/hdsdp_schur.c:256), the reference's per-row classification and strategies,
    and M^-1 * dASinvVec from HKKTFactorize + HKKTSolve,
  * the result of a full HDSDPOptimize (objectives, iteration count, DIMACS errors).
The fixtures pin oracle/hdsdp_oracle.c and the CUDA path on machines where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hdsdp_b200 import problem  # noqa: E402
from oracle import refdrv  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
EXAMPLES = os.environ.get("HDSDP_REFERENCE", "/root/reference") + "/examples"


def operating_points(m):
    i = np.arange(m, dtype=np.float64)
    return [
        # the reference harness's own point (tests/test_file_io.c:421-446): y = 0, tau = 1, R_d = -1e3
        {"y": np.zeros(m), "tau": 1.0, "rd": -1e3},
        {"y": 0.25 * np.sin(1.0 + i), "tau": 0.8, "rd": -40.0},
        {"y": 0.05 * np.cos(0.3 * i), "tau": 1.3, "rd": 0.0},
    ]


def dump(prob, solve=True):
    out = {"m": prob.m, "ncones": len(prob.cones), "rhs": prob.rhs}
    for k, c in enumerate(prob.cones):
        out[f"cone{k}_kind"] = 0 if c.kind == "sdp" else 1
        out[f"cone{k}_dim"] = c.dim
        out[f"cone{k}_beg"] = c.beg
        out[f"cone{k}_idx"] = c.idx
        out[f"cone{k}_elem"] = c.elem
    ref = refdrv.RefKKT(prob)
    for k, c in enumerate(prob.cones):
        if c.kind != "sdp":
            continue
        cl = ref.classification(k)
        out[f"cone{k}_refconekind"] = cl["cone_kind"]
        out[f"cone{k}_types"] = cl["types"]
        out[f"cone{k}_perm"] = cl["perm"]
        out[f"cone{k}_strategies"] = cl["strategies"]
        out[f"cone{k}_r1sign"] = np.array([ref.r1_sign(k, i) for i in range(prob.m + 1)])
    npts = 0
    for pt in operating_points(prob.m):
        y, tau, rd = pt["y"], pt["tau"], pt["rd"]
        try:
            logdet = ref.set_point(y, tau, rd)
        except RuntimeError:
            continue  # S not positive definite at this point for this problem
        if not np.isfinite(logdet):
            continue
        p = f"pt{npts}_"
        out[p + "y"] = y; out[p + "tau"] = tau; out[p + "rd"] = rd; out[p + "logdet"] = logdet
        for k, c in enumerate(prob.cones):
            if c.kind == "sdp":
                out[p + f"S{k}"] = np.tril(ref.get_S(k))
                out[p + f"Ldiag{k}"] = ref.get_Ldiag(k)
        for tname, t in (("inf", 0), ("hsd", 2), ("cor", 1)):
            if t == 1:
                ref.build(0)  # the corrector build re-uses M of the preceding full build
            ref.build(t)
            if t != 1:
                out[p + f"M_{tname}"] = np.tril(ref.get_M())
            v = ref.get_vectors()
            for key, val in v.items():
                out[p + f"{tname}_{key}"] = val
            if t == 0 and len([c for c in prob.cones if c.kind == "sdp"]) == 1:
                out[p + "Sinv"] = ref.get_sinv(prob.cones[0].dim)
        ref.build(0)
        v = ref.get_vectors()
        if ref.factorize() == 0:
            out[p + "sol_asinv"] = ref.solve(v["dASinvVec"])
        npts += 1
    out["npoints"] = npts
    ref.close()
    if solve:
        res = refdrv.optimize(prob)
        out["solve_pObj"] = res["pObj"]; out["solve_dObj"] = res["dObj"]; out["solve_iterations"] = res["iterations"]
        out["solve_status"] = res["status"]; out["solve_dimacs"] = res["dimacs"]; out["solve_y"] = res["y"]
    return out


def main():
    probs = []
    for name in ("mcp100", "theta1", "truss1", "gpp100"):
        probs.append((name, refdrv.read_sdpa(f"{EXAMPLES}/{name}.dat-s")))
    probs.append(("maxcut40", problem.gen_maxcut(40, degree=4, seed=1)))
    probs.append(("theta30", problem.gen_theta(30, 60, seed=2)))
    probs.append(("randsparse", problem.gen_random_sparse(24, 40, nnz_per_row=3, seed=7)))
    probs.append(("multiblock", problem.gen_multiblock(60, n1=12, n2=10, ndense=20, nlp=30, seed=3)))
    for name, prob in probs:
        data = dump(prob)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{name}: m={prob.m} points={data['npoints']} iters={data.get('solve_iterations')} "
              f"pObj={data.get('solve_pObj'):.10e} dObj={data.get('solve_dObj'):.10e} -> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
