"""Full-solve end points of the UNMODIFIED CPU reference (oracle/_ref) on the mid-size problems of tests/test_gpu_integration.py,
one run per OpenBLAS thread count.  Only the summation order inside the BLAS differs between the runs, yet on theta n = 200,
m = 3001 the reference lands on one of two end points (48 iterations / dObj -39.4518775 or 33 iterations / dObj -39.4518854):
the PSDP primal-refinement tail crawls with steps of 1e-2 and stops on a threshold.  Which one a given box produces depends on
its core count and CPU, so the test accepts the end points the reference itself reaches -- those run live on the box plus the
ones recorded here.   Run where /root/reference exists:  python tests/golden/make_endpoints.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools import fullsolve  # noqa: E402

SPECS = ["theta:200:3000", "maxcut:1000:4", "maxcutlp:300:4"]
out = {}
for spec in SPECS:
    s = fullsolve.parse_spec(spec)
    runs = []
    for threads in (1, 2, 4, 8):
        r, log, err = fullsolve.run(s, False, threads)
        assert r is not None, (spec, threads, err[-2000:])
        runs.append({"threads": threads, "iterations": r["iterations"], "dObj": r["dObj"], "pObj": r["pObj"], "status": r["status"],
                     "psdp": "Primal refinement starts" in log})
        print(spec, runs[-1], flush=True)
    out[spec] = runs
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_endpoints.json"), "w") as f:
    json.dump(out, f, indent=1)
