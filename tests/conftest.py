import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Golden fixture -> (Problem, npz dict).  Fixtures come from tests/golden/make_golden.py (reference outputs)."""
    import numpy as np
    from hdsdp_b200.problem import ConeData, Problem
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    m = int(z["m"])
    cones = []
    for k in range(int(z["ncones"])):
        cones.append(ConeData("sdp" if int(z[f"cone{k}_kind"]) == 0 else "lp", int(z[f"cone{k}_dim"]),
                              z[f"cone{k}_beg"].astype(np.int32), z[f"cone{k}_idx"].astype(np.int32), z[f"cone{k}_elem"]))
    return Problem(m=m, cones=cones, rhs=z["rhs"], name=name), z


GOLDEN_NAMES = ["mcp100", "theta1", "truss1", "gpp100", "maxcut40", "theta30", "randsparse", "multiblock"]
