"""One hot-path step of a bench workload (C, D or E) for `ncu --metrics gpu__time_duration.sum`: python tools/ncu_build.py E [steps]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from hdsdp_b200 import _lib, api, problem
ap = argparse.ArgumentParser(); ap.add_argument("key"); ap.add_argument("steps", type=int, nargs="?", default=2)
a = ap.parse_args()
args = argparse.Namespace(n=bench.THETA_N, edges=bench.THETA_EDGES, maxcut_n=8000, multiblock_m=20000)
torch.cuda.set_device(0)
lib = _lib.require_gpu(0)
lib.hdsdpcu_set_option(b"chol_graph", 0)
hp = bench.HotPath(a.key, torch, lib, api, problem, 0, 1, args)
hp.prepare(a.steps)
for s in range(a.steps):
    hp.step_device(s, False)
lib.hdsdpcu_sync()
