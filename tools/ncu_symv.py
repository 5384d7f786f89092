"""Driver for ncu captures of the HBM-bound kernels on the headline workload (theta n = 1500, m = 50 000): one Schur build
(ss_pair_schur_kernel), one factorisation, one 2-rhs solve (trsv_*2_kernel) and two symv (symv_lower_kernel)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from hdsdp_b200 import _lib, api, problem
args = argparse.Namespace(n=bench.THETA_N, edges=bench.THETA_EDGES, maxcut_n=8000, multiblock_m=20000)
torch.cuda.set_device(0)
lib = _lib.require_gpu(0)
hp = bench.HotPath("D", torch, lib, api, problem, 0, 1, args)
hp.prepare(1)
hp.step_device(0, False)
x = np.random.RandomState(0).standard_normal(hp.m)
for _ in range(2):
    y = hp.kkt.symv(x)
lib.hdsdpcu_sync()
print("ok", float(np.abs(y).max()))
