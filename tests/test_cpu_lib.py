"""CPU tests of the boundary and the host logic (no compute calls: there is no GPU here)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_loads_and_exports_every_declared_symbol():
    from hdsdp_b200 import _lib
    lib = _lib.lib()   # binds every name in _SIGNATURES; AttributeError if one is missing
    header = open(os.path.join(ROOT, "include", "hdsdpcu.h")).read()
    declared = set(re.findall(r"\b(hdsdpcu_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed from include/hdsdpcu.h"
    assert declared == set(_lib.declared_symbols()), f"header/binding mismatch: {declared ^ set(_lib.declared_symbols())}"
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    for name in declared:
        assert re.search(rf"\bT {name}\b", exported), f"{name} not exported by libhdsdp_cuda.so"
    assert lib.hdsdpcu_version().startswith(b"hdsdp-b200")


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when no CUDA device is usable."""
    from hdsdp_b200 import _lib
    lib = _lib.lib()
    if lib.hdsdpcu_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.HdsdpCudaError):
        _lib.require_gpu()
    assert lib.hdsdpcu_init(0) == 1   # HDSDP_RETCODE_FAILED
    import ctypes
    h = ctypes.c_void_p()
    assert lib.hdsdpcu_linsys_create(ctypes.byref(h), 8) == 1
    assert lib.hdsdpcu_kkt_create(ctypes.byref(h), 8) == 1


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "hdsdp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{f} mentions the oracle"
                assert "liboracle" not in src and "libhdsdp_ref" not in src


def test_sdpa_roundtrip_and_pack_index(tmp_path):
    from hdsdp_b200 import problem
    prob, _ = load_golden("theta1")
    path = str(tmp_path / "t.dat-s")
    problem.write_sdpa(prob, path)
    back = problem.read_sdpa(path)
    assert back.m == prob.m and len(back.cones) == len(prob.cones)
    for a, b in zip(prob.cones, back.cones):
        assert a.dim == b.dim and np.array_equal(a.beg, b.beg) and np.array_equal(a.idx, b.idx) and np.array_equal(a.elem, b.elem)
    n = 7
    seen = set()
    for c in range(n):
        for r in range(c, n):
            p = int(problem.pack_idx(n, r, c))
            assert problem.unpack_idx(n, p) == (r, c)
            seen.add(p)
    assert seen == set(range(n * (n + 1) // 2))


def test_generators_shapes():
    from hdsdp_b200 import problem
    mc = problem.gen_maxcut(50, degree=6, seed=1)
    assert mc.m == 50 and mc.cones[0].dim == 50 and mc.meta["edges"] == 150
    assert np.all(np.diff(mc.cones[0].beg[1:]) == 1)            # A_i = e_i e_i^T: one entry per constraint
    th = problem.gen_theta(30, 100, seed=2)
    assert th.m == 101 and th.cones[0].beg[1] == 30 * 31 // 2   # dense objective J
    assert th.cones[0].beg[2] - th.cones[0].beg[1] == 30        # A_1 = I
    assert th.rhs[0] == 1.0 and not th.rhs[1:].any()
    mb = problem.gen_multiblock(40, n1=8, n2=6, ndense=10, nlp=12, seed=3)
    assert [c.kind for c in mb.cones] == ["sdp", "sdp", "lp"] and mb.cones[2].dim == 12
    C = problem.cone_to_dense(mc.cones[0], 0)
    assert np.allclose(C, C.T) and C.trace() < 0            # stored objective is -C_sdpa


def test_theta_bench_point_is_interior():
    """bench.py's iterate must give S > 0 (checked with the oracle on a scaled-down instance)."""
    import bench
    from hdsdp_b200 import problem
    from oracle import oracle
    prob = problem.gen_theta(60, 200, seed=2)
    y = bench.theta_point(prob.m, 60, 0)
    c = oracle.OracleCone(prob.cones[0], prob.m)
    c.set_resi(bench.RD)
    ok, _ = c.set_point(y, bench.TAU)
    assert ok


def test_block_cyclic_ownership_arithmetic():
    """hdsdpcu_dist_owner is the single definition of which rank assembles / factors a Schur column (pure host)."""
    from hdsdp_b200 import _lib
    lib = _lib.lib()
    assert lib.hdsdpcu_dist_blob_bytes() == 192
    for nb, P in ((128, 2), (512, 8), (256, 3)):
        cols = np.arange(0, 5000, 37)
        got = np.array([lib.hdsdpcu_dist_owner(int(c), nb, P) for c in cols])
        assert np.array_equal(got, (cols // nb) % P)
        counts = np.bincount((np.arange(50048) // nb) % P, minlength=P)
        assert counts.max() - counts.min() <= nb   # block-cyclic balance


def test_binary_container_roundtrip(tmp_path):
    """the CSC user_data arrays travel as one memory-mappable file (no SDPA text parsing for 50k constraints)."""
    from hdsdp_b200 import problem
    for prob in (load_golden("multiblock")[0], problem.gen_theta(60, 300, seed=2)):
        path = str(tmp_path / "p.hdsdpb")
        problem.save_bin(prob, path)
        for mm in (True, False):
            back = problem.load_bin(path, mmap=mm)
            assert back.m == prob.m and len(back.cones) == len(prob.cones) and np.array_equal(back.rhs, prob.rhs)
            for a, b in zip(prob.cones, back.cones):
                assert a.kind == b.kind and a.dim == b.dim
                assert np.array_equal(a.beg, b.beg) and np.array_equal(a.idx, b.idx) and np.array_equal(a.elem, b.elem)
    with open(path, "r+b") as f:
        f.write(b"X")
    with pytest.raises(ValueError):
        problem.load_bin(path)


def test_bench_operator_checker_against_the_oracle():
    """bench.py's correctness check (config.check) applies the operator form of M with numpy; here the checker itself is checked
    against the oracle's explicit Schur matrix on small theta / max-cut / multi-block (+LP) problems."""
    import bench
    from hdsdp_b200 import problem
    from oracle import oracle
    rs = np.random.RandomState(0)
    for prob, rd, y in ((problem.gen_theta(30, 80, seed=2), -1.0, None), (problem.gen_maxcut(40, degree=4, seed=1), -10.0, None),
                        (problem.gen_multiblock(50, n1=8, n2=6, ndense=10, nlp=12, seed=3), -1e3, "zero")):
        m = prob.m
        if y is None:
            y = bench.theta_point(m, 30, 0) if "theta" in prob.name else bench.maxcut_point(m, 40, 0)
        else:
            y = np.zeros(m)
        ok_ = oracle.OracleKKT(m)
        ok_.clean(0)
        sinvs = []
        for cone in prob.cones:
            if cone.kind != "sdp":
                continue
            oc = oracle.OracleCone(cone, m)
            oc.set_resi(rd)
            ok, _ = oc.set_point(y, 1.0)
            assert ok
            oc.build_schur(ok_, 0)
            S = oc.get("S"); S = np.tril(S) + np.tril(S, -1).T
            sinvs.append(np.linalg.inv(S))
        lp_d2 = None
        for cone in prob.cones:
            if cone.kind == "lp":
                s = oracle.lp_slack(cone, 1.0, y, rd)
                ok_.add_lp(cone, s, rd, 0)
                lp_d2 = (1.0 / s) ** 2
        M = np.tril(ok_.M) + np.tril(ok_.M, -1).T
        x = rs.standard_normal(m)
        got = bench.apply_schur_operator(prob, sinvs, x, lp_d2=lp_d2)
        assert np.abs(got - M @ x).max() <= 1e-10 * np.abs(M @ x).max(), prob.name


def test_integration_hooks_are_linked_in():
    """The drop-in build must carry the three hooks: the reference's HConeSetData / HFpLinsysCreate renamed, ours in their place."""
    lib = os.path.join(ROOT, "integration", "_build", "libhdsdp_integrated.so")
    if not os.path.exists(lib):
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    syms = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    for name in ("HConeSetData", "HConeSetData_ref", "HFpLinsysCreate", "HFpLinsysCreate_ref", "HKKTBuildUp", "fds_syev_dimacs",
                 "hdsdpcu_shim_cone_handle", "shim_prof_report"):
        assert re.search(rf"\b[TtDB] {name}\b", syms), f"{name} missing from libhdsdp_integrated.so"
    und = subprocess.check_output(["nm", "-D", "--undefined-only", lib], text=True)
    for name in ("hdsdpcu_cone_update", "hdsdpcu_cone_ratiotest", "hdsdpcu_kkt_factorize", "hdsdpcu_cone_getprimal", "hdsdpcu_sym_extreme_eig"):
        assert name in und, f"the drop-in does not call {name}"
