"""SDP problem data in the reference's `user_data` layout, an SDPA reader and the synthetic generators.

Layout (reference interface/def_hdsdp_user_data.h:11-32): for an SDP block of dimension n with m
constraints, a CSC matrix of shape [n(n+1)/2] x [m+1]; column 0 holds the objective, column i+1 the
constraint A_i; the row index is the packed lower-triangular slot PACK_IDX(n, r, c) =
(2n - c - 1) c / 2 + r with r >= c (interface/hdsdp_utils.h:50).  An LP block is a CSC matrix of
shape [nLpCol] x [m+1] with the same column convention.

The SDPA reader follows reference interface/hdsdp_file_io.c:34-381: the objective entries
(constraint index 0) are negated (:248-250), entries with |v| < 1e-12 are dropped (:226-232), and
entries keep file order inside a column (the reference goes through a triplet->CSC compress).

The generators are the deterministic synthetic inputs named in SURVEY.md section 8(d)
(BASELINE.json configs C, D, E).
"""
from __future__ import annotations

import os
import random
import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


def pack_idx(n: int, r, c):
    """PACK_IDX(n, r, c), r >= c (vectorised)."""
    r = np.asarray(r, dtype=np.int64)
    c = np.asarray(c, dtype=np.int64)
    return (2 * n - c - 1) * c // 2 + r


@dataclass
class ConeData:
    kind: str            # "sdp" or "lp"
    dim: int             # SDP: n ; LP: number of LP columns
    beg: np.ndarray      # int32 [m + 2]
    idx: np.ndarray      # int32 [nnz]
    elem: np.ndarray     # float64 [nnz]


@dataclass
class Problem:
    m: int
    cones: List[ConeData]
    rhs: np.ndarray      # b, float64 [m]
    name: str = ""
    meta: dict = field(default_factory=dict)


def _csc_from_triplets(ncols: int, cols, rows, vals):
    cols = np.asarray(cols, dtype=np.int64)
    order = np.argsort(cols, kind="stable")
    cols = cols[order]
    idx = np.asarray(rows, dtype=np.int64)[order].astype(np.int32)
    elem = np.asarray(vals, dtype=np.float64)[order]
    beg = np.zeros(ncols + 1, dtype=np.int32)
    np.add.at(beg, cols + 1, 1)
    beg = np.cumsum(beg).astype(np.int32)
    return beg, idx, elem


def read_sdpa(path: str) -> Problem:
    with open(path, "r") as f:
        lines = f.read().splitlines()
    # strip comments (lines starting with " or *)
    lines = [ln for ln in lines if ln.strip() and ln.lstrip()[0] not in ('"', "*")]
    m = int(re.findall(r"-?\d+", lines[0])[0])
    nblk = int(re.findall(r"-?\d+", lines[1])[0])
    dims = [int(x) for x in re.findall(r"-?\d+", lines[2])][:nblk]
    num = r"[-+]?(?:\d+\.?\d*(?:[eEdD][-+]?\d+)?|\.\d+(?:[eEdD][-+]?\d+)?)"
    rhs = np.array([float(x.replace("d", "e").replace("D", "e")) for x in re.findall(num, lines[3])][:m], dtype=np.float64)
    if dims[-1] < 0:
        lp_blk = nblk - 1
        nlp = -dims[-1]
        sdp_dims = dims[:-1]
    else:
        lp_blk, nlp, sdp_dims = -1, 0, dims
    if any(d <= 0 for d in sdp_dims):
        raise ValueError("only one diagonal (LP) block, at the end, is supported (reference hdsdp_file_io.c:108-116)")
    trip = [([], [], []) for _ in sdp_dims]
    lp_trip = ([], [], [])
    for ln in lines[4:]:
        tok = ln.split()
        if len(tok) < 5:
            continue
        con, blk, i, j = int(tok[0]), int(tok[1]) - 1, int(tok[2]) - 1, int(tok[3]) - 1
        v = float(tok[4].replace("d", "e").replace("D", "e"))
        if abs(v) < 1e-12:
            continue
        if con == 0:
            v = -v
        if blk == lp_blk:
            lp_trip[0].append(con); lp_trip[1].append(i); lp_trip[2].append(v)
        else:
            r, c = (i, j) if i >= j else (j, i)
            n = sdp_dims[blk]
            trip[blk][0].append(con); trip[blk][1].append((2 * n - c - 1) * c // 2 + r); trip[blk][2].append(v)
    cones = []
    for b, n in enumerate(sdp_dims):
        beg, idx, elem = _csc_from_triplets(m + 1, *trip[b])
        cones.append(ConeData("sdp", n, beg, idx, elem))
    if nlp > 0:
        beg, idx, elem = _csc_from_triplets(m + 1, *lp_trip)
        cones.append(ConeData("lp", nlp, beg, idx, elem))
    return Problem(m=m, cones=cones, rhs=rhs, name=path.split("/")[-1])


def write_sdpa(prob: Problem, path: str) -> None:
    """Inverse of read_sdpa (used to hand synthetic inputs to the reference CLI)."""
    with open(path, "w") as f:
        f.write(f"{prob.m}\n{len(prob.cones)}\n")
        f.write(" ".join(str(c.dim if c.kind == "sdp" else -c.dim) for c in prob.cones) + "\n")
        f.write(" ".join(repr(float(x)) for x in prob.rhs) + "\n")
        for b, cone in enumerate(prob.cones):
            for col in range(prob.m + 1):
                for e in range(cone.beg[col], cone.beg[col + 1]):
                    v = float(cone.elem[e])
                    if col == 0:
                        v = -v
                    if cone.kind == "lp":
                        i = int(cone.idx[e]) + 1
                        f.write(f"{col} {b + 1} {i} {i} {v!r}\n")
                    else:
                        r, c = unpack_idx(cone.dim, int(cone.idx[e]))
                        f.write(f"{col} {b + 1} {c + 1} {r + 1} {v!r}\n")


def unpack_idx(n: int, p: int):
    c = int(np.floor((2 * n + 1 - np.sqrt((2 * n + 1) ** 2 - 8 * p)) / 2))
    while c > 0 and (2 * n - c - 1) * c // 2 + c > p:
        c -= 1
    while (2 * n - (c + 1) - 1) * (c + 1) // 2 + (c + 1) <= p and c + 1 < n:
        c += 1
    r = p - (2 * n - c - 1) * c // 2
    return r, c


def cone_to_dense(cone: ConeData, col: int) -> np.ndarray:
    """Full symmetric n x n matrix of CSC column `col` (0 = objective) -- for tests on small cases."""
    n = cone.dim
    A = np.zeros((n, n))
    for e in range(cone.beg[col], cone.beg[col + 1]):
        r, c = unpack_idx(n, int(cone.idx[e]))
        A[r, c] = cone.elem[e]
        A[c, r] = cone.elem[e]
    return A


# ---------------------------------------------------------------------------------------------------
# synthetic generators
# ---------------------------------------------------------------------------------------------------
def _sample_edges(n: int, nedges: int, rng: random.Random):
    edges = set()
    while len(edges) < nedges:
        i = rng.randrange(n)
        j = rng.randrange(n)
        if i == j:
            continue
        edges.add((min(i, j), max(i, j)))
    return sorted(edges)


def gen_maxcut(n: int, degree: int = 6, seed: int = 1) -> Problem:
    """Config C: max-cut SDP, A_i = e_i e_i^T, b = 1, C = 1/4 (Diag(deg) - Adj) in SDPA (max) form."""
    rng = random.Random(seed)
    edges = _sample_edges(n, n * degree // 2, rng)
    deg = np.zeros(n)
    for i, j in edges:
        deg[i] += 1.0
        deg[j] += 1.0
    cols, rows, vals = [], [], []
    # objective column (col 0), SDPA value negated by the reader -> stored = -C_sdpa; entries in packed order
    ent = {}
    for i in range(n):
        if deg[i] != 0:
            ent[(i, i)] = -0.25 * deg[i]
    for i, j in edges:
        ent[(j, i)] = 0.25  # (row j > col i): -(-1/4)
    for (r, c) in sorted(ent, key=lambda rc: (rc[1], rc[0])):
        cols.append(0); rows.append(int(pack_idx(n, r, c))); vals.append(ent[(r, c)])
    for i in range(n):
        cols.append(i + 1); rows.append(int(pack_idx(n, i, i))); vals.append(1.0)
    beg, idx, elem = _csc_from_triplets(n + 1, cols, rows, vals)
    return Problem(m=n, cones=[ConeData("sdp", n, beg, idx, elem)], rhs=np.ones(n), name=f"maxcut_n{n}_d{degree}_s{seed}",
                   meta={"edges": len(edges)})


def gen_maxcut_lp(n: int, degree: int = 4, seed: int = 1, lp_cost: float = 1.0) -> Problem:
    """Max-cut with one non-negative LP slack per constraint: X_ii + x_i = 1, cost lp_cost * x_i (an SDP cone + an LP cone,
    strictly feasible on both sides): the smallest well-posed input that drives the LP-cone hooks in a full solve."""
    base = gen_maxcut(n, degree, seed)
    cols = [0] * n + list(range(1, n + 1))
    rows = list(range(n)) + list(range(n))
    vals = [lp_cost] * n + [1.0] * n
    beg, idx, elem = _csc_from_triplets(n + 1, cols, rows, vals)
    return Problem(m=n, cones=[base.cones[0], ConeData("lp", n, beg, idx, elem)], rhs=base.rhs.copy(),
                   name=f"maxcutlp_n{n}_d{degree}_s{seed}", meta=dict(base.meta))


def gen_theta(n: int, nedges: int, seed: int = 2) -> Problem:
    """Config D: Lovasz theta, C = J (all ones), A_1 = I (b_1 = 1), A_k = E_ij for every edge (b_k = 0); m = nedges + 1."""
    rng = random.Random(seed)
    edges = _sample_edges(n, nedges, rng)
    m = len(edges) + 1
    npack = n * (n + 1) // 2
    # objective J: SDPA stores +1 everywhere (max form), reader negates -> -1
    idx0 = np.arange(npack, dtype=np.int64)
    val0 = -np.ones(npack)
    ii = np.arange(n, dtype=np.int64)
    idx1 = pack_idx(n, ii, ii)
    val1 = np.ones(n)
    e = np.asarray(edges, dtype=np.int64)  # (i < j) -> row j, col i
    idxe = pack_idx(n, e[:, 1], e[:, 0])
    vale = np.ones(len(edges))
    idx = np.concatenate([idx0, idx1, idxe]).astype(np.int32)
    elem = np.concatenate([val0, val1, vale])
    beg = np.zeros(m + 2, dtype=np.int32)
    beg[1] = npack
    beg[2] = npack + n
    beg[3:] = npack + n + np.arange(1, len(edges) + 1)
    rhs = np.zeros(m)
    rhs[0] = 1.0
    return Problem(m=m, cones=[ConeData("sdp", n, beg, idx, elem)], rhs=rhs, name=f"theta_n{n}_m{m}_s{seed}",
                   meta={"edges": len(edges)})


def gen_multiblock(m: int, n1: int = 100, n2: int = 120, ndense: int = 3000, nlp: int = 5000, seed: int = 3) -> Problem:
    """Config E: SDP1 (n1) with m dense-rank-one rows given as packed a a^T; SDP2 (n2) with `ndense` dense
    symmetric rows (rest zero); LP cone with 3 nonzeros per column.  Feasible by construction:
    X0 = I, b_i = <A_i, X0>, C = sum_i y0_i A_i + I with y0 ~ U(-1, 1)."""
    rs = np.random.RandomState(seed)
    y0 = rs.uniform(-1.0, 1.0, size=m)
    b = np.zeros(m)
    cones = []
    # SDP1: packed a a^T for every row
    np1 = n1 * (n1 + 1) // 2
    tril_c, tril_r = np.triu_indices(n1)  # pairs (c <= r) enumerated column by column -> packed order
    Arows = rs.standard_normal((m, n1))
    C1 = np.eye(n1)
    elem_blocks = [None] * (m + 1)
    for i in range(m):
        a = Arows[i]
        elem_blocks[i + 1] = a[tril_r] * a[tril_c]
        b[i] += a @ a
    C1 = C1 + (Arows.T * y0) @ Arows
    elem_blocks[0] = -C1[tril_r, tril_c]  # stored negated like the SDPA reader does
    idx = np.tile(np.arange(np1, dtype=np.int32), m + 1)
    elem = np.concatenate(elem_blocks)
    beg = (np.arange(m + 2, dtype=np.int64) * np1).astype(np.int32)
    cones.append(ConeData("sdp", n1, beg, idx, elem))
    # SDP2: first `ndense` rows dense symmetric
    np2 = n2 * (n2 + 1) // 2
    t2c, t2r = np.triu_indices(n2)
    C2 = np.eye(n2)
    blocks = []
    nd = min(ndense, m)
    for i in range(nd):
        G = rs.standard_normal((n2, n2))
        A = 0.5 * (G + G.T)
        blocks.append(A[t2r, t2c])
        b[i] += np.trace(A)
        C2 += y0[i] * A
    elem2 = np.concatenate([-C2[t2r, t2c]] + blocks)
    idx2 = np.tile(np.arange(np2, dtype=np.int32), nd + 1)
    beg2 = np.zeros(m + 2, dtype=np.int64)
    beg2[1:nd + 2] = (np.arange(1, nd + 2) * np2)
    beg2[nd + 2:] = (nd + 1) * np2
    cones.append(ConeData("sdp", n2, beg2.astype(np.int32), idx2, elem2))
    # LP cone: nlp columns, 3 nonzeros each; CSC [nlp x (m+1)], row index = LP column
    rs2 = np.random.RandomState(seed + 1)
    lp_rows = np.stack([rs2.choice(m, size=3, replace=False) for _ in range(nlp)])  # constraints touched by each LP column
    lp_vals = rs2.standard_normal((nlp, 3))
    c_lp = np.ones(nlp)  # x0 = 1: b += A 1 ; c = A^T y0 + 1
    cols, rows, vals = [], [], []
    for j in range(nlp):
        for t in range(3):
            k = int(lp_rows[j, t])
            cols.append(k + 1); rows.append(j); vals.append(lp_vals[j, t])
            b[k] += lp_vals[j, t]
            c_lp[j] += y0[k] * lp_vals[j, t]
    for j in range(nlp):
        cols.append(0); rows.append(j); vals.append(-c_lp[j])
    begl, idxl, eleml = _csc_from_triplets(m + 1, cols, rows, vals)
    cones.append(ConeData("lp", nlp, begl, idxl, eleml))
    return Problem(m=m, cones=cones, rhs=b, name=f"multiblock_m{m}_s{seed}", meta={"ndense": nd})


def gen_random_sparse(n: int, m: int, nnz_per_row: int = 3, seed: int = 7, dense_obj: bool = False) -> Problem:
    """Small mixed-sparsity SDP for parity tests: random sparse symmetric A_i, C = sum y0_i A_i + I."""
    rs = np.random.RandomState(seed)
    cols, rows, vals = [], [], []
    Cm = np.eye(n)
    b = np.zeros(m)
    y0 = rs.uniform(-0.5, 0.5, size=m)
    for i in range(m):
        ent = {}
        for _ in range(nnz_per_row):
            r, c = sorted(rs.randint(0, n, size=2), reverse=True)
            ent[(int(r), int(c))] = float(rs.standard_normal())
        for (r, c) in sorted(ent, key=lambda rc: (rc[1], rc[0])):
            v = ent[(r, c)]
            cols.append(i + 1); rows.append(int(pack_idx(n, r, c))); vals.append(v)
            Cm[r, c] += y0[i] * v
            if r != c:
                Cm[c, r] += y0[i] * v
            else:
                b[i] += v
    for c in range(n):
        for r in range(c, n):
            if Cm[r, c] != 0.0 or dense_obj:
                cols.append(0); rows.append(int(pack_idx(n, r, c))); vals.append(-Cm[r, c])
    beg, idx, elem = _csc_from_triplets(m + 1, cols, rows, vals)
    return Problem(m=m, cones=[ConeData("sdp", n, beg, idx, elem)], rhs=b, name=f"randsparse_n{n}_m{m}_s{seed}")


# --------------------------------------------------------------------------------------------------
# Binary container.  The reference reads SDPA text (HReadSDPA, interface/hdsdp_file_io.c:34) into the
# user_data CSC arrays (def_hdsdp_user_data.h:11-32); config D is ~1.2 M text lines and config E 1.2 GB of CSC, so the
# graded synthetic inputs travel as the CSC arrays themselves: one little-endian file, memory-mappable, no parsing.
#   header  : magic "HDSDPB1\0", int64 m, int64 ncones
#   per cone: int64 kind (0 sdp, 1 lp), int64 dim, int64 nnz, then beg int32[m+2], idx int32[nnz], elem float64[nnz]
#             (each array padded to 8 bytes)
#   trailer : rhs float64[m]
# --------------------------------------------------------------------------------------------------
_MAGIC = b"HDSDPB1\0"


def save_bin(prob: Problem, path: str) -> None:
    def pad8(f, nbytes):
        f.write(b"\0" * ((-nbytes) % 8))
    with open(path, "wb") as f:
        f.write(_MAGIC)
        np.array([prob.m, len(prob.cones)], dtype="<i8").tofile(f)
        for c in prob.cones:
            np.array([0 if c.kind == "sdp" else 1, c.dim, len(c.elem)], dtype="<i8").tofile(f)
            beg = np.ascontiguousarray(c.beg, dtype="<i4"); idx = np.ascontiguousarray(c.idx, dtype="<i4")
            assert len(beg) == prob.m + 2 and len(idx) == len(c.elem)
            beg.tofile(f); pad8(f, beg.nbytes)
            idx.tofile(f); pad8(f, idx.nbytes)
            np.ascontiguousarray(c.elem, dtype="<f8").tofile(f)
        np.ascontiguousarray(prob.rhs, dtype="<f8").tofile(f)


def load_bin(path: str, mmap: bool = True) -> Problem:
    """Zero-copy when mmap=True: the CSC arrays are views of the file, handed straight to hdsdpcu_cone_create."""
    raw = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    if bytes(raw[:8]) != _MAGIC:
        raise ValueError(f"{path}: not an HDSDPB1 file")
    off = 8
    m, ncones = (int(v) for v in raw[off:off + 16].view("<i8")); off += 16
    cones = []
    for _ in range(ncones):
        kind, dim, nnz = (int(v) for v in raw[off:off + 24].view("<i8")); off += 24
        nb = 4 * (m + 2)
        beg = raw[off:off + nb].view("<i4"); off += nb + ((-nb) % 8)
        nb = 4 * nnz
        idx = raw[off:off + nb].view("<i4"); off += nb + ((-nb) % 8)
        elem = raw[off:off + 8 * nnz].view("<f8"); off += 8 * nnz
        cones.append(ConeData("sdp" if kind == 0 else "lp", dim, beg, idx, elem))
    rhs = raw[off:off + 8 * m].view("<f8")
    return Problem(m=m, cones=cones, rhs=rhs, name=os.path.basename(path))
