// hdsdp_b200/csrc/classify.cpp -- host-side classification of SDP coefficient matrices.
//
// Mirrors the reference's presolve so that the device image uses the same storage class per
// coefficient (ZERO / SPARSE / DENSE / SPR1 / DSR1):
//   sdpDataMatSetData      linalg/hdsdp_sdpdata.c:2321-2345   (zero / dense if nnz > 0.3 n(n+1)/2 / sparse)
//   tsp_decompress         linalg/sparse_opts.c:428-443        (packed CSC row index -> (row, col))
//   tsp_r1_extract         linalg/sparse_opts.c:453-516        (sparse rank-one test, tol 1e-10)
//   pds_r1_extract         linalg/dense_opts.c:233-285         (dense rank-one test)
//   sdpDataMatBuildUpEigs  linalg/hdsdp_sdpdata.c:2373-2449    (SPR1 if #{|a_k|>1e-10} <= 0.5 n else DSR1)
//   normalisation          linalg/hdsdp_sdpdata.c:880-899      (a <- a/|a|, sign <- sign |a|^2)
// This is our own implementation of those rules (plain C++ on std::vector), not a copy.
#include "cone.h"
#include <cmath>
#include <algorithm>

namespace {

inline long pack_col_start(long n, long j) { return j * n - j * (j - 1) / 2; }

void unpack_index(int n, long p, int &row, int &col) {
    // largest j with pack_col_start(n, j) <= p
    long lo = 0, hi = n - 1;
    while (lo < hi) {
        long mid = (lo + hi + 1) / 2;
        if (pack_col_start(n, mid) <= p) lo = mid; else hi = mid - 1;
    }
    col = (int) lo;
    row = (int) (p - pack_col_start(n, lo) + lo);
}

bool sparse_rank_one(int n, const std::vector<int> &Ai, const std::vector<int> &Aj, const std::vector<double> &Ax,
                     double &sgn, std::vector<double> &a) {
    const int nnz = (int) Ax.size();
    a.assign(n, 0.0);
    int i = Ai[0], j = Aj[0];
    double v = Ax[0];
    if (i != j) return false;
    if (nnz == 1) { sgn = Ax[0]; a[i] = 1.0; return true; }
    double s = (v > 0) ? 1.0 : -1.0;
    v = std::sqrt(std::fabs(v));
    int k = 0, anz = 0;
    for (k = 0; k < nnz; ++k) {
        if (Aj[k] > i) break;
        a[Ai[k]] = Ax[k] / v;
        anz += 1;
    }
    if (nnz != anz * (anz + 1) / 2) return false;
    if (k == n) return false;
    double eps = 0.0;
    for (k = 0; k < nnz; ++k) eps += std::fabs(Ax[k] - s * a[Ai[k]] * a[Aj[k]]);
    if (eps > 1e-10) return false;
    sgn = s;
    return true;
}

bool dense_rank_one(int n, const std::vector<double> &A, double &sgn, std::vector<double> &a) {
    int i = 0; long k = 0;
    for (i = 0; i < n; ++i) {
        if (A[k] != 0) break;
        k += n - i;
    }
    if (i == n) return false;
    double s = (A[k] > 0) ? 1.0 : -1.0;
    double v = std::sqrt(std::fabs(A[k]));
    a.assign(n, 0.0);
    // column i of the symmetric matrix: PACK_ENTRY(A, n, k, i) (the reference reads the packed slot
    // (row k, col i) also for k < i, i.e. a slot of an earlier column; we restate that literally)
    for (int r = 0; r < n; ++r) {
        long slot = (long) ((2L * n - i - 1) * i / 2) + r;
        a[r] = A[slot] / v;
    }
    double eps = 0.0; long id = 0;
    for (int c = 0; c < n; ++c) {
        for (int jj = 0; jj < n - c; ++jj) eps += std::fabs(A[id + jj] - s * a[c] * a[c + jj]);
        id += n - c;
        if (eps > 1e-10) return false;
    }
    sgn = s;
    return true;
}

} // namespace

// Classify column `k` of the user CSC ([n(n+1)/2] x [m+1], lower packed) into a HostCoeff.
void classify_coeff(int n, int nnz, const int *Ci, const double *Cx, HostCoeff &out) {
    const long npack = (long) n * (n + 1) / 2;
    out = HostCoeff();
    if (nnz == 0) { out.type = COEFF_ZERO; return; }
    std::vector<double> a;
    double sgn = 0.0;
    bool r1 = false;
    if ((double) nnz > 0.3 * (double) npack) {
        out.type = COEFF_DENSE;
        out.packed.assign(npack, 0.0);
        for (int e = 0; e < nnz; ++e) out.packed[Ci[e]] = Cx[e];
        r1 = dense_rank_one(n, out.packed, sgn, a);
    } else {
        out.type = COEFF_SPARSE;
        out.row.resize(nnz); out.col.resize(nnz); out.val.assign(Cx, Cx + nnz);
        for (int e = 0; e < nnz; ++e) unpack_index(n, Ci[e], out.row[e], out.col[e]);
        r1 = sparse_rank_one(n, out.row, out.col, out.val, sgn, a);
    }
    if (!r1) return;
    int nz = 0;
    for (int r = 0; r < n; ++r) if (std::fabs(a[r]) > 1e-10) nz += 1;
    bool dense = (double) nz > 0.5 * (double) n;
    out.row.clear(); out.col.clear(); out.val.clear(); out.packed.clear();
    out.sign = sgn;
    if (dense) {
        out.type = COEFF_DSR1;
        out.fac = a;
        double nrm = 0.0;
        for (double x : out.fac) nrm += x * x;
        nrm = std::sqrt(nrm);
        out.sign *= nrm * nrm;
        for (double &x : out.fac) x /= nrm;
    } else {
        out.type = COEFF_SPR1;
        double nrm = 0.0;
        for (int r = 0; r < n; ++r) if (std::fabs(a[r]) > 1e-10) { out.idx.push_back(r); out.fac.push_back(a[r]); nrm += a[r] * a[r]; }
        nrm = std::sqrt(nrm);
        out.sign *= nrm * nrm;
        for (double &x : out.fac) x /= nrm;
    }
}
