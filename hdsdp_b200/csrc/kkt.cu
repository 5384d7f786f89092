// hdsdp_b200/csrc/kkt.cu -- device-resident Schur complement object (hdsdp_kkt twin).
//
// Reference counterparts: interface/hdsdp_schur.c
//   HKKTIAllocDenseKKT :11-44, HKKTClean :141-165, HKKTBuildUp :256-268, HKKTExport :293-326,
//   HKKTFactorize :328-336, HKKTSolve :338-346, HKKTRegularize :348-373,
// plus the host writers into M that must have device twins:
//   sBoundConeGetKKT   interface/hdsdp_conic_bound.c:201-249   (diag(M) += 1/sl^2 + 1/su^2)
//   LPConeGetKKT       interface/hdsdp_conic_lp.c:254-330       (M += A diag(s^-2) A^T, lower)
// The solve replaces the reference's PCG-on-M (linalg/hdsdp_linsolver.c:1446-1588) by a direct
// Cholesky solve; M itself is left untouched by the factorisation (a copy is factored), matching
// the reference contract that M stays valid between HKKTFactorize and the last HKKTSolve.
#include "cone.h"
#include <cmath>
#include <cstring>

namespace {

__global__ void min_diag_kernel(const double *__restrict__ M, long ld, int m, double *out, int rank, int nranks, int nb) {
    __shared__ double red[256];
    double v = 1e300;
    for (int i = threadIdx.x; i < m; i += 256)
        if (nranks <= 1 || (i / nb) % nranks == rank) v = fmin(v, M[(long) i * ld + i]); // only owned columns are assembled
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// reg <- min(reg * mindiag, 1e-5); zero if < 1e-14; M_ii += reg   (hdsdp_schur.c:348-373)
__global__ void regularize_kernel(double *M, long ld, int m, double reg, const double *mindiag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double r = reg * (*mindiag);
    r = fmin(r, 1e-05);
    if (r < 1e-14) r = 0.0;
    M[(long) i * ld + i] += r;
}

__global__ void min_vec_kernel(const double *__restrict__ v, int n, double *out) {
    if (threadIdx.x == 0) {
        double r = v[0];
        for (int i = 1; i < n; ++i) r = fmin(r, v[i]);
        *out = r;
    }
}

// copy the lower trapezoids of the block columns owned by this rank (rows >= the block's first row)
__global__ void copy_owned_kernel(double *__restrict__ dst, const double *__restrict__ src, long ld, int mp, int nb, int rank, int nranks) {
    const int col = blockIdx.x;   // columns on gridDim.x (no 65535 limit)
    if ((col / nb) % nranks != rank) return;
    const int r0 = (col / nb) * nb;
    for (int r = r0 + blockIdx.y * blockDim.x + threadIdx.x; r < mp; r += gridDim.y * blockDim.x)
        dst[(long) col * ld + r] = src[(long) col * ld + r];
}

// M <- 0 on the part that is ever read: rows >= the first row of each 128-column block, and only the block columns this
// rank owns (HKKTClean memsets all of M, hdsdp_schur.c:156-162; the strict upper triangle is never referenced)
__global__ void zero_lower_kernel(double *M, long ld, int mp, int nb, int rank, int nranks) {
    const int col = blockIdx.x;   // columns on gridDim.x (no 65535 limit)
    if (nranks > 1 && (col / nb) % nranks != rank) return;
    const int r0 = (col / HD_LEAF) * HD_LEAF;
    double2 *p = reinterpret_cast<double2 *>(M + (long) col * ld + r0);
    const int cnt = (mp - r0) / 2;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < cnt; i += gridDim.y * blockDim.x) p[i] = make_double2(0.0, 0.0);
}

__global__ void add_diag_vec_kernel(double *M, long ld, int m, const double *__restrict__ d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) M[(long) i * ld + i] += d[i];
}

__global__ void vec_axpy_kernel(double *y, const double *__restrict__ x, int m, double a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) y[i] += a * x[i];
}

__global__ void scale_vec_kernel(double *x, int m, double a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) x[i] *= a;
}

// LP cone (reference LPConeGetKKT, interface/hdsdp_conic_lp.c:254-330): for LP column c with rows r_0..r_k (CSC by LP column),
// weight w_c = s_c^-2:  M[r_a, r_b] += a_a a_b w_c (r_a >= r_b);  asinv[r] += a / s_c ;  asinvrd[r] += rd a w_c ;
// dTraceSinv += 1 / s_c (rd != 0);  HOMOGENEOUS: dCSinv += c_c / s_c, dCSinvCSinv += (c_c / s_c)^2, asinvc[r] += a c_c w_c
__global__ void lp_schur_kernel(const int *__restrict__ colptr, const int *__restrict__ rowidx, const double *__restrict__ val,
                                const double *__restrict__ sinv, const double *__restrict__ obj, int ncol, double rd, int build_matrix,
                                int hsd, double *M, long ldm, double *asinv, double *asinvrd, double *asinvc, double *scal) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    double tr = 0.0, cs = 0.0, cscs = 0.0;
    if (c < ncol) {
        const double si = sinv[c], w = si * si;
        const double oc = hsd ? obj[c] : 0.0;
        if (rd != 0.0) tr = si;
        if (hsd) { cs = oc * si; cscs = cs * cs; }
        for (int a = colptr[c]; a < colptr[c + 1]; ++a) {
            const int ra = rowidx[a];
            const double va = val[a];
            atomicAdd(&asinv[ra], va * si);
            if (rd != 0.0) atomicAdd(&asinvrd[ra], rd * va * w);
            if (hsd) atomicAdd(&asinvc[ra], va * oc * w);
            if (!build_matrix) continue;
            for (int b = colptr[c]; b <= a; ++b) {
                const int rb = rowidx[b];
                const int r = max(ra, rb), q = min(ra, rb);
                atomicAdd(&M[(long) q * ldm + r], va * val[b] * w);
            }
        }
    }
    // block reduction of the three scalars, one atomic per block
    __shared__ double red[3][128];
    red[0][threadIdx.x] = tr; red[1][threadIdx.x] = cs; red[2][threadIdx.x] = cscs;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int q = 0; q < 3; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (rd != 0.0) atomicAdd(&scal[3], red[0][0]);
        if (hsd) { atomicAdd(&scal[1], red[1][0]); atomicAdd(&scal[0], red[2][0]); }
    }
}

// y += M x for a symmetric M of which only the lower triangle is stored (the dsymv('L') of the reference's PCG loop,
// linalg/hdsdp_linsolver.c:1446-1588), nrhs <= 2 vectors.  HBM-bound: every entry of the lower triangle is read ONCE (4 m^2 bytes)
// and used for both y_i += T x_j and y_j += T^T x_i.  One CTA (256 threads, two per SM so that one CTA's loads overlap the other's
// reductions) per 128 x 64 half-tile: thread (row, half) holds 32 entries of its tile row in registers (coalesced loads, all in
// flight together); the row product is reduced over the two halves through shared memory, the column products over the 32
// lanes by shuffles and over the 4 warps of a half through shared memory; results are added to y with atomics.
constexpr int SV_N = 64;
template <int NRHS>
__global__ void __launch_bounds__(256, 2) symv_lower_kernel(const double *__restrict__ M, long ld, int nblk, const double *__restrict__ x,
                                                           long ldx, double *y, long ldy, double alpha, double *part) {
    __shared__ double xi[NRHS][HD_LEAF], xj[NRHS][SV_N];
    __shared__ double rowred[NRHS][2][HD_LEAF];
    __shared__ double colred[NRHS][4][SV_N];
    // half-tile (i, j2): row block i, 64-column block j2 <= 2 i + 1; row block i owns 2 (i + 1) of them: i = floor((sqrt(4 b + 1) - 1) / 2)
    const long b = blockIdx.x;
    int i = (int) ((sqrt(4.0 * (double) b + 1.0) - 1.0) * 0.5);
    while ((long) (i + 1) * (i + 2) <= b) ++i;
    while ((long) i * (i + 1) > b) --i;
    const int j2 = (int) (b - (long) i * (i + 1));
    if (i >= nblk) return;
    const int t = threadIdx.x, row = t & 127, half = t >> 7, lane = t & 31, w4 = (t >> 5) & 3;
    const bool diag = (j2 >> 1) == i;
    const int c0 = (j2 & 1) * SV_N;                 // first column of this half-tile inside its leaf
    const double *T = M + ((long) j2 * SV_N + half * 32) * ld + (long) i * HD_LEAF + row;
    double tl[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) tl[q] = __ldcs(&T[(long) q * ld]);
    if (t < HD_LEAF) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) xi[r][t] = x[(long) r * ldx + (long) i * HD_LEAF + t];
    } else if (t < HD_LEAF + SV_N) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) xj[r][t - HD_LEAF] = x[(long) r * ldx + (long) j2 * SV_N + (t - HD_LEAF)];
    }
    if (diag) { // only the lower triangle of a diagonal leaf is data
#pragma unroll
        for (int q = 0; q < 32; ++q) if (c0 + half * 32 + q > row) tl[q] = 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) acc += tl[q] * xj[r][half * 32 + q];
        rowred[r][half][row] = acc;
    }
    // transposed part: column c = half*32 + q gets sum_row T[row, c] x_i[row]
    // Reduction of the 32 column products over the 32 lanes by recursive halving: at distance s a lane keeps the half of its
    // values whose index has bit s equal to its own lane bit and hands the other half to its partner -- 16 + 8 + 4 + 2 + 1 = 31
    // shuffles instead of 32 x 5, and lane q ends with the total of column q (the tile's compute phase was what kept the
    // kernel at 63 % of HBM: 2 CTAs/SM cannot hide 160 shuffles per thread behind the other CTA's loads).
    if (diag) { // on a diagonal leaf the diagonal itself was already used by the row product
#pragma unroll
        for (int q = 0; q < 32; ++q) if (c0 + half * 32 + q == row) tl[q] = 0.0;
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
        const double xv = xi[r][row];
        double v[16];
        const bool up16 = (lane & 16) != 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const double lo = tl[q] * xv, hi = tl[q + 16] * xv;
            v[q] = (up16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, up16 ? lo : hi, 16);
        }
#pragma unroll
        for (int sdist = 8; sdist >= 1; sdist >>= 1) {
            const bool up = (lane & sdist) != 0;
#pragma unroll
            for (int q = 0; q < sdist; ++q) v[q] = (up ? v[q + sdist] : v[q]) + __shfl_xor_sync(0xffffffffu, up ? v[q] : v[q + sdist], sdist);
        }
        colred[r][w4][half * 32 + lane] = v[0];
    }
    __syncthreads();
    if (part) {
        // deterministic mode: the tile's two partial vectors go to part[b][r][0..127 | 128..191]; symv_reduce_kernel adds them in order
        double *pb = part + (b * NRHS) * (HD_LEAF + SV_N);
        if (t < HD_LEAF) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) pb[r * (HD_LEAF + SV_N) + t] = rowred[r][0][t] + rowred[r][1][t];
        } else if (t < HD_LEAF + SV_N) {
            const int c = t - HD_LEAF;
#pragma unroll
            for (int r = 0; r < NRHS; ++r) pb[r * (HD_LEAF + SV_N) + HD_LEAF + c] = colred[r][0][c] + colred[r][1][c] + colred[r][2][c] + colred[r][3][c];
        }
        return;
    }
    if (t < HD_LEAF) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) atomicAdd(&y[(long) r * ldy + (long) i * HD_LEAF + t], alpha * (rowred[r][0][t] + rowred[r][1][t]));
    } else if (t < HD_LEAF + SV_N) {
        const int c = t - HD_LEAF;
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
            atomicAdd(&y[(long) r * ldy + (long) j2 * SV_N + c], alpha * (colred[r][0][c] + colred[r][1][c] + colred[r][2][c] + colred[r][3][c]));
    }
}

// y[i-block] += alpha * (sum over the half-tiles of block row i of their row parts + sum over the half-tiles of block column i of
// their column parts), in a fixed order: the result does not depend on the scheduling of symv_lower_kernel's CTAs
constexpr int SVR_G = 8; // groups of 128 threads per block row: group g adds the partial vectors of tiles g, g + 8, ...
template <int NRHS>
__global__ void __launch_bounds__(HD_LEAF * SVR_G) symv_reduce_kernel(const double *__restrict__ part, int nblk, double *y, long ldy, double alpha) {
    __shared__ double gs[SVR_G][NRHS][HD_LEAF];
    const int i = blockIdx.x, t = threadIdx.x & (HD_LEAF - 1), g = threadIdx.x >> 7;
    constexpr int PW = HD_LEAF + SV_N;
    double s[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) s[r] = 0.0;
    const long rowbase = (long) i * (i + 1);                   // first half-tile of block row i
    for (int j2 = g; j2 < 2 * (i + 1); j2 += SVR_G) {
        const double *pb = part + ((rowbase + j2) * NRHS) * PW;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) s[r] += pb[r * PW + t];
    }
    const int j2c = 2 * i + (t >> 6), c = t & 63;              // my 64-column block inside leaf i
    for (int ii = i + g; ii < nblk; ii += SVR_G) {
        const double *pb = part + (((long) ii * (ii + 1) + j2c) * NRHS) * PW;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) s[r] += pb[r * PW + HD_LEAF + c];
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) gs[g][r][t] = s[r];
    __syncthreads();
    if (g == 0) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            double v = gs[0][r][t];
#pragma unroll
            for (int q = 1; q < SVR_G; ++q) v += gs[q][r][t];  // fixed order: the sum does not depend on the schedule
            y[(long) r * ldy + (long) i * HD_LEAF + t] += alpha * v;
        }
    }
}

// r <- b - r (r holds M x), and max |r_i| / max |b_i| per vector into out[2 v], out[2 v + 1]
__global__ void residual_finish_kernel(double *r, const double *__restrict__ b, long ld, int m, int nrhs, double *out) {
    __shared__ double red[2][256];
    const int v = blockIdx.x;
    double mr = 0.0, mb = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        const double bi = b[(long) v * ld + i];
        const double ri = bi - r[(long) v * ld + i];
        r[(long) v * ld + i] = ri;
        mr = fmax(mr, fabs(ri)); mb = fmax(mb, fabs(bi));
        if (ri != ri) mr = INFINITY;
    }
    red[0][threadIdx.x] = mr; red[1][threadIdx.x] = mb;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] = fmax(red[0][threadIdx.x], red[0][threadIdx.x + o]);
            red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * v] = red[0][0]; out[2 * v + 1] = red[1][0]; }
}

__global__ void vec_add2_kernel(double *x, const double *__restrict__ d, long ld, int m, int nrhs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) for (int v = 0; v < nrhs; ++v) x[(long) v * ld + i] += d[(long) v * ld + i];
}

// ------------------------------------------------------------------------------------------------------------------------
// Jacobi-preconditioned conjugate gradients on M: the reference's DEFAULT solver of the Schur system (conjGradSolve,
// linalg/hdsdp_linsolver.c:1446-1588; HDSDP_LINSYS_DENSE_ITERATIVE, interface/hdsdp_schur.c:19-35), restated with every vector
// in HBM: one symv_lower_kernel per step (4 m^2 bytes, HBM-bound) plus one single-CTA kernel for the vector updates and dots.
// scal: 0 r.z  1 d.Md  2 alpha  3 rnew.z  4 ||r||^2  5 beta  6 ||b||^2
// ------------------------------------------------------------------------------------------------------------------------
constexpr int PCG_T = 1024;
__device__ __forceinline__ double pcg_block_sum(double v, double *red) {
    __syncthreads();
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = PCG_T / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    return red[0];
}
// x = 0, r = b, d = z = b / jac ; ||b||^2
__global__ void __launch_bounds__(PCG_T) pcg_init_kernel(const double *__restrict__ b, const double *__restrict__ M, long ld, double *jac, double *x,
                                                        double *r, double *d, double *z, int m, double *scal) {
    __shared__ double red[PCG_T];
    double nb = 0.0;
    for (int i = threadIdx.x; i < m; i += PCG_T) {
        const double bi = b[i], ji = M[(long) i * ld + i];
        jac[i] = ji; x[i] = 0.0; r[i] = bi; d[i] = bi / ji; z[i] = bi / ji;
        nb += bi * bi;
    }
    nb = pcg_block_sum(nb, red);
    if (threadIdx.x == 0) { scal[6] = nb; scal[4] = nb; }
}
// alpha = r.z / d.Md ; x += alpha d ; (full step only) rnew = r - alpha Md ; z = rnew / jac ; beta = rnew.z / r.z ; d = z + beta d
__global__ void __launch_bounds__(PCG_T) pcg_update_kernel(const double *__restrict__ jac, const double *__restrict__ Md, double *x, double *r,
                                                          double *d, double *z, int m, int restart, double *scal) {
    __shared__ double red[PCG_T];
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < m; i += PCG_T) { a += z[i] * r[i]; c += d[i] * Md[i]; }
    const double rz = pcg_block_sum(a, red);
    const double dMd = pcg_block_sum(c, red);
    const double alpha = rz / dMd;
    double e = 0.0, f = 0.0;
    for (int i = threadIdx.x; i < m; i += PCG_T) {
        x[i] += alpha * d[i];
        if (!restart) {
            const double rn = r[i] - alpha * Md[i];
            const double zn = rn / jac[i];
            r[i] = rn; z[i] = zn;
            e += rn * zn; f += rn * rn;
        }
    }
    if (restart) return;
    const double rnz = pcg_block_sum(e, red);
    const double rr = pcg_block_sum(f, red);
    const double beta = rnz / rz;
    for (int i = threadIdx.x; i < m; i += PCG_T) d[i] = z[i] + beta * d[i];
    if (threadIdx.x == 0) { scal[0] = rz; scal[1] = dMd; scal[2] = alpha; scal[3] = rnz; scal[4] = rr; scal[5] = beta; }
}
// restart: r = b - t (t = M x) ; d = z = r / jac
__global__ void __launch_bounds__(PCG_T) pcg_restart_kernel(const double *__restrict__ b, const double *__restrict__ t, const double *__restrict__ jac,
                                                           double *r, double *d, double *z, int m) {
    for (int i = threadIdx.x; i < m; i += PCG_T) {
        const double ri = b[i] - t[i];
        r[i] = ri; d[i] = ri / jac[i]; z[i] = ri / jac[i];
    }
}

inline unsigned nblk(long total, int threads) { return (unsigned) ((total + threads - 1) / threads); }

} // namespace

int hd_scale_vec(cudaStream_t st, double *x, int m, double a) {
    HDK(scale_vec_kernel)<<<nblk(m, 256), 256, 0, st>>>(x, m, a);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

static int kkt_alloc(KktCU *k) {
    size_t bytes = sizeof(double) * (size_t) k->mp * k->mp;
    if (cudaMalloc(&k->d_M, bytes) != cudaSuccess) { cudaGetLastError(); k->d_M = nullptr; return HD_MEMORY; }
    HD_CUDA(cudaMemset(k->d_M, 0, bytes));
    HD_CALL(chol_create(&k->chol, k->m));
    HD_CUDA(cudaMalloc(&k->d_asinv, sizeof(double) * k->mp));
    HD_CUDA(cudaMalloc(&k->d_asinvrd, sizeof(double) * k->mp));
    HD_CUDA(cudaMalloc(&k->d_asinvc, sizeof(double) * k->mp));
    HD_CUDA(cudaMemset(k->d_asinv, 0, sizeof(double) * k->mp));
    HD_CUDA(cudaMemset(k->d_asinvrd, 0, sizeof(double) * k->mp));
    HD_CUDA(cudaMemset(k->d_asinvc, 0, sizeof(double) * k->mp));
    HD_CUDA(cudaMalloc(&k->d_scal, sizeof(double) * 16));
    HD_CUDA(cudaMemset(k->d_scal, 0, sizeof(double) * 16));
    HD_CUDA(cudaMallocHost(&k->h_scal, sizeof(double) * 16));
    HD_CUDA(cudaMalloc(&k->d_rhs, sizeof(double) * (size_t) k->mp * 8));
    HD_CUDA(cudaMemset(k->d_rhs, 0, sizeof(double) * (size_t) k->mp * 8));
    HD_CUDA(cudaMallocHost(&k->h_vec, sizeof(double) * (size_t) k->mp * 8));
    return HD_OK;
}

int kkt_create(KktCU **pk, int nRow) {
    KktCU *k = new KktCU();
    k->m = nRow;
    k->mp = hd_pad(nRow);
    const int rc = kkt_alloc(k);
    if (rc != HD_OK) { kkt_destroy(k); return rc; } // a failure half-way (20 GB for M at m = 50k) releases what exists already
    *pk = k;
    return HD_OK;
}

void kkt_destroy(KktCU *k) {
    if (!k) return;
    cudaFree(k->d_M); cudaFree(k->d_asinv); cudaFree(k->d_asinvrd); cudaFree(k->d_asinvc);
    cudaFree(k->d_scal); cudaFree(k->d_rhs);
    if (k->h_scal) cudaFreeHost(k->h_scal);
    if (k->h_vec) cudaFreeHost(k->h_vec);
    if (k->dist) dist_destroy(k->dist);
    if (k->d_gather) cudaFree(k->d_gather);
    if (k->d_ref) cudaFree(k->d_ref);
    if (k->d_pcg) cudaFree(k->d_pcg);
    if (k->d_symv_ws) cudaFree(k->d_symv_ws);
    chol_destroy(k->chol);
    delete k;
}

// HKKTClean (hdsdp_schur.c:141-165)
int kkt_clean(KktCU *k, int typeKKT) {
    cudaStream_t st = hd_stream();
    HD_CUDA(cudaMemsetAsync(k->d_asinv, 0, sizeof(double) * k->mp, st));
    HD_CUDA(cudaMemsetAsync(k->d_asinvrd, 0, sizeof(double) * k->mp, st));
    if (typeKKT == KKT_HOMOGENEOUS) {
        HD_CUDA(cudaMemsetAsync(k->d_asinvc, 0, sizeof(double) * k->mp, st));
        HD_CUDA(cudaMemsetAsync(k->d_scal, 0, sizeof(double) * 3, st));
    }
    HD_CUDA(cudaMemsetAsync(k->d_scal + 3, 0, sizeof(double), st));
    if (typeKKT == KKT_INFEASIBLE || typeKKT == KKT_HOMOGENEOUS || typeKKT == KKT_PRIMAL) {
        HDK(zero_lower_kernel)<<<dim3(k->mp, 16), 256, 0, st>>>(k->d_M, k->mp, k->mp, k->shard_nb, k->rank, k->nranks);
        HD_CUDA(cudaGetLastError());
        k->factored = false;
        k->fresh = true;
    }
    return HD_OK;
}

int kkt_build_up(KktCU *k, int typeKKT) {
    HD_CALL(kkt_clean(k, typeKKT));
    for (size_t i = 0; i < k->cones.size(); ++i) {
        HD_CALL(cone_build_schur(k->cones[i], (int) i, k, typeKKT));
        k->fresh = false;
    }
    return HD_OK;
}

int kkt_regularize(KktCU *k, double reg) {
    cudaStream_t st = hd_stream();
    HDK(min_diag_kernel)<<<1, 256, 0, st>>>(k->d_M, k->mp, k->m, k->d_scal + 4, k->rank, k->nranks, k->shard_nb);
    if (k->dist && k->nranks > 1) {
        // global min of diag(M) over the ranks' column shards (peer-memory all-gather of one double, dist.cu)
        HD_CALL(dist_allgather_small(k->dist, k->d_scal + 4, 1, k->d_gather));
        HDK(min_vec_kernel)<<<1, 32, 0, st>>>(k->d_gather, k->nranks, k->d_scal + 4);
    }
    HDK(regularize_kernel)<<<nblk(k->m, 256), 256, 0, st>>>(k->d_M, k->mp, k->m, reg, k->d_scal + 4);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// Host-computed contributions of the cheap cones (bound cone, scalar parts of the LP cone):
// diag(M) += diagAdd, dASinvVec += asinvAdd, dASinvRdSinvVec += asinvrdAdd, dASinvCSinvVec += asinvcAdd,
// scalars {dCSinvCSinv, dCSinv, dCSinvRdSinv, dTraceSinv} += scalarsAdd4.  Any pointer may be NULL.
int kkt_add_host(KktCU *k, const double *diagAdd, const double *asinvAdd, const double *asinvrdAdd, const double *asinvcAdd,
                 const double *scalarsAdd4) {
    cudaStream_t st = hd_stream();
    const int m = k->m;
    k->fresh = false;
    const double *src[4] = {diagAdd, asinvAdd, asinvrdAdd, asinvcAdd};
    double *dst[4] = {nullptr, k->d_asinv, k->d_asinvrd, k->d_asinvc};
    for (int v = 0; v < 4; ++v) {
        if (!src[v]) continue;
        memcpy(k->h_vec + (size_t) v * k->mp, src[v], sizeof(double) * m);
        HD_CUDA(cudaMemcpyAsync(k->d_rhs + (size_t) v * k->mp, k->h_vec + (size_t) v * k->mp, sizeof(double) * m,
                                cudaMemcpyHostToDevice, st));
        if (v == 0) HDK(add_diag_vec_kernel)<<<nblk(m, 256), 256, 0, st>>>(k->d_M, k->mp, m, k->d_rhs);
        else HDK(vec_axpy_kernel)<<<nblk(m, 256), 256, 0, st>>>(dst[v], k->d_rhs + (size_t) v * k->mp, m, 1.0);
    }
    if (scalarsAdd4) {
        memcpy(k->h_vec + (size_t) 4 * k->mp, scalarsAdd4, sizeof(double) * 4);
        HD_CUDA(cudaMemcpyAsync(k->d_rhs + (size_t) 4 * k->mp, k->h_vec + (size_t) 4 * k->mp, sizeof(double) * 4,
                                cudaMemcpyHostToDevice, st));
        HDK(vec_axpy_kernel)<<<1, 32, 0, st>>>(k->d_scal, k->d_rhs + (size_t) 4 * k->mp, 4, 1.0);
    }
    HD_CUDA(cudaGetLastError());
    HD_CUDA(cudaStreamSynchronize(st)); // h_vec is reused by the next call
    return HD_OK;
}

// LP cone twin.  colptr/rowidx/val: CSC by LP column of the [nLpCol x m]^T data (row index = constraint);
// sInvHost: 1/s per LP column at the current iterate.
int kkt_add_lp(KktCU *k, int nLpCol, const int *d_colptr, const int *d_rowidx, const double *d_val, const double *d_obj,
               const double *sInvHost, double *d_sinv_stage, double rd, int typeKKT) {
    cudaStream_t st = hd_stream();
    k->fresh = false;
    HD_CUDA(cudaMemcpyAsync(d_sinv_stage, sInvHost, sizeof(double) * nLpCol, cudaMemcpyHostToDevice, st));
    HDK(lp_schur_kernel)<<<nblk(nLpCol, 128), 128, 0, st>>>(d_colptr, d_rowidx, d_val, d_sinv_stage, d_obj, nLpCol, rd,
                                                        typeKKT != KKT_CORRECTOR, typeKKT == KKT_HOMOGENEOUS, k->d_M, k->mp, k->d_asinv,
                                                        k->d_asinvrd, k->d_asinvc, k->d_scal);
    HD_CUDA(cudaGetLastError());
    HD_CUDA(cudaStreamSynchronize(st));
    return HD_OK;
}

// HKKTExport (hdsdp_schur.c:293-326): D2H of the side vectors and scalars
int kkt_export(KktCU *k, double *asinv, double *asinvrd, double *asinvc, double *csinvcsinv, double *csinv,
               double *csinvrd, double *tracesinv) {
    cudaStream_t st = hd_stream();
    const int m = k->m;
    if (asinv) HD_CUDA(cudaMemcpyAsync(k->h_vec, k->d_asinv, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    if (asinvrd) HD_CUDA(cudaMemcpyAsync(k->h_vec + k->mp, k->d_asinvrd, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    if (asinvc) HD_CUDA(cudaMemcpyAsync(k->h_vec + 2 * (size_t) k->mp, k->d_asinvc, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    HD_CUDA(cudaMemcpyAsync(k->h_scal, k->d_scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
    HD_CUDA(cudaStreamSynchronize(st));
    if (asinv) memcpy(asinv, k->h_vec, sizeof(double) * m);
    if (asinvrd) memcpy(asinvrd, k->h_vec + k->mp, sizeof(double) * m);
    if (asinvc) memcpy(asinvc, k->h_vec + 2 * (size_t) k->mp, sizeof(double) * m);
    if (csinvcsinv) *csinvcsinv = k->h_scal[0];
    if (csinv) *csinv = k->h_scal[1];
    if (csinvrd) *csinvrd = k->h_scal[2];
    if (tracesinv) *tracesinv = k->h_scal[3];
    return HD_OK;
}

// HKKTFactorize: copy M (lower) into the factor buffer, pad, Cholesky.  info > 0 -> HD_FAILED
// (the reference falls back to dsytrf LDL there, hdsdp_linsolver.c:2036-2039: here the LDL^T mode of chol.cu, see below)
int kkt_factorize(KktCU *k, int *info_out) {
    cudaStream_t st = hd_stream();
    if (k->solver_mode == 1 && k->use_jacobi && !(k->dist && k->nranks > 1)) {
        // conjGradLinSolverNumeric with the Jacobi preconditioner (:1359-1372, :1405-1431): nothing to factor, diag(M) is read by the solve
        if (info_out) *info_out = 0;
        k->factored = true;
        return HD_OK;
    }
    if (k->dist && k->nranks > 1) {
        // multi-GPU: this rank's block columns of M go into the factor buffer, the peers' columns arrive as factor panels
        HDK(copy_owned_kernel)<<<dim3(k->mp, 8), 256, 0, st>>>(k->chol->L, k->d_M, k->mp, k->mp, k->shard_nb, k->rank, k->nranks);
        HD_CUDA(cudaGetLastError());
        HD_CALL(hd_pad_identity(st, k->chol->L, k->mp, k->m, k->mp));
        int info = 0;
        int rc = dist_factor(k->dist, &info, k->chol->ldl);
        if (rc == HD_OK && info > 0 && !k->chol->ldl) {
            // every rank sees the same info (it travels with the panels): all switch to LDL^T together, for good
            fprintf(stderr, "[hdsdpcu] KKT system is almost indefinite (pivot %d). Switch to LDL.\n", info);
            HDK(copy_owned_kernel)<<<dim3(k->mp, 8), 256, 0, st>>>(k->chol->L, k->d_M, k->mp, k->mp, k->shard_nb, k->rank, k->nranks);
            HD_CALL(hd_pad_identity(st, k->chol->L, k->mp, k->m, k->mp));
            rc = dist_factor(k->dist, &info, true);
        }
        if (info_out) *info_out = info;
        k->factored = (rc == HD_OK && info == 0);
        return k->factored ? HD_OK : HD_FAILED;
    }
    // only the lower trapezoids are ever read by the factorisation: half the traffic of a full copy
    HDK(copy_owned_kernel)<<<dim3(k->mp, 8), 256, 0, st>>>(k->chol->L, k->d_M, k->mp, k->mp, HD_LEAF, 0, 1);
    HD_CUDA(cudaGetLastError());
    HD_CALL(hd_pad_identity(st, k->chol->L, k->mp, k->m, k->mp));
    int info = 0;
    HD_CALL(chol_factor(st, k->chol, &info));
    if (info > 0 && !k->chol->ldl) {
        // reference HFpLinsysNumeric (linalg/hdsdp_linsolver.c:2030-2039): "KKT system is almost indefinite. Switch to LDL."
        // -- permanently, as HFpLinsysSwitchToIndefinite does.  Here: unpivoted LDL^T with static pivoting on the GPU.
        fprintf(stderr, "[hdsdpcu] KKT system is almost indefinite (pivot %d). Switch to LDL.\n", info);
        k->chol->ldl = true;
        HDK(copy_owned_kernel)<<<dim3(k->mp, 8), 256, 0, st>>>(k->chol->L, k->d_M, k->mp, k->mp, HD_LEAF, 0, 1);
        HD_CALL(hd_pad_identity(st, k->chol->L, k->mp, k->m, k->mp));
        HD_CALL(chol_factor(st, k->chol, &info));
    }
    if (info_out) *info_out = info;
    k->factored = (info == 0);
    return info == 0 ? HD_OK : HD_FAILED;
}

// y (nrhs vectors of stride ldy) += alpha * A x for a symmetric np x np matrix of which the lower triangle is stored (np % 128 == 0,
// vectors zero-padded to np): used for the Schur matrix (PCG, residuals) and for the dual step dS of the Lanczos operator
// ws: device workspace of hd_symv_ws_doubles(np) doubles for the deterministic (ordered) reduction; null = atomics
long hd_symv_ws_doubles(int np) {
    const long nb = np / HD_LEAF;
    return nb * (nb + 1) * 2 * (HD_LEAF + SV_N);
}
int hd_symv_lower(cudaStream_t st, const double *A, long ld, int np, const double *d_x, long ldx, double *d_y, long ldy, int nRhs, double alpha,
                  double *ws) {
    const int nb = np / HD_LEAF;
    const long tiles = (long) nb * (nb + 1);   // 128 x 64 half-tiles of the lower triangle
    for (int r0 = 0; r0 < nRhs; r0 += 2) {
        const int nr = (nRhs - r0 >= 2) ? 2 : 1;
        g_hd_launches += ws ? 2 : 1;
        if (nr == 2) {
            symv_lower_kernel<2><<<(unsigned) tiles, 256, 0, st>>>(A, ld, nb, d_x + (long) r0 * ldx, ldx, d_y + (long) r0 * ldy, ldy, alpha, ws);
            if (ws) symv_reduce_kernel<2><<<nb, HD_LEAF * SVR_G, 0, st>>>(ws, nb, d_y + (long) r0 * ldy, ldy, alpha);
        } else {
            symv_lower_kernel<1><<<(unsigned) tiles, 256, 0, st>>>(A, ld, nb, d_x + (long) r0 * ldx, ldx, d_y + (long) r0 * ldy, ldy, alpha, ws);
            if (ws) symv_reduce_kernel<1><<<nb, HD_LEAF * SVR_G, 0, st>>>(ws, nb, d_y + (long) r0 * ldy, ldy, alpha);
        }
    }
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// y (nrhs vectors of stride mp) += M x, M = the assembled Schur matrix (lower triangle), single GPU only
int kkt_symv_dev(KktCU *k, const double *d_x, double *d_y, int nRhs) {
    if (k->dist && k->nranks > 1) return HD_FAILED; // every rank holds only its own block columns of M
    if (!k->d_symv_ws) HD_CUDA(cudaMalloc(&k->d_symv_ws, sizeof(double) * (size_t) hd_symv_ws_doubles(k->mp)));
    return hd_symv_lower(hd_stream(), k->d_M, k->mp, k->mp, d_x, k->mp, d_y, k->mp, nRhs, 1.0, k->d_symv_ws);
}


// The reference's iteration limits and tolerances for the Schur system (HKKTIAllocDenseKKT, interface/hdsdp_schur.c:21-35;
// defaults of conjGradLinSolverCreate, linalg/hdsdp_linsolver.c:1313-1318)
static void pcg_params(int m, double *absTol, double *relTol, int *maxIter) {
    double acc = 1e-12; // KKT_ACCURACY
    int it = -1;
    if (m > 20000) { acc *= 100.0; it = 500; }
    else if (m > 15000) { acc *= 50.0; it = 450; }
    else if (m > 5000) { acc *= 5.0; it = 120; }
    *absTol = acc; *relTol = 5.0 * acc;
    *maxIter = it > 0 ? it : (m / 20 > 50 ? m / 20 : 50);
}

// conjGradSolve with the Jacobi preconditioner on one device vector (in place).  Returns HD_OK with *converged = 0 when the
// reference would give up on Jacobi (iteration limit, or iter > 20 with ||r|| > 0.01 ||b||): the caller then switches to Cholesky.
static int pcg_jacobi_dev(KktCU *k, double *d_x, int *converged) {
    cudaStream_t st = hd_stream();
    const int m = k->m, mp = k->mp;
    if (!k->d_pcg) {
        HD_CUDA(cudaMalloc(&k->d_pcg, sizeof(double) * ((size_t) 8 * mp + 8)));
        HD_CUDA(cudaMemset(k->d_pcg, 0, sizeof(double) * ((size_t) 8 * mp + 8)));
    }
    double *jac = k->d_pcg, *x = jac + mp, *r = x + mp, *d = r + mp, *z = d + mp, *Md = z + mp, *b = Md + mp, *t = b + mp, *scal = t + mp;
    double absTol, relTol; int maxIter;
    pcg_params(m, &absTol, &relTol, &maxIter);
    HD_CUDA(cudaMemcpyAsync(b, d_x, sizeof(double) * mp, cudaMemcpyDeviceToDevice, st));
    HDK(pcg_init_kernel)<<<1, PCG_T, 0, st>>>(b, k->d_M, k->mp, jac, x, r, d, z, m, scal);
    HD_CUDA(cudaMemcpyAsync(k->h_scal + 8, scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
    HD_CUDA(cudaStreamSynchronize(st));
    const double rhsNorm = sqrt(k->h_scal[8 + 6]);
    double cgTol = absTol < rhsNorm * relTol ? absTol : rhsNorm * relTol;
    if (cgTol < 0.1 * absTol) cgTol = 0.1 * absTol;
    *converged = 1;
    k->last_cg_iters = 0;
    if (rhsNorm < cgTol) { // the zero vector already qualifies (hdsdp_linsolver.c:1489-1492)
        HD_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * mp, st));
        return HD_OK;
    }
    HD_CUDA(cudaMemsetAsync(Md, 0, sizeof(double) * mp, st));
    HD_CALL(kkt_symv_dev(k, d, Md, 1));
    int iter = 0;
    bool giveup = false;
    for (iter = 0; iter < maxIter; ++iter) {
        const bool restart = (iter % 20 == 5);
        HDK(pcg_update_kernel)<<<1, PCG_T, 0, st>>>(jac, Md, x, r, d, z, m, restart ? 1 : 0, scal);
        if (restart) { // r = b - M x from scratch (hdsdp_linsolver.c:1509-1523)
            HD_CUDA(cudaMemsetAsync(t, 0, sizeof(double) * mp, st));
            HD_CALL(kkt_symv_dev(k, x, t, 1));
            HDK(pcg_restart_kernel)<<<1, PCG_T, 0, st>>>(b, t, jac, r, d, z, m);
            HD_CUDA(cudaMemsetAsync(Md, 0, sizeof(double) * mp, st));
            HD_CALL(kkt_symv_dev(k, d, Md, 1));
            continue;
        }
        HD_CUDA(cudaMemcpyAsync(k->h_scal + 8, scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
        HD_CUDA(cudaStreamSynchronize(st));
        const double resiNorm = sqrt(k->h_scal[8 + 4]);
        if (resiNorm != resiNorm) { giveup = true; break; }           // ITERATIVE_STATUS_NUMERICAL
        if (iter > 20 && resiNorm > 0.01 * rhsNorm) { giveup = true; break; }
        if (resiNorm < cgTol) break;
        HD_CUDA(cudaMemsetAsync(Md, 0, sizeof(double) * mp, st));
        HD_CALL(kkt_symv_dev(k, d, Md, 1));
    }
    k->last_cg_iters = iter;
    if (giveup || iter >= maxIter) { *converged = 0; return HD_OK; }
    HD_CUDA(cudaMemcpyAsync(d_x, x, sizeof(double) * mp, cudaMemcpyDeviceToDevice, st));
    return HD_OK;
}

static int solve_once(KktCU *k, cudaStream_t st, double *d_x, int nRhs) {
    HD_CALL(chol_fsolve(st, k->chol, d_x, nRhs, k->mp));
    HD_CALL(chol_dsolve(st, k->chol, d_x, nRhs, k->mp));
    return chol_bsolve(st, k->chol, d_x, nRhs, k->mp);
}

// nRhs device vectors of stride mp (padded with zeros) solved in place.
// LDL^T mode (static pivoting, no symmetric pivoting) on one GPU: M itself is intact, so the solve is followed by fixed-precision
// iterative refinement on the residual b - M x (at most 3 steps) and FAILS if the residual does not come down to
// 1e-9 max|b| -- a perturbed pivot or element growth can then never return an inaccurate Newton direction silently.
int kkt_solve_dev(KktCU *k, double *d_x, int nRhs) {
    if (!k->factored) return HD_FAILED;
    cudaStream_t st = hd_stream();
    if (k->solver_mode == 1 && k->use_jacobi && !(k->dist && k->nranks > 1)) {
        // reference policy: Jacobi-PCG on M; the first failure switches to the Cholesky factor for good (:1558-1567)
        for (int v = 0; v < nRhs; ++v) {
            int ok = 0;
            HD_CALL(pcg_jacobi_dev(k, d_x + (size_t) v * k->mp, &ok));
            k->cg_solves += 1;
            if (!ok) {
                k->use_jacobi = false;
                k->cg_fallbacks += 1;
                HD_CALL(kkt_factorize(k, nullptr));     // M is intact: factor it now, then solve the remaining vectors directly
                return kkt_solve_dev(k, d_x + (size_t) v * k->mp, nRhs - v);
            }
        }
        return HD_OK;
    }
    const bool refine = k->chol->ldl && !(k->dist && k->nranks > 1);
    if (!refine) return solve_once(k, st, d_x, nRhs);
    if (nRhs > 4) { // the refinement workspace holds 4 right-hand sides
        for (int r0 = 0; r0 < nRhs; r0 += 4) HD_CALL(kkt_solve_dev(k, d_x + (size_t) r0 * k->mp, nRhs - r0 < 4 ? nRhs - r0 : 4));
        return HD_OK;
    }
    const size_t bytes = sizeof(double) * (size_t) k->mp * nRhs;
    if (!k->d_ref) HD_CUDA(cudaMalloc(&k->d_ref, sizeof(double) * (size_t) k->mp * 8));
    double *d_b = k->d_ref, *d_r = k->d_ref + (size_t) 4 * k->mp;
    HD_CUDA(cudaMemcpyAsync(d_b, d_x, bytes, cudaMemcpyDeviceToDevice, st));
    HD_CALL(solve_once(k, st, d_x, nRhs));
    k->last_refine_steps = 0;
    for (int it = 0; it < 4; ++it) {
        HD_CUDA(cudaMemsetAsync(d_r, 0, bytes, st));
        HD_CALL(kkt_symv_dev(k, d_x, d_r, nRhs));
        HDK(residual_finish_kernel)<<<nRhs, 256, 0, st>>>(d_r, d_b, k->mp, k->m, nRhs, k->d_scal + 8);
        HD_CUDA(cudaMemcpyAsync(k->h_scal + 8, k->d_scal + 8, sizeof(double) * 2 * nRhs, cudaMemcpyDeviceToHost, st));
        HD_CUDA(cudaStreamSynchronize(st));
        double worst = 0.0;
        for (int v = 0; v < nRhs; ++v) {
            const double rel = k->h_scal[8 + 2 * v] / (k->h_scal[8 + 2 * v + 1] > 0.0 ? k->h_scal[8 + 2 * v + 1] : 1.0);
            if (!(rel <= worst)) worst = rel;   // NaN propagates
        }
        k->last_residual = worst;
        if (worst <= 1e-13) return HD_OK;
        if (it == 3) break;
        HD_CALL(solve_once(k, st, d_r, nRhs));
        HDK(vec_add2_kernel)<<<nblk(k->m, 256), 256, 0, st>>>(d_x, d_r, k->mp, k->m, nRhs);
        k->last_refine_steps = it + 1;
    }
    if (!(k->last_residual <= 1e-9)) {
        fprintf(stderr, "[hdsdpcu] LDL^T solve of the Schur system: residual %.2e after %d refinement steps (%d perturbed pivots): FAILED\n",
                k->last_residual, k->last_refine_steps, k->chol->nperturbed);
        return HD_FAILED;
    }
    return HD_OK;
}

// HKKTSolve: host rhs -> host lhs (lhs == NULL: in place), nRhs <= 8 per call
int kkt_solve(KktCU *k, int nRhs, const double *rhs, double *lhs) {
    if (!k->factored) return HD_FAILED;
    cudaStream_t st = hd_stream();
    const int m = k->m;
    for (int r0 = 0; r0 < nRhs; r0 += 8) {
        int nb = (nRhs - r0 < 8) ? nRhs - r0 : 8;
        memset(k->h_vec, 0, sizeof(double) * (size_t) k->mp * nb);
        for (int r = 0; r < nb; ++r) memcpy(k->h_vec + (size_t) r * k->mp, rhs + (size_t) (r0 + r) * m, sizeof(double) * m);
        HD_CUDA(cudaMemcpyAsync(k->d_rhs, k->h_vec, sizeof(double) * (size_t) k->mp * nb, cudaMemcpyHostToDevice, st));
        int rc = kkt_solve_dev(k, k->d_rhs, nb);
        if (rc == HD_OK) {
            HD_CUDA(cudaMemcpyAsync(k->h_vec, k->d_rhs, sizeof(double) * (size_t) k->mp * nb, cudaMemcpyDeviceToHost, st));
            HD_CUDA(cudaStreamSynchronize(st));
        }
        // reference HFpLinsysSolve (linalg/hdsdp_linsolver.c:2088-2103): a failed solve or a NaN in sol[0] / rhs[0] makes the dense
        // back-end of M switch to the indefinite factorisation for good ("KKT system is unstable. Switch to LDL.") and solve again
        const bool bad = rc != HD_OK || k->h_vec[0] != k->h_vec[0] || rhs[(size_t) r0 * m] != rhs[(size_t) r0 * m];
        if (bad && !k->chol->ldl && !(k->dist && k->nranks > 1)) {
            fprintf(stderr, "[hdsdpcu] KKT system is unstable. Switch to LDL.\n");
            k->chol->ldl = true;
            HD_CALL(kkt_factorize(k, nullptr));
            memset(k->h_vec, 0, sizeof(double) * (size_t) k->mp * nb);
            for (int r = 0; r < nb; ++r) memcpy(k->h_vec + (size_t) r * k->mp, rhs + (size_t) (r0 + r) * m, sizeof(double) * m);
            HD_CUDA(cudaMemcpyAsync(k->d_rhs, k->h_vec, sizeof(double) * (size_t) k->mp * nb, cudaMemcpyHostToDevice, st));
            HD_CALL(kkt_solve_dev(k, k->d_rhs, nb));
            HD_CUDA(cudaMemcpyAsync(k->h_vec, k->d_rhs, sizeof(double) * (size_t) k->mp * nb, cudaMemcpyDeviceToHost, st));
            HD_CUDA(cudaStreamSynchronize(st));
        } else if (bad) {
            return HD_FAILED;
        }
        double *out = lhs ? lhs : const_cast<double *>(rhs);
        for (int r = 0; r < nb; ++r) memcpy(out + (size_t) (r0 + r) * m, k->h_vec + (size_t) r * k->mp, sizeof(double) * m);
    }
    return HD_OK;
}

int kkt_get_matrix(KktCU *k, double *Mhost) {
    cudaStream_t st = hd_stream();
    HD_CUDA(cudaMemcpy2DAsync(Mhost, (size_t) k->m * 8, k->d_M, (size_t) k->mp * 8, (size_t) k->m * 8, k->m,
                              cudaMemcpyDeviceToHost, st));
    HD_CUDA(cudaStreamSynchronize(st));
    return HD_OK;
}
