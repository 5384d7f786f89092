// hdsdp_b200/csrc/chol.cu -- dense FP64 Cholesky / inverse / triangular solves on sm_100a.
//
// Replaces the LAPACK calls behind the reference's dense linear-system back-end
// (reference linalg/hdsdp_linsolver.c: dpotrf :1096, dtrsm :1158/:1184, dpotrs :1210,
//  dpotri + symmetrise :1238-1260, diag extraction :1227).
//
// Algorithm: cache-oblivious *recursive* right-looking Cholesky on a matrix padded to a multiple
// of 128.  All O(n^3) work is the DMMA GEMM of gemm_nt.cu:
//     potrf(A)      : potrf(A11); A21 <- A21 L11^-T; A22 -= A21 A21^T (lower tiles); potrf(A22)
//     trsm(B, L)    : B1 <- B1 L11^-T; B2 -= B1 L21^T; B2 <- B2 L22^-T
//     leaf (128)    : one CTA factors the 128x128 block in shared memory and also emits its
//                     explicit inverse, so every leaf triangular solve is again a GEMM / GEMV.
// The recursion runs on the host and only enqueues kernels on one stream; shapes are fixed per
// matrix size, so a whole factorisation can be captured into a CUDA graph by the caller.
#include "common.h"
#include <cuda.h>
#include <cmath>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// Leaf: Cholesky of one 128x128 block + its inverse.  One CTA, 512 threads, 128 KB smem.
// ------------------------------------------------------------------------------------------
constexpr int LEAF_THREADS = 512;

__global__ void __launch_bounds__(LEAF_THREADS, 1)
potf2_leaf_kernel(double *A, long lda, double *Dinv, int *info, int base) {
    extern __shared__ __align__(16) double Ls[]; // 128 x 128 column-major, + 128 column buffer
    double *colbuf = Ls + HD_LEAF * HD_LEAF;
    const int tid = threadIdx.x;
    // load lower triangle (upper := 0)
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += LEAF_THREADS) {
        int i = e & 127, k = e >> 7;
        Ls[e] = (i >= k) ? A[(long) k * lda + i] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < HD_LEAF; ++j) {
        double d = Ls[j * HD_LEAF + j];
        bool bad = !(d > 0.0) || isinf(d);
        if (bad) {
            if (tid == 0) atomicCAS(info, 0, base + j + 1);
            d = 1.0;
        }
        double r = sqrt(d);
        double rinv = 1.0 / r;
        __syncthreads(); // everyone has read the pivot
        for (int i = j + tid; i < HD_LEAF; i += LEAF_THREADS) {
            Ls[j * HD_LEAF + i] = (i == j) ? r : Ls[j * HD_LEAF + i] * rinv;
        }
        __syncthreads();
        // trailing update of columns k = j+1 .. 127, rows i >= k : thread -> (row ti, column phase tk)
        {
            const int ti = tid & 127, tk = tid >> 7;
            if (ti > j) {
                const double lij = Ls[j * HD_LEAF + ti];
                for (int k = j + 1 + tk; k <= ti; k += 4) Ls[k * HD_LEAF + ti] -= lij * Ls[j * HD_LEAF + k];
            }
        }
        __syncthreads();
    }
    // write back L (lower only; strict upper of the global block is left untouched)
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += LEAF_THREADS) {
        int i = e & 127, k = e >> 7;
        if (i >= k) A[(long) k * lda + i] = Ls[e];
    }
    __syncthreads();
    // in-place inverse of the lower-triangular factor (column sweep from the last column):
    //   inv[j][j] = 1/L[j][j];  inv[j+1:, j] = -inv[j][j] * (inv[j+1:, j+1:] * L[j+1:, j])
    // 4 threads cooperate on one row (k-range split), 128 rows x 4 = 512 threads.
    const int row = tid >> 2, part = tid & 3;
    for (int j = HD_LEAF - 1; j >= 0; --j) {
        if (tid < HD_LEAF) colbuf[tid] = Ls[j * HD_LEAF + tid]; // copy column j of L
        __syncthreads();
        double djj = 1.0 / colbuf[j];
        double s = 0.0;
        if (row > j) {
            // x_row = sum_{k=j+1}^{row} inv[row][k] * L[k][j]
            for (int k = j + 1 + part; k <= row; k += 4) s += Ls[k * HD_LEAF + row] * colbuf[k];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
            if (row > j) Ls[j * HD_LEAF + row] = -djj * s;
            else if (row == j) Ls[j * HD_LEAF + row] = djj;
        }
        __syncthreads();
    }
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += LEAF_THREADS) Dinv[e] = Ls[e]; // upper part is exactly 0
}


// ------------------------------------------------------------------------------------------
// Leaf v2: the same contract (L and L^-1 of one 128x128 block), organised for the tensor pipe.
// 8 panels of 16 columns.  Per panel: (a) warp 0 factors the 16x16 diagonal block and inverts it with register
// rows + shuffles; (b) the rows below become X = A21 W^T and (c) the trailing block gets -= X X^T, both as
// m8n8k4 DMMA tiles on shared memory (row stride 132: conflict-free fragment loads).  The inverse is then built
// block column by block column, one warp per block column with no CTA-wide barrier:
//   X_rc = -W_r * sum_{k=c}^{r-1} L_rk X_kc   (X_cc = W_c),   stored transposed in the unused upper triangle.
// ------------------------------------------------------------------------------------------
constexpr int L2_LD = 132, L2_WLD = 20, L2_THREADS = 256;
constexpr int L2_SMEM = (HD_LEAF * L2_LD + 8 * 16 * L2_WLD + 8 * 16 * L2_WLD + 2 * HD_LEAF) * 8;

__device__ __forceinline__ void dmma_tile(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// phase clocks of the leaf (tools/leafclk.py): compiled in only with -DHDSDPCU_LEAFCLK (stores into one global array would
// race between leaves running concurrently on the main and side streams and add traffic to the latency-critical kernel)
__device__ long long g_leaf_clk[40];
#ifdef HDSDPCU_LEAFCLK
#define LEAF_CLK(i) do { if (tid == 0) g_leaf_clk[i] = clock64(); } while (0)
#else
#define LEAF_CLK(i) do { } while (0)
#endif

// LDL = true: signed Cholesky A = L J L^T with J = diag(+-1) (unpivoted LDL^T with D = J, the scale folded into L);
// pivots with |d| <= *floorp are replaced by +floorp (static pivoting) and counted.  sgn[j] receives J_jj.
template <bool LDL>
__global__ void __launch_bounds__(L2_THREADS, 1)
potf2_leaf2_kernel(double *A, long lda, double *Dinv, int *info, int base, double *sgn, const double *floorp, int *nperturb) {
    extern __shared__ __align__(16) double sm[];
    double *As = sm;                          // 128 x 132, column-major: L in the lower triangle, X^T in the upper
    double *Wd = sm + HD_LEAF * L2_LD;        // 8 diagonal-block inverses, W[row][col] at col * 20 + row
    double *Sc = Wd + 8 * 16 * L2_WLD;        // per-warp 16 x 16 scratch, same layout
    double *Ri = Sc + 8 * 16 * L2_WLD;        // 1 / L_jj
    double *Sg = Ri + HD_LEAF;                // J_jj (LDL only)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    LEAF_CLK(0);
    // 16 independent loads in flight per thread (a plain loop is latency-bound: one HBM round trip per element)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double tmp[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int e = tid + L2_THREADS * (16 * b + u), i = e & 127, j = e >> 7;
            tmp[u] = (i >= j) ? __ldcs(&A[(long) j * lda + i]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int e = tid + L2_THREADS * (16 * b + u), i = e & 127, j = e >> 7;
            As[j * L2_LD + i] = tmp[u];
        }
    }
    __syncthreads();
    LEAF_CLK(1);
    // (a) 16 x 16 diagonal block c0: lane l (and its mirror l + 16) holds row l in registers; column j is broadcast with
    //     shuffles.  No branches in the update: entries above the diagonal hold garbage that is never read.  1 / L_jj comes from
    //     rsqrt (<= 1 ulp) and is kept for the substitutions.  Executed by warp 0 only.
    auto factor_diag = [&](const int c0) {
        // Lane l (and its mirror l + 16) holds row l of the block in registers; fully unrolled over the 16 pivots.  Column j is
        // broadcast through shared memory (double-buffered): 15 warp-wide double shuffles per pivot cost ~120 issue cycles on the
        // single warp that runs this chain, 8 broadcast LDS.128 cost ~16.  Entries above the diagonal hold garbage that is
        // never read.  Measured alternatives (tools/leafclk.py, cycles per 16 x 16 block, hot / first execution of a launch):
        // this version 4.4k / 19k (the unrolled code is fetched cold by every launch); rolled loop with a shifting register row
        // 7.9k / 8.6k; rolled loop in place on shared memory 14k / 15k.
        const int l = lane & 15;
        double a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = As[(c0 + k) * L2_LD + c0 + l];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            double d = __shfl_sync(0xffffffffu, a[j], j);
            double sj = 1.0;
            if (LDL) {
                if (d < 0.0) { sj = -1.0; d = -d; }
                const double fl = *floorp;
                if (!(d > fl) || isinf(d)) { // tiny, zero or NaN pivot: static pivoting
                    if (lane == 0) atomicAdd(nperturb, 1);
                    d = fl; sj = 1.0;
                }
            } else if (!(d > 0.0) || isinf(d)) {
                if (lane == 0) atomicCAS(info, 0, base + c0 + j + 1);
                d = 1.0;
            }
            const double ri = rsqrt(d);
            const double r = d * ri;
            a[j] = (l == j) ? r : a[j] * (LDL ? ri * sj : ri);
            if (lane == 0) { Ri[c0 + j] = ri; if (LDL) Sg[c0 + j] = sj; }
            const double aj = LDL ? sj * a[j] : a[j];
            double *cb = Sc + (j & 1) * 16;
            if (lane < 16) cb[l] = a[j];
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < 16; ++k) a[k] = fma(-aj, cb[k], a[k]);
        }
        __syncwarp();
        if (lane < 16) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k <= l) As[(c0 + k) * L2_LD + c0 + l] = a[k];
        }
    };
    // (c) one 8-row tile row mt of the trailing update A22 -= X X^T (lower 8 x 8 tiles, four column tiles at a time so that
    //     four independent DMMA chains are in flight); X = columns c0 .. c0+15, trailing block starts at r0
    auto update_tile_row = [&](const int c0, const int r0, const int mt) {
        const int m0 = r0 + 8 * mt;
        double av[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) av[kk] = As[(c0 + 4 * kk + tig) * L2_LD + m0 + gid] * (LDL ? Sg[c0 + 4 * kk + tig] : 1.0);
        for (int nt0 = 0; nt0 <= mt; nt0 += 4) {
            double acc[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                double bv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) bv[q] = As[(c0 + 4 * kk + tig) * L2_LD + r0 + 8 * min(nt0 + q, mt) + gid];
#pragma unroll
                for (int q = 0; q < 4; ++q) dmma_tile(acc[q][0], acc[q][1], av[kk], bv[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (nt0 + q > mt) continue;
                const int n0 = r0 + 8 * (nt0 + q);
                As[(n0 + 2 * tig) * L2_LD + m0 + gid] -= acc[q][0];
                As[(n0 + 2 * tig + 1) * L2_LD + m0 + gid] -= acc[q][1];
            }
        }
    };
    if (warp == 0) factor_diag(0);
    __syncthreads();
    LEAF_CLK(2);
    for (int p = 0; p < 7; ++p) {
        const int c0 = 16 * p, r0 = c0 + 16, R = HD_LEAF - r0;
        // (b) X = A21 L11^-T by (right-looking) forward substitution, one thread per row -- as dtrsm would: no explicit
        //     inverse on the factor itself, so a single-leaf matrix gets a classical, backward-stable Cholesky
        if (tid < R) {
            const int m = r0 + tid;
            double x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = As[(c0 + j) * L2_LD + m];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double xj = LDL ? x[j] * Ri[c0 + j] * Sg[c0 + j] : x[j] * Ri[c0 + j];
                x[j] = xj;
                const double xs = LDL ? xj * Sg[c0 + j] : xj;
#pragma unroll
                for (int k = j + 1; k < 16; ++k) x[k] = fma(-xs, As[(c0 + j) * L2_LD + c0 + k], x[k]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) As[(c0 + j) * L2_LD + m] = x[j];
        }
        __syncthreads();
        LEAF_CLK(3 + 3 * p);
        // (c) + look-ahead: warp 0 updates the next diagonal block (tile rows 0 and 1) and factors it at once -- the 16
        //     sequential pivots of (a) are the longest dependency chain of a panel -- while warps 1..7 update the other tile
        //     rows, two per warp (rows 2 + i and T - 1 - i: T + 2 tiles together, balanced)
        {
            const int T = R / 8;
            if (warp == 0) {
                update_tile_row(c0, r0, 0);
                update_tile_row(c0, r0, 1);
                __syncwarp();
                factor_diag(r0);
            } else {
                const int lo = 2 + (warp - 1), hi = T - 1 - (warp - 1);
                if (lo <= hi) {
                    update_tile_row(c0, r0, lo);
                    if (hi != lo) update_tile_row(c0, r0, hi);
                }
            }
        }
        __syncthreads();
        LEAF_CLK(4 + 3 * p);
    }
    LEAF_CLK(26);
    if (LDL && tid < HD_LEAF) sgn[tid] = Sg[tid];
    // L back to global (lower only)
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += L2_THREADS) {
        const int i = e & 127, j = e >> 7;
        if (i >= j) A[(long) j * lda + i] = As[j * L2_LD + i];
    }
    LEAF_CLK(27);
    // inverses of the 8 diagonal blocks, one per warp: lane c < 16 solves L_rr w = e_c by forward substitution
    {
        const int r = warp, rr0 = 16 * r;
        if (lane < 16) {
            const double *Bd = As + rr0 * L2_LD + rr0;
            double w[16];
#pragma unroll
            for (int l = 0; l < 16; ++l) w[l] = (l == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double wk = w[k] * Ri[rr0 + k];
                w[k] = wk;
#pragma unroll
                for (int l = k + 1; l < 16; ++l) w[l] = fma(-wk, Bd[k * L2_LD + l], w[l]);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) Wd[r * 16 * L2_WLD + lane * L2_WLD + k] = (k >= lane) ? w[k] : 0.0; // W[k][lane]
        }
    }
    __syncthreads();
    LEAF_CLK(28);
    // inverse: warp c builds block column c (blocks r = c+1 .. 7) on its own
    {
        const int c = warp, cc0 = 16 * c;
        double *S = Sc + warp * 16 * L2_WLD;
        for (int r = c + 1; r < 8; ++r) {
            const int rr0 = 16 * r;
            double acc[2][2][2], acb[2][2][2]; // two accumulator sets (even / odd k-steps): 8 independent DMMA chains
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acb[i][j][0] = acb[i][j][1] = 0.0;
            for (int k = c; k < r; ++k) {
                const int kk0 = 16 * k;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) av[i] = As[(kk0 + 4 * kk + tig) * L2_LD + rr0 + 8 * i + gid]; // L[rr0+m][kk0+k']
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        bv[j] = (k == c) ? Wd[c * 16 * L2_WLD + (8 * j + gid) * L2_WLD + 4 * kk + tig]       // W_c[k'][n]
                                         : As[(kk0 + 4 * kk + tig) * L2_LD + cc0 + 8 * j + gid];             // X[kk0+k'][cc0+n]
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (kk & 1) dmma_tile(acb[i][j][0], acb[i][j][1], av[i], bv[j]);
                            else dmma_tile(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
                        }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) { acc[i][j][0] += acb[i][j][0]; acc[i][j][1] += acb[i][j][1]; }
            // S (m, n) -> scratch S[n * WLD + m]
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    S[(8 * j + 2 * tig) * L2_WLD + 8 * i + gid] = acc[i][j][0];
                    S[(8 * j + 2 * tig + 1) * L2_WLD + 8 * i + gid] = acc[i][j][1];
                }
            __syncwarp();
            double out[2][2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) out[i][j][0] = out[i][j][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                double av[2], bv[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) av[i] = Wd[r * 16 * L2_WLD + (4 * kk + tig) * L2_WLD + 8 * i + gid]; // W_r[m][k']
#pragma unroll
                for (int j = 0; j < 2; ++j) bv[j] = S[(8 * j + gid) * L2_WLD + 4 * kk + tig];                     // S[k'][n]
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) dmma_tile(out[i][j][0], out[i][j][1], av[i], bv[j]);
            }
            __syncwarp();
            // X[rr0+m][cc0+n] = -out, stored transposed: As[(rr0+m) * LD + cc0+n]
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    As[(rr0 + 8 * i + gid) * L2_LD + cc0 + 8 * j + 2 * tig] = -out[i][j][0];
                    As[(rr0 + 8 * i + gid) * L2_LD + cc0 + 8 * j + 2 * tig + 1] = -out[i][j][1];
                }
            __syncwarp();
        }
    }
    __syncthreads();
    LEAF_CLK(29);
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += L2_THREADS) {
        const int i = e & 127, j = e >> 7;
        double v = 0.0;
        if (i >= j) v = ((i >> 4) == (j >> 4)) ? Wd[(i >> 4) * 16 * L2_WLD + (j & 15) * L2_WLD + (i & 15)] : As[i * L2_LD + j];
        Dinv[e] = v;
    }
    LEAF_CLK(30);
}

// ------------------------------------------------------------------------------------------------------------
// LDL^T leaf with BOUNDED Bunch-Kaufman pivoting (the reference's indefinite fallback is LAPACK dsytrf,
// linalg/hdsdp_linsolver.c:1662-1825: symmetric pivoting with 1 x 1 and 2 x 2 pivots).  Pivots are searched inside the
// 128 x 128 diagonal leaf only, so the block recursion, the panel GEMMs and the multi-GPU layout stay as they are:
//   P^T A P = L D L^T  (Bunch-Kaufman partial pivoting, alpha = (1 + sqrt 17) / 8, D block diagonal),
//   D = Q Lambda Q^T per block  =>  A = G J G^T,  G = P L Q |Lambda|^(1/2),  J = sign(Lambda).
// G is not triangular, but nothing outside the leaf ever reads a diagonal leaf of the factor: panels are solved with
// the explicit G^-1 (trsm_rec), the triangular solves multiply by G^-1 / G^-T (trsv.cu), the updates use J.  The
// kernel writes Dinv = |Lambda|^(-1/2) Q^T L^-1 P^T, the signs and (for checks only) G into the full leaf.  |lambda| <= *floorp
// is replaced by +floorp and counted (static pivoting stays as the backstop for a leaf that is singular as a whole: pivots
// are never taken from another leaf, so this is weaker than dsytrf -- the residual gate of the KKT solve covers the rest).
// One CTA, the full symmetric leaf in shared memory; a fallback path, written for clarity: 0.35 ms per leaf against 45 us for
// the Cholesky leaf (tools/probe_ldl.py: LDL^T of n = 8192 32.8 ms, unpivoted 11.7 ms; n = 20 000 127 ms vs 101 ms).
constexpr int BK_LD = HD_LEAF + 1, BK_THREADS = 256;
constexpr int BK_SMEM = (HD_LEAF * BK_LD + 4 * HD_LEAF) * 8;

__device__ __forceinline__ void bk_argmax(double v, int idx, double *rv, int *ri, double &outv, int &outi) {
    // largest v (ties: smallest idx) over the CTA; v < 0 / NaN never wins.  Ends with the scratch free for reuse.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!(v >= 0.0)) { v = -1.0; idx = 1 << 30; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { rv[warp] = v; ri[warp] = idx; }
    __syncthreads();
    outv = rv[0]; outi = ri[0];
#pragma unroll
    for (int w = 1; w < BK_THREADS / 32; ++w)
        if (rv[w] > outv || (rv[w] == outv && ri[w] < outi)) { outv = rv[w]; outi = ri[w]; }
    __syncthreads();
}

__global__ void __launch_bounds__(BK_THREADS, 1)
ldl_bk_leaf_kernel(double *A, long lda, double *Dinv, double *sgn, const double *floorp, int *nperturb) {
    extern __shared__ __align__(16) double sm[];
    double *As = sm;                       // 128 x 129, column-major, FULL symmetric trailing matrix; L below the diagonal of finished columns
    double *cu = sm + HD_LEAF * BK_LD;     // pivot column(s) and multipliers of the current step
    double *cv = cu + HD_LEAF, *lu = cv + HD_LEAF, *lv = lu + HD_LEAF;
    __shared__ int perm[HD_LEAF], bt[HD_LEAF];           // bt: 1 = 1 x 1 pivot, 2 / 0 = first / second column of a 2 x 2 pivot
    __shared__ double lam[HD_LEAF], rcs[HD_LEAF], rsn[HD_LEAF];
    __shared__ double rv[BK_THREADS / 32];
    __shared__ int ri[BK_THREADS / 32];
    const int tid = threadIdx.x, row = tid & 127, half = tid >> 7;
    const double alpha = 0.6403882032022076, fl = *floorp;
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += BK_THREADS) {
        const int i = e & 127, j = e >> 7;
        As[j * BK_LD + i] = (i >= j) ? A[(long) j * lda + i] : A[(long) i * lda + j];
    }
    if (tid < HD_LEAF) perm[tid] = tid;
    __syncthreads();
    // symmetric interchange of rows / columns p and q (whole rows: also the finished columns of L)
    auto swap_rc = [&](const int p, const int q) {
        if (tid < HD_LEAF) { const double t = As[tid * BK_LD + p]; As[tid * BK_LD + p] = As[tid * BK_LD + q]; As[tid * BK_LD + q] = t; }
        __syncthreads();
        if (tid < HD_LEAF) { const double t = As[p * BK_LD + tid]; As[p * BK_LD + tid] = As[q * BK_LD + tid]; As[q * BK_LD + tid] = t; }
        if (tid == 0) { const int t = perm[p]; perm[p] = perm[q]; perm[q] = t; }
        __syncthreads();
    };
    int k = 0;
    while (k < HD_LEAF) {
        double lmax, sigma; int r, dummy;
        bk_argmax((tid < HD_LEAF && tid > k) ? fabs(As[k * BK_LD + tid]) : -1.0, tid, rv, ri, lmax, r);
        const double akk = fabs(As[k * BK_LD + k]);
        int kind = 1;
        if (lmax > 0.0 && !(akk >= alpha * lmax)) { // uniform over the CTA: every thread sees the same shared values
            bk_argmax((tid < HD_LEAF && tid >= k && tid != r) ? fabs(As[r * BK_LD + tid]) : -1.0, tid, rv, ri, sigma, dummy);
            const double arr = fabs(As[r * BK_LD + r]);
            if (akk * sigma >= alpha * lmax * lmax) { }
            else if (arr >= alpha * sigma) swap_rc(k, r);
            else { kind = 2; if (r != k + 1) swap_rc(k + 1, r); }
        }
        if (kind == 1) {
            double d = As[k * BK_LD + k];
            if (!(fabs(d) > fl) || isinf(d)) { d = fl; if (tid == 0) atomicAdd(nperturb, 1); }
            if (tid < HD_LEAF) { const double c = tid > k ? As[k * BK_LD + tid] : 0.0; cu[tid] = c; lu[tid] = c / d; }
            __syncthreads();
            if (row > k)
                for (int j = k + 1 + half; j < HD_LEAF; j += 2) // (row, j) and (j, row) get the identical product: exact symmetry
                    As[j * BK_LD + row] -= (row >= j) ? lu[row] * cu[j] : lu[j] * cu[row];
            if (tid < HD_LEAF && tid > k) As[k * BK_LD + tid] = lu[tid];
            if (tid == 0) { bt[k] = 1; lam[k] = d; rcs[k] = 1.0; rsn[k] = 0.0; }
            k += 1;
        } else {
            const double a = As[k * BK_LD + k], b = As[k * BK_LD + k + 1], c = As[(k + 1) * BK_LD + k + 1];
            double cs = 1.0, sn = 0.0, l1 = a, l2 = c; // Jacobi rotation: Q^T [a b; b c] Q = diag(l1, l2), Q = [cs sn; -sn cs]
            if (b != 0.0) {
                const double tau = (c - a) / (2.0 * b);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                cs = rsqrt(1.0 + t * t); sn = t * cs; l1 = a - t * b; l2 = c + t * b;
            }
            if (!(fabs(l1) > fl) || isinf(l1)) { l1 = fl; if (tid == 0) atomicAdd(nperturb, 1); }
            if (!(fabs(l2) > fl) || isinf(l2)) { l2 = fl; if (tid == 0) atomicAdd(nperturb, 1); }
            if (tid < HD_LEAF) {
                const double u = tid > k + 1 ? As[k * BK_LD + tid] : 0.0, v = tid > k + 1 ? As[(k + 1) * BK_LD + tid] : 0.0;
                const double q1 = (cs * u - sn * v) / l1, q2 = (sn * u + cs * v) / l2; // [u v] Q Lambda^-1
                cu[tid] = u; cv[tid] = v;
                lu[tid] = q1 * cs + q2 * sn; lv[tid] = -q1 * sn + q2 * cs;             // ... Q^T = [u v] D^-1
            }
            __syncthreads();
            if (row > k + 1)
                for (int j = k + 2 + half; j < HD_LEAF; j += 2)
                    As[j * BK_LD + row] -= (row >= j) ? lu[row] * cu[j] + lv[row] * cv[j] : lu[j] * cu[row] + lv[j] * cv[row];
            if (tid < HD_LEAF && tid > k + 1) { As[k * BK_LD + tid] = lu[tid]; As[(k + 1) * BK_LD + tid] = lv[tid]; }
            if (tid == 0) {
                As[k * BK_LD + k + 1] = 0.0; // L is unit lower with an identity 2 x 2 diagonal block
                bt[k] = 2; bt[k + 1] = 0; lam[k] = l1; lam[k + 1] = l2;
                rcs[k] = rcs[k + 1] = cs; rsn[k] = rsn[k + 1] = sn;
            }
            k += 2;
        }
        __syncthreads();
    }
    // G = P L Q |Lambda|^(1/2) back to the leaf of the factor (all 128 x 128 entries: G is not triangular once a pivot moved).
    // Nothing on the solver path reads it; it is there so that A = G J G^T can be checked from outside.
    {
        auto Lf = [&](const int ii, const int cc) { return ii > cc ? As[cc * BK_LD + ii] : (ii == cc ? 1.0 : 0.0); };
        for (int e = tid; e < HD_LEAF * HD_LEAF; e += BK_THREADS) {
            const int ii = e & 127, cc = e >> 7, b = bt[cc];
            double g;
            if (b == 1) g = Lf(ii, cc);
            else if (b == 2) g = rcs[cc] * Lf(ii, cc) - rsn[cc] * Lf(ii, cc + 1);
            else g = rsn[cc] * Lf(ii, cc - 1) + rcs[cc] * Lf(ii, cc);
            A[(long) cc * lda + perm[ii]] = g * sqrt(fabs(lam[cc]));
        }
    }
    __syncthreads();
    // Y = L^-1 in place (unit lower triangular; the diagonal is implicit): columns from the right,
    // Y[j+1:, j] = -Y[j+1:, j+1:] L[j+1:, j]
    for (int j = HD_LEAF - 2; j >= 0; --j) {
        if (tid < HD_LEAF) cu[tid] = tid > j ? As[j * BK_LD + tid] : 0.0;
        __syncthreads();
        if (tid < HD_LEAF && tid > j) {
            double s = cu[tid];
            for (int t = j + 1; t < tid; ++t) s = fma(As[t * BK_LD + tid], cu[t], s);
            As[j * BK_LD + tid] = -s;
        }
        __syncthreads();
    }
    // Dinv = |Lambda|^(-1/2) Q^T Y P^T : column perm[i] of Dinv is column i of Z = |Lambda|^(-1/2) Q^T Y
    auto Y = [&](const int rr, const int ii) { return ii < rr ? As[ii * BK_LD + rr] : (ii == rr ? 1.0 : 0.0); };
    for (int e = tid; e < HD_LEAF * HD_LEAF; e += BK_THREADS) {
        const int rr = e & 127, ii = e >> 7, b = bt[rr];
        const double sc = rsqrt(fabs(lam[rr]));
        double z;
        if (b == 1) z = Y(rr, ii);
        else if (b == 2) z = rcs[rr] * Y(rr, ii) - rsn[rr] * Y(rr + 1, ii);
        else z = rsn[rr] * Y(rr - 1, ii) + rcs[rr] * Y(rr, ii);
        Dinv[(long) perm[ii] * HD_LEAF + rr] = z * sc;
    }
    if (tid < HD_LEAF) sgn[tid] = lam[tid] < 0.0 ? -1.0 : 1.0;
}

// X leaf (upper triangular) = Dinv^T
__global__ void leaf_transpose_kernel(const double *__restrict__ Dinv, double *X, long ldx) {
    __shared__ double t[32][33];
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y; // 32 x 8
    for (int r = ty; r < 32; r += 8) t[r][tx] = Dinv[(long) (by + r) * HD_LEAF + bx + tx]; // t[k][i] = Dinv[i,k]
    __syncthreads();
    for (int r = ty; r < 32; r += 8) X[(long) (bx + r) * ldx + by + tx] = t[tx][r]; // X[k, i] = Dinv[i, k]
}

// transposed copies of every inverse leaf (coalesced access for the L^T solve), one CTA per leaf
__global__ void leaf_transpose_all_kernel(const double *__restrict__ Dinv, double *DinvT) {
    __shared__ double t[32][33];
    const double *src = Dinv + (long) blockIdx.z * HD_LEAF * HD_LEAF;
    double *dst = DinvT + (long) blockIdx.z * HD_LEAF * HD_LEAF;
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8) t[r][tx] = src[(long) (by + r) * HD_LEAF + bx + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) dst[(long) (bx + r) * HD_LEAF + by + tx] = t[tx][r];
}

__global__ void logdet_kernel(const double *L, long ld, int n, double *out, double *diag) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        double d = L[(long) i * ld + i];
        if (diag) diag[i] = d;
        s += log(d);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && out) *out = 2.0 * red[0];
}

const int LEAF_SMEM = (HD_LEAF * HD_LEAF + HD_LEAF) * 8;

// LDL^T mode of the recursion (reference fallback dsytrf, linalg/hdsdp_linsolver.c:1662-1825): set by chol_factor around
// the enqueue; the sign vector of a sub-block is found from its inverse-leaf pointer (leaf index = column / 128)
struct LdlCtx { double *sgn; const double *dinv_base; const double *floorp; int *nperturb; };
const LdlCtx *g_ldl = nullptr;
inline double *sgn_of(const double *dinv) {
    return g_ldl ? g_ldl->sgn + ((dinv - g_ldl->dinv_base) / (HD_LEAF * HD_LEAF)) * HD_LEAF : nullptr;
}

const double *g_peer_local = nullptr;
double *g_peer_dst = nullptr;

int split_leaves(int n) { return ((n / HD_LEAF) / 2) * HD_LEAF; } // n1 (multiple of 128, >= 128 when n >= 256)

// B (rows x n, ld ldb) <- B * L^-T,  L n x n lower (ld ldl), dinv = inverse leaves of L
int trsm_rec(cudaStream_t st, double *B, long ldb, int rows, const double *L, long ldl, int n, const double *dinv) {
    if (rows <= 0) return HD_OK;
    if (n == HD_LEAF) {
        GemmArgs g{};
        g.M = rows; g.N = HD_LEAF; g.K = HD_LEAF;
        g.A = B; g.lda = ldb; g.B = dinv; g.ldb = HD_LEAF; g.C = B; g.ldc = ldb;
        g.alpha = 1.0; g.beta = 0.0; g.flags = 0;
        if (g_ldl) { g.flags = HD_GEMM_EPI_COLSCALE; g.sb = sgn_of(dinv); } // X = (B L^-T) J
        if (g_peer_dst) g.peerC = g_peer_dst + (B - g_peer_local);
        return hd_gemm_nt(st, g); // in place: each CTA reads exactly the rows it later writes
    }
    int n1 = split_leaves(n), n2 = n - n1;
    HD_CALL(trsm_rec(st, B, ldb, rows, L, ldl, n1, dinv));
    GemmArgs g{};
    g.M = rows; g.N = n2; g.K = n1;
    g.A = B; g.lda = ldb; g.B = L + n1; g.ldb = ldl; g.C = B + (long) n1 * ldb; g.ldc = ldb;
    g.alpha = -1.0; g.beta = 1.0; g.flags = 0; g.ksign = sgn_of(dinv);
    HD_CALL(hd_gemm_nt(st, g));
    return trsm_rec(st, B + (long) n1 * ldb, ldb, rows, L + (long) n1 * ldl + n1, ldl, n2,
                    dinv + (long) (n1 / HD_LEAF) * HD_LEAF * HD_LEAF);
}

int g_leaf_version = 2;
int g_ldl_pivot = 1; // 1: bounded Bunch-Kaufman inside the leaf (default); 0: unpivoted L J L^T with static pivoting only
int potrf_rec(cudaStream_t st, double *A, long lda, int n, double *dinv, int *info, int base) {
    if (n == HD_LEAF) {
        if (g_ldl && g_ldl_pivot) { ++g_hd_launches; ldl_bk_leaf_kernel<<<1, BK_THREADS, BK_SMEM, st>>>(A, lda, dinv, sgn_of(dinv), g_ldl->floorp, g_ldl->nperturb); }
        else if (g_ldl) { ++g_hd_launches; potf2_leaf2_kernel<true><<<1, L2_THREADS, L2_SMEM, st>>>(A, lda, dinv, info, base, sgn_of(dinv), g_ldl->floorp, g_ldl->nperturb); }
        else if (g_leaf_version == 2) { ++g_hd_launches; potf2_leaf2_kernel<false><<<1, L2_THREADS, L2_SMEM, st>>>(A, lda, dinv, info, base, nullptr, nullptr, nullptr); }
        else HDK(potf2_leaf_kernel)<<<1, LEAF_THREADS, LEAF_SMEM, st>>>(A, lda, dinv, info, base);
        HD_CUDA(cudaGetLastError());
        return HD_OK;
    }
    int n1 = split_leaves(n), n2 = n - n1;
    HD_CALL(potrf_rec(st, A, lda, n1, dinv, info, base));
    HD_CALL(trsm_rec(st, A + n1, lda, n2, A, lda, n1, dinv));
    GemmArgs g{};
    g.M = n2; g.N = n2; g.K = n1;
    g.A = A + n1; g.lda = lda; g.B = A + n1; g.ldb = lda; g.C = A + (long) n1 * lda + n1; g.ldc = lda;
    g.alpha = -1.0; g.beta = 1.0; g.flags = HD_GEMM_LOWER; g.ksign = sgn_of(dinv);
    HD_CALL(hd_gemm_nt(st, g));
    return potrf_rec(st, A + (long) n1 * lda + n1, lda, n2, dinv + (long) (n1 / HD_LEAF) * HD_LEAF * HD_LEAF, info,
                     base + n1);
}

// X (n x n upper triangular, ld ldx) = L^-T.   Block formula:
//   X11 = L11^-T, X22 = L22^-T, X12 = -(X11 L21^T) L22^-T  (the right factor applied by trsm_rec)
// The two diagonal sub-blocks of the recursion are independent.  Below INV_FORK_N their products are a handful of CTAs each
// (latency-bound), so the second one is enqueued on another stream of a small pool and joined before the off-diagonal
// block: up to 8 branches of the tree run side by side (forks nest).  Events are reused round-robin: a wait captures the
// event's state when it is enqueued, and every record is followed by its wait before the event comes round again (a node
// of 2048 has ~30 events pending at most).
constexpr int INV_POOL = 8, INV_EVENTS = 256;
int g_inv_fork_n = 2048; // "invert_fork": largest node that forks (0 = off)
cudaStream_t g_inv_pool[INV_POOL] = {nullptr};
cudaEvent_t g_inv_ev[INV_EVENTS] = {nullptr};
int g_inv_next_stream = 0, g_inv_next_event = 0;
cudaEvent_t inv_event() {
    cudaEvent_t &e = g_inv_ev[g_inv_next_event++ % INV_EVENTS];
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}
cudaStream_t inv_stream() {
    cudaStream_t &s = g_inv_pool[g_inv_next_stream++ % INV_POOL];
    if (!s) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    return s;
}

int invert_rec(cudaStream_t st, const double *L, long ldl, int n, const double *dinv, double *X, long ldx) {
    if (n == HD_LEAF) {
        HDK(leaf_transpose_kernel)<<<dim3(4, 4), dim3(32, 8), 0, st>>>(dinv, X, ldx);
        HD_CUDA(cudaGetLastError());
        return HD_OK;
    }
    int n1 = split_leaves(n), n2 = n - n1;
    const double *dinv2 = dinv + (long) (n1 / HD_LEAF) * HD_LEAF * HD_LEAF;
    cudaEvent_t joined = nullptr;
    if (n <= g_inv_fork_n && n2 >= 2 * HD_LEAF) {
        cudaStream_t s2 = inv_stream();
        cudaEvent_t fork = inv_event(), done = inv_event();
        HD_CUDA(cudaEventRecord(fork, st));
        HD_CUDA(cudaStreamWaitEvent(s2, fork, 0));
        HD_CALL(invert_rec(s2, L + (long) n1 * ldl + n1, ldl, n2, dinv2, X + (long) n1 * ldx + n1, ldx));
        HD_CUDA(cudaEventRecord(done, s2));
        HD_CALL(invert_rec(st, L, ldl, n1, dinv, X, ldx));
        joined = done; // X22 is not read by the two products below (they use L22 and its inverse leaves): join after them
    } else {
        HD_CALL(invert_rec(st, L, ldl, n1, dinv, X, ldx));
        HD_CALL(invert_rec(st, L + (long) n1 * ldl + n1, ldl, n2, dinv2, X + (long) n1 * ldx + n1, ldx));
    }
    GemmArgs g{};
    g.M = n1; g.N = n2; g.K = n1;
    g.A = X; g.lda = ldx; g.B = L + n1; g.ldb = ldl; g.C = X + (long) n1 * ldx; g.ldc = ldx;
    g.alpha = -1.0; g.beta = 0.0; g.flags = n1 >= 1024 ? HD_GEMM_KTRI_A : 0; // X11 is upper triangular: half the k-range on average
    HD_CALL(hd_gemm_nt(st, g));
    HD_CALL(trsm_rec(st, X + (long) n1 * ldx, ldx, n1, L + (long) n1 * ldl + n1, ldl, n2, dinv2));
    if (joined) HD_CUDA(cudaStreamWaitEvent(st, joined, 0));
    return HD_OK;
}

} // namespace

static int ensure_leaf_attr() {
    static unsigned long long attr = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr >> (dev & 63) & 1ull)) {
        HD_CUDA(cudaFuncSetAttribute(potf2_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
        HD_CUDA(cudaFuncSetAttribute(potf2_leaf2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM));
        HD_CUDA(cudaFuncSetAttribute(potf2_leaf2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM));
        HD_CUDA(cudaFuncSetAttribute(ldl_bk_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
        attr |= 1ull << (dev & 63);
    }
    return HD_OK;
}
int hd_potrf_rec(cudaStream_t st, double *A, long lda, int n, double *dinv, int *info, int base) {
    HD_CALL(ensure_leaf_attr());
    return potrf_rec(st, A, lda, n, dinv, info, base);
}
int hd_trsm_rec(cudaStream_t st, double *B, long ldb, int rows, const double *L, long ldl, int n, const double *dinv) {
    return trsm_rec(st, B, ldb, rows, L, ldl, n, dinv);
}
void hd_trsm_set_peer(const double *local, double *peer) { g_peer_local = local; g_peer_dst = peer; }
// LDL^T mode for the enqueue calls that follow (dist.cu drives hd_potrf_rec / hd_trsm_rec itself); c == nullptr ends it
void hd_ldl_scope(DenseChol *c) {
    static LdlCtx ctx;
    if (!c) { g_ldl = nullptr; return; }
    ctx.sgn = c->sgn; ctx.dinv_base = c->Dinv; ctx.floorp = c->dfloor; ctx.nperturb = c->dperturb;
    g_ldl = &ctx;
}
int hd_chol_finish(cudaStream_t st, DenseChol *c) {
    HDK(leaf_transpose_all_kernel)<<<dim3(4, 4, c->np / HD_LEAF), dim3(32, 8), 0, st>>>(c->Dinv, c->DinvT);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int chol_create(DenseChol **pc, int n) {
    if (n <= 0) return HD_FAILED;
    DenseChol *c = (DenseChol *) calloc(1, sizeof(DenseChol));
    if (!c) return HD_MEMORY;
    c->n = n;
    c->np = hd_pad(n);
    size_t bytes = (size_t) c->np * c->np * sizeof(double);
    if (cudaMalloc(&c->L, bytes) != cudaSuccess) { free(c); cudaGetLastError(); return HD_MEMORY; }
    // inverse leaves, followed by the np sign entries of the LDL^T mode (one allocation: one IPC handle covers both)
    if (cudaMalloc(&c->Dinv, ((size_t) c->np * HD_LEAF + c->np) * sizeof(double)) != cudaSuccess) {
        cudaFree(c->L); free(c); cudaGetLastError(); return HD_MEMORY;
    }
    if (cudaMalloc(&c->DinvT, (size_t) c->np * HD_LEAF * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&c->sync, sizeof(int) * (2 * (c->np / HD_LEAF) + 4)) != cudaSuccess) {
        cudaFree(c->L); cudaFree(c->Dinv); free(c); cudaGetLastError(); return HD_MEMORY;
    }
    if (cudaMalloc(&c->dinfo, sizeof(int)) != cudaSuccess || cudaMallocHost(&c->hinfo, sizeof(int)) != cudaSuccess) {
        cudaFree(c->L); cudaFree(c->Dinv); free(c); cudaGetLastError(); return HD_MEMORY;
    }
    c->sgn = c->Dinv + (size_t) c->np * HD_LEAF;
    c->work = nullptr;
    c->factored = false;
    *pc = c;
    return HD_OK;
}

void chol_destroy(DenseChol *c) {
    if (!c) return;
    cudaFree(c->L);
    cudaFree(c->Dinv);
    cudaFree(c->DinvT);
    cudaFree(c->sync);
    cudaFree(c->dinfo);
    cudaFreeHost(c->hinfo);
    if (c->work) cudaFree(c->work);
    if (c->dfloor) { cudaFree(c->dfloor); cudaFree(c->dperturb); }
    if (c->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t) c->graph_exec);
    free(c);
}

int chol_ensure_work(DenseChol *c) {
    if (c->work) return HD_OK;
    HD_CUDA(cudaMalloc(&c->work, (size_t) c->np * c->np * sizeof(double)));
    return HD_OK;
}

int chol_load_dev(cudaStream_t st, DenseChol *c, const double *dS, long lds) {
    HD_CALL(hd_copy2d(st, c->L, c->np, dS, lds, c->n, c->n));
    HD_CALL(hd_pad_identity(st, c->L, c->np, c->n, c->np));
    c->factored = false;
    return HD_OK;
}

// Blocked right-looking Cholesky with one step of look-ahead (used for large matrices, i.e. the Schur matrix M).
// The O(n^3) trailing update of step k (one lower-triangular DMMA GEMM with K = NB) runs on the main stream while a
// second, high-priority stream factors the next diagonal block and solves the next panel (the recursive kernels
// above: latency-bound leaves and thin GEMMs), so the low-efficiency panel work is hidden behind the big GEMM.
//   main : [col k+1 -= P_k P_k^T] -> e_col ...... [trailing cols k+2.. -= P_k P_k^T (LOWER)] -> wait e_panel
//   side :            wait e_col -> potrf(A_{k+1,k+1}) -> A_{k+2..,k+1} <- A L^-T -> e_panel
static cudaStream_t g_side = nullptr;
static cudaEvent_t g_ev_col = nullptr, g_ev_panel = nullptr, g_ev_trail = nullptr;
// -1: by size (measured on B200, tools/trace_potrf.py); 1: one-step look-ahead with whole panels on the side stream;
// 2: strip chain + two-step look-ahead (potrf_blocked2); 3: as 2 with the two look-ahead block columns updated by one GEMM
static int g_sched = -1;
void hd_chol_set_sched(int v) { g_sched = v; }
static int g_lookahead_nb = -1; // block size; 0 disables the blocked path; -1 = by size (measured on B200, tools/probe_block.py)

void hd_chol_set_block(int nb) { g_lookahead_nb = nb; }
void hd_chol_set_leaf(int v) { g_leaf_version = v; }
void hd_chol_set_ldl_pivot(int v) { g_ldl_pivot = v != 0; }
void hd_chol_set_invert_fork(int v) { g_inv_fork_n = v; }
int hd_leaf_clocks(long long *out) { return cudaMemcpyFromSymbol(out, g_leaf_clk, sizeof(long long) * 40) == cudaSuccess ? HD_OK : HD_FAILED; }

static int potrf_blocked2(cudaStream_t st, double *A, long lda, int np, double *dinv, int *info, int NB, int base0);
static int potrf_blocked(cudaStream_t st, double *A, long lda, int np, double *dinv, int *info, int NB, int base0, bool tails);
static int auto_block(int np) { return np <= 7168 ? 128 : (np < 14336 ? 256 : (np < 24000 ? 512 : (np < 40000 ? 1024 : 2048))); }
static int g_tail = 1; // "chol_tail": hand the trailing matrix over to the schedule of its own size (see potrf_blocked)
static int g_tail_pct = 100; // probe knob: values > 1 scale the hand-over thresholds (percent)
void hd_chol_set_tail(int v) { g_tail = v != 0; g_tail_pct = v > 1 ? v : 100; }

// Factor a trailing sub-matrix (fully updated, nothing of it factored yet) with the block size and schedule a matrix of that size
// would get on its own.
static int potrf_tail(cudaStream_t st, double *A, long lda, int np, double *dinv, int *info, int base0) {
    const int nb = auto_block(np);
    const int sched = g_sched > 0 ? g_sched : (np <= 10240 ? 3 : 1);
    if (np >= 4 * nb && !g_ldl && sched >= 2) return potrf_blocked2(st, A, lda, np, dinv, info, nb, base0);
    if (np >= 4 * nb) return potrf_blocked(st, A, lda, np, dinv, info, nb, base0, true);
    return potrf_rec(st, A, lda, np, dinv, info, base0);
}

static int potrf_blocked(cudaStream_t st, double *A, long lda, int np, double *dinv, int *info, int NB, int base0, bool tails) {
    if (!g_side) {
        int lo = 0, hi = 0;
        HD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HD_CUDA(cudaStreamCreateWithPriority(&g_side, cudaStreamNonBlocking, hi));
        HD_CUDA(cudaEventCreateWithFlags(&g_ev_col, cudaEventDisableTiming));
        HD_CUDA(cudaEventCreateWithFlags(&g_ev_panel, cudaEventDisableTiming));
    }
    const int nblk = (np + NB - 1) / NB;
    const bool trace = getenv("HDSDPCU_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&]() { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    auto start = [&](int k) { return k * NB; };
    auto size = [&](int k) { return (k == nblk - 1) ? np - k * NB : NB; };
    auto leaves = [&](int k) { return dinv + (long) (start(k) / HD_LEAF) * HD_LEAF * HD_LEAF; };
    // panel 0 on the main stream
    HD_CALL(potrf_rec(st, A, lda, size(0), leaves(0), info, base0));
    mark();
    if (nblk > 1) HD_CALL(trsm_rec(st, A + size(0), lda, np - size(0), A, lda, size(0), leaves(0)));
    mark();
    // Below ~24k the column update of block k+1 is two latency-bound launches: it runs on the side stream (chained in
    // front of the panel work it feeds) so that the main stream only ever issues the large trailing GEMMs.  For large
    // matrices it is real GEMM work that fills the GPU and stays on the main stream.
    const bool col_on_side = np < 24000;
    if (!g_ev_trail) HD_CUDA(cudaEventCreateWithFlags(&g_ev_trail, cudaEventDisableTiming));
    HD_CUDA(cudaEventRecord(g_ev_trail, st));
    for (int k = 0; k + 1 < nblk; ++k) {
        const int s0 = start(k), b0 = size(k), s1 = start(k + 1), b1 = size(k + 1);
        const double *P = A + s0 * lda; // panel k: rows s1.. are the solved block column
        // Tail hand-over.  Late in the factorisation the panel chain (leaves and thin products of a NB-wide diagonal block, each
        // waiting for a CTA slot behind trailing-update CTAs that run for NB / 7 us) outlasts the trailing update: at m = 50 000,
        // NB = 2048 the main stream waited 2.5 ms per step over the last nine steps (HDSDPCU_TRACE).  Once the rest is small
        // enough, panel k is applied to ALL of it and the rest is factored with the block size / schedule of its own size.
        // Measured (tools/trace_potrf.py, hand-over off / on): n = 50 000 1297.8 / 1290.6 ms, 30 000 297.3 / 294.7, 20 000 95.2 / 93.8.
        const int rem = np - s1;
        const int tail_n = (int) ((long) (NB >= 2048 ? 12288 : (NB >= 1024 ? 8704 : (NB >= 512 ? 5632 : (NB >= 256 ? 3584 : 0)))) * g_tail_pct / 100);
        if (tails && !trace && rem <= tail_n && rem >= 8 * HD_LEAF) {
            GemmArgs g{};
            g.M = rem; g.N = rem; g.K = b0;
            g.A = P + s1; g.lda = lda; g.B = P + s1; g.ldb = lda; g.C = A + (long) s1 * lda + s1; g.ldc = lda;
            g.alpha = -1.0; g.beta = 1.0; g.flags = HD_GEMM_LOWER; g.ksign = sgn_of(leaves(k));
            HD_CALL(hd_gemm_nt(st, g));
            return potrf_tail(st, A + (long) s1 * lda + s1, lda, rem, leaves(k + 1), info, base0 + s1);
        }
        cudaStream_t cst = col_on_side ? g_side : st;
        if (col_on_side) HD_CUDA(cudaStreamWaitEvent(g_side, g_ev_trail, 0)); // trailing update k-1 (and panel 0) done
        // (1) update block column k+1: diagonal block (lower tiles) and the rectangle below it
        GemmArgs g{};
        g.M = b1; g.N = b1; g.K = b0;
        g.A = P + s1; g.lda = lda; g.B = P + s1; g.ldb = lda; g.C = A + (long) s1 * lda + s1; g.ldc = lda;
        g.alpha = -1.0; g.beta = 1.0; g.flags = HD_GEMM_LOWER; g.ksign = sgn_of(leaves(k));
        HD_CALL(hd_gemm_nt(cst, g));
        const int below = np - (s1 + b1);
        if (below > 0) {
            g.M = below; g.N = b1; g.flags = 0;
            g.A = P + s1 + b1; g.C = A + (long) s1 * lda + s1 + b1;
            HD_CALL(hd_gemm_nt(cst, g));
        }
        if (!col_on_side) {
            HD_CUDA(cudaEventRecord(g_ev_col, st));
            HD_CUDA(cudaStreamWaitEvent(g_side, g_ev_col, 0));
        }
        mark();
        // (2) side stream: factor block k+1 and solve its panel
        HD_CALL(potrf_rec(g_side, A + (long) s1 * lda + s1, lda, b1, leaves(k + 1), info, base0 + s1));
        if (below > 0) HD_CALL(trsm_rec(g_side, A + (long) s1 * lda + s1 + b1, lda, below, A + (long) s1 * lda + s1, lda, b1, leaves(k + 1)));
        HD_CUDA(cudaEventRecord(g_ev_panel, g_side));
        // (3) main stream: the rest of the trailing matrix (columns of blocks k+2..)
        if (below > 0) {
            g.M = below; g.N = below; g.K = b0;
            g.A = P + s1 + b1; g.B = P + s1 + b1; g.C = A + (long) (s1 + b1) * lda + s1 + b1;
            g.flags = HD_GEMM_LOWER;
            HD_CALL(hd_gemm_nt(st, g));
        }
        HD_CUDA(cudaEventRecord(g_ev_trail, st));
        mark();
        HD_CUDA(cudaStreamWaitEvent(st, g_ev_panel, 0));
        mark();
    }
    if (trace) {
        cudaStreamSynchronize(st);
        float t;
        cudaEventElapsedTime(&t, tev[0], tev[1]);
        fprintf(stderr, "[trace] potrf_blocked np=%d NB=%d panel0 trsm %.3f ms\n", np, NB, t);
        double tc = 0, tg = 0, tw = 0;
        for (int k = 0; k + 1 < nblk; ++k) {
            float a, b, c;
            cudaEventElapsedTime(&a, tev[1 + 3 * k], tev[2 + 3 * k]);
            cudaEventElapsedTime(&b, tev[2 + 3 * k], tev[3 + 3 * k]);
            cudaEventElapsedTime(&c, tev[3 + 3 * k], tev[4 + 3 * k]);
            tc += a; tg += b; tw += c;
            int below = np - (k + 2) * NB; if (below < 0) below = 0;
            double fl = (double) below * below * NB;
            if (k % 4 == 0 || c > 0.5) fprintf(stderr, "[trace] k=%3d col %.3f ms  trailing %.3f ms (%.1f TF)  wait-panel %.3f ms\n", k, a, b, b > 0 ? fl / b / 1e9 : 0.0, c);
        }
        fprintf(stderr, "[trace] totals: col %.1f ms trailing %.1f ms wait %.1f ms\n", tc, tg, tw);
        for (auto e : tev) cudaEventDestroy(e);
    }
    return HD_OK;
}


// Blocked Cholesky, second schedule (Cholesky mode, single GPU): the critical chain is cut down to what it needs.
// Diagonal block k+1 depends on panel k only through the NB rows just below diagonal block k, so the high-priority side stream
// runs  POTRF(k) -> STRIP(k) (solve of those NB rows only) -> DIAGUPD(k+1) (A_{k+1,k+1} -= strip strip^T) -> POTRF(k+1) ...
// with thin-tile products (gemm_nt.cu), while the main stream does the bulk one to two steps behind:
//   TRSMB(k)  rows below the strip of panel k           (needs POTRF(k))
//   UPDCOL(k) block column k+1 below its diagonal block (needs STRIP(k); feeds STRIP(k+1))
//   UPDCOL2(k) block column k+2 from its diagonal block down (feeds DIAGUPD(k+2): two steps of look-ahead)
//   UPDREST(k) everything from block column k+3 on (one lower-triangular GEMM, K = NB)
// The side stream waits for the main stream only when the bulk falls more than one step behind.
static std::vector<cudaEvent_t> g_evs[4]; // potrf, strip (side -> main); col, col2 (main -> side)

// Optional SM partition for the strip schedule (CUDA green contexts; driver API resolved through the runtime so that the
// library has no link-time dependency on libcuda): the chain gets 8 SMs of its own (the smallest partition sm_100 allows),
// the bulk GEMMs the other 140.  The idea: a high-priority stream only wins the NEXT free CTA slot, so a chain kernel waits
// behind bulk CTAs that run for 20 .. 40 us.  MEASURED ON B200 AND REJECTED (tools/trace_potrf.py, ms without / with the
// partition): n = 4096 2.44 / 2.45, 6144 5.28 / 5.38, 8192 9.63 / 9.78, 10240 16.3 / 16.8 -- the 5 % of SMs the bulk loses
// cost more than the slot waits.  Kept behind hdsdpcu_set_option("chol_partition", 1), off by default.
struct SmPartition { bool tried = false; cudaStream_t chain = nullptr, bulk = nullptr; int chain_sms = 0, bulk_sms = 0; };
static SmPartition g_part[64];
static int g_partition = 0;
static cudaEvent_t g_ev_join = nullptr;
void hd_chol_set_partition(int v) { g_partition = v; }

template <typename F> static bool drv_entry(const char *name, F &fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(ptr);
    return true;
}

static SmPartition *sm_partition() {
    int dev = 0;
    cudaGetDevice(&dev);
    SmPartition &P = g_part[dev & 63];
    if (P.tried) return P.chain ? &P : nullptr;
    P.tried = true;
    decltype(&cuDeviceGet) pDeviceGet = nullptr;
    decltype(&cuDeviceGetDevResource) pGetRes = nullptr;
    decltype(&cuDevSmResourceSplitByCount) pSplit = nullptr;
    decltype(&cuDevResourceGenerateDesc) pDesc = nullptr;
    decltype(&cuGreenCtxCreate) pCreate = nullptr;
    decltype(&cuGreenCtxStreamCreate) pStream = nullptr;
    if (!drv_entry("cuDeviceGet", pDeviceGet) || !drv_entry("cuDeviceGetDevResource", pGetRes) ||
        !drv_entry("cuDevSmResourceSplitByCount", pSplit) || !drv_entry("cuDevResourceGenerateDesc", pDesc) ||
        !drv_entry("cuGreenCtxCreate", pCreate) || !drv_entry("cuGreenCtxStreamCreate", pStream)) {
        fprintf(stderr, "[hdsdpcu] green contexts not available in this driver: SM partition off\n");
        return nullptr;
    }
    CUdevice cudev;
    CUdevResource all, grp[1], rest;
    unsigned int ngrp = 1;
    CUdevResourceDesc d_chain = nullptr, d_bulk = nullptr;
    CUgreenCtx c_chain = nullptr, c_bulk = nullptr;
    CUstream s_chain = nullptr, s_bulk = nullptr;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (pDeviceGet(&cudev, dev) != CUDA_SUCCESS || pGetRes(cudev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS ||
        pSplit(grp, &ngrp, &all, &rest, 0, 8) != CUDA_SUCCESS || ngrp < 1 || rest.sm.smCount < 64 ||
        pDesc(&d_chain, &grp[0], 1) != CUDA_SUCCESS || pDesc(&d_bulk, &rest, 1) != CUDA_SUCCESS ||
        pCreate(&c_chain, d_chain, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
        pCreate(&c_bulk, d_bulk, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
        pStream(&s_chain, c_chain, CU_STREAM_NON_BLOCKING, hi) != CUDA_SUCCESS ||
        pStream(&s_bulk, c_bulk, CU_STREAM_NON_BLOCKING, lo) != CUDA_SUCCESS) {
        fprintf(stderr, "[hdsdpcu] SM partition could not be created: off\n");
        cudaGetLastError();
        return nullptr;
    }
    P.chain = (cudaStream_t) s_chain; P.bulk = (cudaStream_t) s_bulk;
    P.chain_sms = (int) grp[0].sm.smCount; P.bulk_sms = (int) rest.sm.smCount;
    if (getenv("HDSDPCU_TRACE")) fprintf(stderr, "[trace] SM partition: chain %d SMs, bulk %d SMs\n", P.chain_sms, P.bulk_sms);
    return &P;
}
int hd_chol_partition_sms(int *chain, int *bulk) {
    SmPartition *P = sm_partition();
    if (chain) *chain = P ? P->chain_sms : 0;
    if (bulk) *bulk = P ? P->bulk_sms : 0;
    return P ? HD_OK : HD_FAILED;
}
static cudaEvent_t g_ev_fork = nullptr;

static int potrf_blocked2(cudaStream_t st, double *A, long lda, int np, double *dinv, int *info, int NB, int base0) {
    if (!g_side) {
        int lo = 0, hi = 0;
        HD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HD_CUDA(cudaStreamCreateWithPriority(&g_side, cudaStreamNonBlocking, hi));
        HD_CUDA(cudaEventCreateWithFlags(&g_ev_col, cudaEventDisableTiming));
        HD_CUDA(cudaEventCreateWithFlags(&g_ev_panel, cudaEventDisableTiming));
    }
    if (!g_ev_fork) HD_CUDA(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
    const int nblk = (np + NB - 1) / NB;
    for (int t = 0; t < 4; ++t)
        while ((int) g_evs[t].size() < nblk) {
            cudaEvent_t e;
            HD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            g_evs[t].push_back(e);
        }
    auto start = [&](int k) { return k < nblk ? k * NB : np; };
    auto size = [&](int k) { return k >= nblk ? 0 : ((k == nblk - 1) ? np - k * NB : NB); };
    auto leaves = [&](int k) { return dinv + (long) (start(k) / HD_LEAF) * HD_LEAF * HD_LEAF; };
    auto at = [&](int r, int c) { return A + (long) c * lda + r; };
    cudaStream_t side = g_side;
    cudaStream_t const caller = st;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    SmPartition *part = (g_partition > 0 && cap == cudaStreamCaptureStatusNone) ? sm_partition() : nullptr;
    HD_CUDA(cudaEventRecord(g_ev_fork, st));
    if (part) {
        side = part->chain;
        st = part->bulk; // from here on "st" is the bulk stream; the caller's stream only forks and joins
        HD_CUDA(cudaStreamWaitEvent(st, g_ev_fork, 0));
    }
    HD_CUDA(cudaStreamWaitEvent(side, g_ev_fork, 0));
    for (int k = 0; k < nblk; ++k) {
        const int s0 = start(k), b0 = size(k), s1 = start(k + 1), b1 = size(k + 1), s2 = start(k + 2), b2 = size(k + 2), s3 = start(k + 3);
        // ---- side: the chain ----
        if (k >= 1) {
            if (k >= 2) HD_CUDA(cudaStreamWaitEvent(side, g_evs[3][k - 2], 0));          // UPDCOL2(k-2) and all earlier bulk updates of A_kk
            GemmArgs g{};
            g.M = b0; g.N = b0; g.K = size(k - 1);
            g.A = at(s0, start(k - 1)); g.lda = lda; g.B = g.A; g.ldb = lda; g.C = at(s0, s0); g.ldc = lda;
            g.alpha = -1.0; g.beta = 1.0; g.flags = HD_GEMM_LOWER;
            HD_CALL(hd_gemm_nt(side, g));                                                  // DIAGUPD(k)
        }
        HD_CALL(potrf_rec(side, at(s0, s0), lda, b0, leaves(k), info, base0 + s0));        // POTRF(k)
        HD_CUDA(cudaEventRecord(g_evs[0][k], side));
        if (b1 > 0) {
            if (k >= 1) HD_CUDA(cudaStreamWaitEvent(side, g_evs[2][k - 1], 0));          // UPDCOL(k-1): column block k below its diagonal
            HD_CALL(trsm_rec(side, at(s1, s0), lda, b1, at(s0, s0), lda, b0, leaves(k)));  // STRIP(k)
            HD_CUDA(cudaEventRecord(g_evs[1][k], side));
        }
        // ---- main: the bulk ----
        const int rowsR = np - s2;
        if (b1 > 0 && rowsR > 0) {
            HD_CUDA(cudaStreamWaitEvent(st, g_evs[0][k], 0));
            HD_CALL(trsm_rec(st, at(s2, s0), lda, rowsR, at(s0, s0), lda, b0, leaves(k))); // TRSMB(k)
            HD_CUDA(cudaStreamWaitEvent(st, g_evs[1][k], 0));
            GemmArgs g{};
            g.K = b0; g.lda = lda; g.ldb = lda; g.ldc = lda; g.alpha = -1.0; g.beta = 1.0;
            const int rows3 = np - s3;
            if (g_sched != 2) {
                // LOOK(k): block columns k+1 and k+2 from row s2 down as ONE rectangle (the strictly upper part of diagonal
                // block k+2 receives values nobody reads: the factorisation only ever touches lower triangles of diagonal blocks)
                g.M = rowsR; g.N = b1 + b2; g.flags = 0;
                g.A = at(s2, s0); g.B = at(s1, s0); g.C = at(s2, s1);
                HD_CALL(hd_gemm_nt(st, g));
                HD_CUDA(cudaEventRecord(g_evs[2][k], st));
                HD_CUDA(cudaEventRecord(g_evs[3][k], st));
            } else {
                g.M = rowsR; g.N = b1; g.flags = 0;
                g.A = at(s2, s0); g.B = at(s1, s0); g.C = at(s2, s1);
                HD_CALL(hd_gemm_nt(st, g));                                                    // UPDCOL(k)
                HD_CUDA(cudaEventRecord(g_evs[2][k], st));
                g.M = b2; g.N = b2; g.flags = HD_GEMM_LOWER;
                g.A = at(s2, s0); g.B = at(s2, s0); g.C = at(s2, s2);
                HD_CALL(hd_gemm_nt(st, g));                                                    // UPDCOL2(k): diagonal block k+2 ...
                if (rows3 > 0) {
                    g.M = rows3; g.N = b2; g.flags = 0;
                    g.A = at(s3, s0); g.B = at(s2, s0); g.C = at(s3, s2);
                    HD_CALL(hd_gemm_nt(st, g));                                                // ... and the rows below it
                }
                HD_CUDA(cudaEventRecord(g_evs[3][k], st));
            }
            if (rows3 > 0) {
                g.M = rows3; g.N = rows3; g.flags = HD_GEMM_LOWER;
                g.A = at(s3, s0); g.B = at(s3, s0); g.C = at(s3, s3);
                HD_CALL(hd_gemm_nt(st, g));                                                    // UPDREST(k)
            }
        } else if (b1 > 0) {
            // last panel: nothing below the strip; the events the side stream may wait for still have to exist in stream order
            HD_CUDA(cudaStreamWaitEvent(st, g_evs[1][k], 0));
            HD_CUDA(cudaEventRecord(g_evs[2][k], st));
            HD_CUDA(cudaEventRecord(g_evs[3][k], st));
        }
    }
    HD_CUDA(cudaStreamWaitEvent(st, g_evs[0][nblk - 1], 0)); // join
    if (part) {
        if (!g_ev_join) HD_CUDA(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
        HD_CUDA(cudaEventRecord(g_ev_join, st));
        HD_CUDA(cudaStreamWaitEvent(caller, g_ev_join, 0));
    }
    return HD_OK;
}

namespace {
__global__ void ldl_floor_kernel(const double *A, long ld, int n, double *floorp, int *nperturb, int nb, int rank, int nranks) {
    __shared__ double red[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256)
        if (nranks <= 1 || (i / nb) % nranks == rank) v = fmax(v, fabs(A[(long) i * ld + i])); // only owned columns hold the matrix
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) { *floorp = fmax(1e-13 * red[0], 1e-300); *nperturb = 0; }
}
__global__ void sign_scale_kernel(double *x, long ldx, int n, int nrhs, const double *__restrict__ sgn) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        for (int r = 0; r < nrhs; ++r) x[(long) r * ldx + i] *= sgn[i];
}
__global__ void count_negative_kernel(const double *__restrict__ sgn, int n, int *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && sgn[i] < 0.0) atomicAdd(out, 1);
}
} // namespace

// allocate the LDL state on first use and set the static-pivoting floor from the (owned part of the) diagonal in c->L
int chol_ldl_prepare(cudaStream_t st, DenseChol *c, int nb, int rank, int nranks) {
    if (!c->dfloor) {
        HD_CUDA(cudaMalloc(&c->dfloor, sizeof(double)));
        HD_CUDA(cudaMalloc(&c->dperturb, 2 * sizeof(int)));
    }
    HDK(ldl_floor_kernel)<<<1, 256, 0, st>>>(c->L, c->np, c->n, c->dfloor, c->dperturb, nb, rank, nranks);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// Cholesky (c->ldl == false; *info = LAPACK dpotrf info) or LDL^T with unit-magnitude D (c->ldl == true; always
// "succeeds": *info = 0, c->nperturbed / c->nnegative report the static pivots and the inertia)
// everything chol_factor enqueues for the factorisation proper (main + side stream); also what gets captured in a graph
static int enqueue_factor(cudaStream_t st, DenseChol *c, int nb) {
    HD_CUDA(cudaMemsetAsync(c->dinfo, 0, sizeof(int), st));
    LdlCtx ctx{};
    if (c->ldl) {
        HD_CALL(chol_ldl_prepare(st, c, HD_LEAF, 0, 1));
        ctx.sgn = c->sgn; ctx.dinv_base = c->Dinv; ctx.floorp = c->dfloor; ctx.nperturb = c->dperturb;
        g_ldl = &ctx;
    }
    int rc;
    const int sched = g_sched > 0 ? g_sched : (c->np <= 10240 ? 3 : 1);
    if (nb >= HD_LEAF && c->np >= 4 * nb && !c->ldl && sched >= 2)
        rc = potrf_blocked2(st, c->L, c->np, c->np, c->Dinv, c->dinfo, (nb / HD_LEAF) * HD_LEAF, 0);
    else if (nb >= HD_LEAF && c->np >= 4 * nb)
        rc = potrf_blocked(st, c->L, c->np, c->np, c->Dinv, c->dinfo, (nb / HD_LEAF) * HD_LEAF, 0, g_tail && g_lookahead_nb < 0);
    else
        rc = potrf_rec(st, c->L, c->np, c->np, c->Dinv, c->dinfo, 0);
    g_ldl = nullptr;
    return rc;
}

static int g_use_graph = 1;
static int g_graph_max = 6144;
void hd_chol_set_graph(int on) { g_use_graph = on != 0; if (on > 1) g_graph_max = on; }

int chol_factor(cudaStream_t st, DenseChol *c, int *info) {
    HD_CALL(ensure_leaf_attr());
    int nb = g_lookahead_nb;
    // measured (tools/trace_potrf.py, B200): strip schedule with NB = 128 up to 7k, 256 up to 10k; one-step look-ahead beyond
    if (nb < 0) nb = auto_block(c->np);
    // Up to n = 6k the factorisation is a launch-bound chain of a few hundred small kernels on two streams whose shapes
    // depend only on (n, block, mode): the second call with the same configuration captures it into a CUDA graph, later
    // calls replay the graph (one launch, dependencies resolved on the device).  The first call runs eagerly so that
    // every lazy allocation / function attribute exists before the capture.  Measured (tools/probe_block.py): -4 % at n = 1500,
    // -3 % at 4096, +2 % at 8192 and beyond (stream priorities are not honoured inside a graph), hence the size limit.
    const unsigned long long key = 1ull | ((unsigned long long) nb << 8) | ((unsigned long long) (c->ldl ? 1 : 0) << 1) |
                                   ((unsigned long long) g_leaf_version << 2) | ((unsigned long long) hd_gemm_get_variant() << 4) |
                                   ((unsigned long long) (g_sched + 1) << 40) | ((unsigned long long) g_ldl_pivot << 44) | ((unsigned long long) (g_partition > 0) << 45) | ((unsigned long long) g_tail << 46);
    const bool graph_ok = g_use_graph && c->np <= g_graph_max && getenv("HDSDPCU_TRACE") == nullptr;
    if (graph_ok && c->graph_exec && c->graph_key == key) {
        HD_CUDA(cudaGraphLaunch((cudaGraphExec_t) c->graph_exec, st));
        g_hd_launches += c->graph_kernels; // the kernels inside the replayed graph
    } else if (graph_ok && c->eager_key == key) {
        if (c->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t) c->graph_exec); c->graph_exec = nullptr; }
        cudaGraph_t graph = nullptr;
        HD_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const long before = g_hd_launches;
        int rc = enqueue_factor(st, c, nb);
        c->graph_kernels = g_hd_launches - before;
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != HD_OK || ce != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            fprintf(stderr, "[hdsdpcu] graph capture of the factorisation failed (%s); running eagerly\n", cudaGetErrorString(ce));
            g_use_graph = 0;
            HD_CALL(enqueue_factor(st, c, nb));
        } else {
            cudaGraphExec_t exec = nullptr;
            HD_CUDA(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            c->graph_exec = exec; c->graph_key = key;
            HD_CUDA(cudaGraphLaunch(exec, st));
        }
    } else {
        HD_CALL(enqueue_factor(st, c, nb));
        c->eager_key = key;
    }
    HDK(leaf_transpose_all_kernel)<<<dim3(4, 4, c->np / HD_LEAF), dim3(32, 8), 0, st>>>(c->Dinv, c->DinvT);
    HD_CUDA(cudaGetLastError());
    HD_CUDA(cudaMemcpyAsync(c->hinfo, c->dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    int counts[2] = {0, 0};
    if (c->ldl) {
        HD_CUDA(cudaMemsetAsync(c->dperturb + 1, 0, sizeof(int), st));
        HDK(count_negative_kernel)<<<(c->n + 255) / 256, 256, 0, st>>>(c->sgn, c->n, c->dperturb + 1);
        HD_CUDA(cudaMemcpyAsync(counts, c->dperturb, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    HD_CUDA(cudaStreamSynchronize(st));
    int inf = *c->hinfo;
    if (inf > c->n) inf = 0; // padding pivots are exactly 1
    if (c->ldl) { inf = 0; c->nperturbed = counts[0]; c->nnegative = counts[1]; }
    if (info) *info = inf;
    c->factored = (inf == 0);
    return HD_OK;
}

int chol_dsolve(cudaStream_t st, DenseChol *c, double *x, int nrhs, long ldx) {
    if (!c->ldl) return HD_OK;
    HDK(sign_scale_kernel)<<<(c->n + 255) / 256, 256, 0, st>>>(x, ldx, c->n, nrhs, c->sgn);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int chol_invert(cudaStream_t st, DenseChol *c, double *inv) {
    if (c->ldl) { // X X^T would drop J = diag(+-1): the inverse of an L J L^T factor is not provided (nothing on the path needs it)
        fprintf(stderr, "[hdsdpcu] chol_invert: not available for an LDL^T factorisation\n");
        return HD_FAILED;
    }
    HD_CALL(chol_ensure_work(c));
    // X = L^-T (upper triangular) in work; blocks below the diagonal leaves must read as zero
    HD_CUDA(cudaMemsetAsync(c->work, 0, (size_t) c->np * c->np * sizeof(double), st));
    HD_CALL(invert_rec(st, c->L, c->np, c->np, c->Dinv, c->work, c->np));
    GemmArgs g{};
    g.M = c->np; g.N = c->np; g.K = c->np;
    g.A = c->work; g.lda = c->np; g.B = c->work; g.ldb = c->np; g.C = inv; g.ldc = c->np;
    g.alpha = 1.0; g.beta = 0.0; g.flags = HD_GEMM_LOWER | HD_GEMM_KTRI_MAX;
    HD_CALL(hd_gemm_nt(st, g));
    return hd_symmetrize_lower(st, inv, c->np, c->np);
}

int chol_fsolve(cudaStream_t st, DenseChol *c, double *x, int nrhs, long ldx) {
    return hd_trsv(st, false, c->L, c->np, c->Dinv, c->DinvT, c->np, x, nrhs, ldx, c->sync);
}
int chol_bsolve(cudaStream_t st, DenseChol *c, double *x, int nrhs, long ldx) {
    return hd_trsv(st, true, c->L, c->np, c->Dinv, c->DinvT, c->np, x, nrhs, ldx, c->sync);
}

int chol_logdet(cudaStream_t st, DenseChol *c, double *dlogdet, double *ddiag) {
    HDK(logdet_kernel)<<<1, 256, 0, st>>>(c->L, c->np, c->n, dlogdet, ddiag);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}
