// hdsdp_b200/csrc/trsv.cu -- one-launch triangular solves  L x = b  and  L^T x = b  for a few right-hand sides.
//
// Replaces dtrsv/dpotrs behind the reference's HFpLinsysSolve / FSolve / BSolve
// (linalg/hdsdp_linsolver.c:1146-1222) and, for M, the PCG loop of conjGradSolve (:1446-1588).
//
// HBM-bound: the algorithmic traffic is one read of the lower triangle of L per pass (4 n^2 bytes).
// One CTA owns one 128-row block of x.  CTAs take their block index from a ticket counter, so a CTA only ever
// waits on blocks that were started before it (deadlock-free without a cooperative launch): CTA i consumes the
// published x_j (j < i) in order, streaming its 128 x 128 tiles of L, then applies the explicit inverse of its
// diagonal leaf (stored by the factorisation) and publishes x_i through a release flag.  The dependency chain per
// block is one tile product + one leaf product; all other tile traffic overlaps across the resident CTAs.
#include "common.h"

namespace {

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// sync[0] = ticket counter, sync[1 + j] = ready flag of block j
template <int NRHS>
__global__ void __launch_bounds__(128) trsv_fwd_kernel(const double *__restrict__ L, long ldl, const double *__restrict__ Dinv,
                                                      double *x, long ldx, int nblk, int *sync) {
    __shared__ int s_blk;
    __shared__ double xs[NRHS][HD_LEAF];
    const int t = threadIdx.x;
    if (t == 0) s_blk = atomicAdd(&sync[0], 1);
    __syncthreads();
    const int i = s_blk;
    if (i >= nblk) return;
    double acc[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) acc[r] = x[(long) r * ldx + (long) i * HD_LEAF + t];
    const double *Lrow = L + (long) i * HD_LEAF + t;
    for (int j = 0; j < i; ++j) {
        if (t == 0) while (ld_acquire(&sync[1 + j]) == 0) { }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < NRHS; ++r) xs[r][t] = __ldcg(&x[(long) r * ldx + (long) j * HD_LEAF + t]);
        __syncthreads();
        const double *Lt = Lrow + (long) j * HD_LEAF * ldl;
#pragma unroll 16
        for (int k = 0; k < HD_LEAF; ++k) {
            const double l = __ldcs(&Lt[(long) k * ldl]);
#pragma unroll
            for (int r = 0; r < NRHS; ++r) acc[r] -= l * xs[r][k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xs[r][t] = acc[r];
    __syncthreads();
    const double *D = Dinv + (long) i * HD_LEAF * HD_LEAF + t; // x_i[t] = sum_{k<=t} Dinv[t,k] acc[k]
    double out[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) out[r] = 0.0;
#pragma unroll 8
    for (int k = 0; k < HD_LEAF; ++k) {
        const double d = D[k * HD_LEAF]; // exact zeros above the diagonal
#pragma unroll
        for (int r = 0; r < NRHS; ++r) out[r] += d * xs[r][k];
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) x[(long) r * ldx + (long) i * HD_LEAF + t] = out[r];
    __threadfence();
    __syncthreads();
    if (t == 0) st_release(&sync[1 + i], 1);
}

// L^T x = b: block i (descending).  Warp w owns columns c = 32 w .. 32 w + 31 of the tile column; lanes stride the
// 128 rows of every tile L[j-block, i-block] (coalesced along the rows), partial sums per (column, lane), reduced by
// shuffles once at the end.  DinvT holds the transposed leaf inverses so the final product is coalesced as well.
template <int NRHS>
__global__ void __launch_bounds__(128) trsv_bwd_kernel(const double *__restrict__ L, long ldl, const double *__restrict__ DinvT,
                                                      double *x, long ldx, int nblk, int *sync) {
    __shared__ int s_blk;
    __shared__ double xs[NRHS][HD_LEAF];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) s_blk = atomicAdd(&sync[0], 1);
    __syncthreads();
    if (s_blk >= nblk) return;
    const int i = nblk - 1 - s_blk;
    double acc[NRHS][32];
#pragma unroll
    for (int r = 0; r < NRHS; ++r)
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[r][c] = 0.0;
    for (int j = nblk - 1; j > i; --j) {
        if (t == 0) while (ld_acquire(&sync[1 + j]) == 0) { }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < NRHS; ++r) xs[r][t] = __ldcg(&x[(long) r * ldx + (long) j * HD_LEAF + t]);
        __syncthreads();
        const double *Lt = L + ((long) i * HD_LEAF + w * 32) * ldl + (long) j * HD_LEAF + lane;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double *col = Lt + (long) c * ldl;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double l = __ldcs(&col[32 * q]);
#pragma unroll
                for (int r = 0; r < NRHS; ++r) acc[r][c] += l * xs[r][32 * q + lane];
            }
        }
    }
    __syncthreads();
    // reduce over lanes; column c of warp w -> element w*32 + c of the block
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            double s = acc[r][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == c) xs[r][w * 32 + c] = x[(long) r * ldx + (long) i * HD_LEAF + w * 32 + c] - s;
        }
    }
    __syncthreads();
    const double *D = DinvT + (long) i * HD_LEAF * HD_LEAF + t; // x_i[t] = sum_{k>=t} Dinv[k,t] rhs[k] = sum_k DinvT[t,k] rhs[k]
    double out[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) out[r] = 0.0;
#pragma unroll 8
    for (int k = 0; k < HD_LEAF; ++k) {
        const double d = D[k * HD_LEAF];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) out[r] += d * xs[r][k];
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) x[(long) r * ldx + (long) i * HD_LEAF + t] = out[r];
    __threadfence();
    __syncthreads();
    if (t == 0) st_release(&sync[1 + i], 1);
}


// ------------------------------------------------------------------------------------------------------------
// v2: same one-launch / ticket / flag protocol, but the dependency chain per 128-row block is made of on-chip work
// only.  256 threads per CTA: the NEXT 128 x 128 tile of L is already in registers (64 doubles per thread, loads
// issued before the CTA spins on the flag of the block it multiplies) and the inverse leaf of the CTA's own
// diagonal block is staged in shared memory once, so after a flag flips a block costs: one L2 read of x_j, 64 FMAs
// per thread, a shared-memory reduction and, for the last tile, the leaf product from shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int T2_THREADS = 256;
constexpr int T2_SMEM = (HD_LEAF * HD_LEAF) * 8;

template <int NRHS>
__global__ void __launch_bounds__(T2_THREADS, 1) trsv_fwd2_kernel(const double *__restrict__ L, long ldl, const double *__restrict__ Dinv,
                                                                 double *x, long ldx, int nblk, int *sync) {
    extern __shared__ __align__(16) double Dsh[];      // Dinv_i, column-major 128 x 128 (exact zeros above the diagonal)
    __shared__ int s_blk;
    __shared__ double xs[NRHS][HD_LEAF];
    __shared__ double red[NRHS][2][HD_LEAF];
    const int t = threadIdx.x, row = t & 127, half = t >> 7;
    if (t == 0) s_blk = atomicAdd(&sync[0], 1);
    __syncthreads();
    const int i = s_blk;
    if (i >= nblk) return;
    {   // stage the inverse leaf (off the critical path: overlaps with the first tiles)
        const double *D = Dinv + (long) i * HD_LEAF * HD_LEAF;
        for (int e = t; e < HD_LEAF * HD_LEAF / 2; e += T2_THREADS)
            reinterpret_cast<double2 *>(Dsh)[e] = __ldcs(reinterpret_cast<const double2 *>(D) + e);
    }
    double acc[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
    const double *Lrow = L + (long) i * HD_LEAF + row + (long) (half * 64) * ldl;
    double tl[64];
    if (i > 0) {
#pragma unroll
        for (int q = 0; q < 64; ++q) tl[q] = __ldcs(&Lrow[(long) q * ldl]);
    }
    for (int j = 0; j < i; ++j) {
        if (t == 0) while (ld_acquire(&sync[1 + j]) == 0) { }
        __syncthreads();
        if (t < HD_LEAF) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) xs[r][t] = __ldcg(&x[(long) r * ldx + (long) j * HD_LEAF + t]);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 64; ++q) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) acc[r] += tl[q] * xs[r][half * 64 + q];
        }
        if (j + 1 < i) {
            const double *Lt = Lrow + (long) (j + 1) * HD_LEAF * ldl;
#pragma unroll
            for (int q = 0; q < 64; ++q) tl[q] = __ldcs(&Lt[(long) q * ldl]);
        }
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) red[r][half][row] = acc[r];
    __syncthreads();
    if (t < HD_LEAF) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) xs[r][t] = x[(long) r * ldx + (long) i * HD_LEAF + t] - (red[r][0][t] + red[r][1][t]);
    }
    __syncthreads();
    double out[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) out[r] = 0.0;
    const double *Dr = Dsh + row + (half * 64) * HD_LEAF;
#pragma unroll 16
    for (int k = 0; k < 64; ++k) {
        const double dv = Dr[k * HD_LEAF];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) out[r] += dv * xs[r][half * 64 + k];
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) red[r][half][row] = out[r];
    __syncthreads();
    if (t < HD_LEAF) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) x[(long) r * ldx + (long) i * HD_LEAF + t] = red[r][0][t] + red[r][1][t];
    }
    __threadfence();
    __syncthreads();
    if (t == 0) st_release(&sync[1 + i], 1);
}

// L^T x = b, block i descending.  Tile L[j-block rows, i-block cols]: warp w (of 8) owns columns 16 w .. 16 w + 15,
// lane l the rows l, l + 32, l + 64, l + 96 (coalesced); 16 partial sums per thread and right-hand side, reduced over
// the lanes once at the end.
template <int NRHS>
__global__ void __launch_bounds__(T2_THREADS, 1) trsv_bwd2_kernel(const double *__restrict__ L, long ldl, const double *__restrict__ DinvT,
                                                                 double *x, long ldx, int nblk, int *sync) {
    extern __shared__ __align__(16) double Dsh[];      // DinvT_i : DinvT[t, k] = Dinv[k, t], column-major
    __shared__ int s_blk;
    __shared__ double xs[NRHS][HD_LEAF];
    __shared__ double red[NRHS][2][HD_LEAF];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, row = t & 127, half = t >> 7;
    if (t == 0) s_blk = atomicAdd(&sync[0], 1);
    __syncthreads();
    if (s_blk >= nblk) return;
    const int i = nblk - 1 - s_blk;
    {
        const double *D = DinvT + (long) i * HD_LEAF * HD_LEAF;
        for (int e = t; e < HD_LEAF * HD_LEAF / 2; e += T2_THREADS)
            reinterpret_cast<double2 *>(Dsh)[e] = __ldcs(reinterpret_cast<const double2 *>(D) + e);
    }
    double acc[NRHS][16];
#pragma unroll
    for (int r = 0; r < NRHS; ++r)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[r][c] = 0.0;
    const double *Lcol = L + ((long) i * HD_LEAF + w * 16) * ldl + lane;
    double tl[64];
    if (i < nblk - 1) {
        const double *Lt = Lcol + (long) (nblk - 1) * HD_LEAF;
#pragma unroll
        for (int c = 0; c < 16; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) tl[c * 4 + q] = __ldcs(&Lt[(long) c * ldl + 32 * q]);
    }
    for (int j = nblk - 1; j > i; --j) {
        if (t == 0) while (ld_acquire(&sync[1 + j]) == 0) { }
        __syncthreads();
        if (t < HD_LEAF) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) xs[r][t] = __ldcg(&x[(long) r * ldx + (long) j * HD_LEAF + t]);
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < NRHS; ++r) acc[r][c] += tl[c * 4 + q] * xs[r][32 * q + lane];
        if (j - 1 > i) {
            const double *Lt = Lcol + (long) (j - 1) * HD_LEAF;
#pragma unroll
            for (int c = 0; c < 16; ++c)
#pragma unroll
                for (int q = 0; q < 4; ++q) tl[c * 4 + q] = __ldcs(&Lt[(long) c * ldl + 32 * q]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double sacc = acc[r][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == c) xs[r][w * 16 + c] = x[(long) r * ldx + (long) i * HD_LEAF + w * 16 + c] - sacc;
        }
    }
    __syncthreads();
    double out[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) out[r] = 0.0;
    const double *Dr = Dsh + row + (half * 64) * HD_LEAF;
#pragma unroll 16
    for (int k = 0; k < 64; ++k) {
        const double dv = Dr[k * HD_LEAF];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) out[r] += dv * xs[r][half * 64 + k];
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) red[r][half][row] = out[r];
    __syncthreads();
    if (t < HD_LEAF) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) x[(long) r * ldx + (long) i * HD_LEAF + t] = red[r][0][t] + red[r][1][t];
    }
    __threadfence();
    __syncthreads();
    if (t == 0) st_release(&sync[1 + i], 1);
}

// ------------------------------------------------------------------------------------------------------------
// v3: one work item per 128 x 128 TILE instead of one CTA per row block.  v2 streams row block i through one CTA (i tiles
// in sequence), so the last row blocks -- the longest -- run alone at the end, each limited to what one SM can pull from HBM
// (68 % of the HBM roofline at m = 50 000).  Here persistent CTAs take tickets over the tiles in the order
//   column 0: diag(0), (1,0), (2,0), ... ; column 1: diag(1), (2,1), ...        (L^T x = b: mirrored, last column first)
// so every dependency of an item has a smaller ticket (no deadlock however few CTAs are resident):
//   diag(j)  : waits until all updates of block j are applied (cnt[j]), x_j <- Dinv_j x_j, publishes ready[j];
//   (i, j)   : its tile is already in registers; waits for ready[j], multiplies, then waits for its TURN on block i
//              (cnt[i] == number of updates applied so far) and subtracts -- the turn order makes the sum deterministic.
// The next item's tile loads are issued before the current item waits for its turn.  All tiles off the critical chain
// diag(j) -> (j+1, j) -> diag(j+1) stream in the background at full bandwidth.
// sync layout: [0] ticket, [1 .. nblk] ready, [1 + nblk .. 2 nblk] cnt, [1 + 2 nblk] error flag (spin timeout)
// ------------------------------------------------------------------------------------------------------------
constexpr int T3_THREADS = 256;
// updates block i has already received when the one from block j is due: forward j = 0 .. i-1 in order, transposed j = nblk-1 .. i+1
__device__ __forceinline__ int need_i_of(bool trans, int nblk, int i, int j) { return trans ? nblk - 1 - j : j; }
__device__ __forceinline__ unsigned long long t3_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want (thread 0 of the CTA); gives up after 10 s and raises the error flag instead of hanging the GPU
__device__ __forceinline__ void t3_wait(int *p, int want, int *err) {
    if (ld_acquire(p) >= want) return;
    const unsigned long long t0 = t3_now();
    int spins = 0;
    while (ld_acquire(p) < want) {
        if ((++spins & 1023) == 0 && t3_now() - t0 > 10000000000ull) { atomicExch(err, 1); break; }
    }
}

template <int NRHS, bool TRANS>
__global__ void __launch_bounds__(T3_THREADS, 1) trsv3_kernel(const double *__restrict__ L, long ldl, const double *__restrict__ Dleaf,
                                                             double *x, long ldx, int nblk, int *sync) {
    __shared__ int s_t;
    __shared__ double vs[NRHS][HD_LEAF];          // input block
    __shared__ double red[NRHS][4][HD_LEAF];      // partial results (2 halves, or 4 warps of a half for the transposed product)
    int *ready = sync + 1, *cnt = sync + 1 + nblk, *err = sync + 1 + 2 * nblk;
    const int t = threadIdx.x, row = t & 127, half = t >> 7, lane = t & 31, w4 = (t >> 5) & 3;
    const long nitems = (long) nblk * (nblk + 1) / 2;
    // item of a ticket: column c (nblk - c items: the diagonal one first), position r in it
    auto decode = [&](long tk, int &c, int &r) {
        const double b = 2.0 * nblk + 1.0;
        c = (int) ((b - sqrt(b * b - 8.0 * (double) tk)) * 0.5);
        if (c < 0) c = 0;
        if (c > nblk - 1) c = nblk - 1;
        while (c > 0 && (long) c * nblk - (long) c * (c - 1) / 2 > tk) --c;
        while ((long) (c + 1) * nblk - (long) (c + 1) * c / 2 <= tk) ++c;
        r = (int) (tk - ((long) c * nblk - (long) c * (c - 1) / 2));
    };
    auto tile_ptr = [&](int c, int r) -> const double * {
        const int j = TRANS ? nblk - 1 - c : c;                     // block whose solution the item reads (off-diagonal) or produces
        if (r == 0) return Dleaf + (long) j * HD_LEAF * HD_LEAF;    // Dinv_j (forward) / DinvT_j (transposed): plain product
        const int i = TRANS ? j - r : j + r;                        // block the item updates
        return TRANS ? L + (long) i * HD_LEAF * ldl + (long) j * HD_LEAF   // L[j-block rows, i-block cols], used transposed
                     : L + (long) j * HD_LEAF * ldl + (long) i * HD_LEAF;  // L[i-block rows, j-block cols]
    };
    double tl[64];
    if (t == 0) s_t = atomicAdd(&sync[0], 1);
    __syncthreads();
    long tk = s_t;
    int c = 0, r = 0;
    if (tk < nitems) {
        decode(tk, c, r);
        const double *T = tile_ptr(c, r) + row + (long) (half * 64) * (r == 0 ? HD_LEAF : ldl);
        const long ld = r == 0 ? HD_LEAF : ldl;
#pragma unroll
        for (int q = 0; q < 64; ++q) tl[q] = __ldcs(&T[(long) q * ld]);
    }
    while (tk < nitems) {
        const int j = TRANS ? nblk - 1 - c : c;
        const int i = r == 0 ? j : (TRANS ? j - r : j + r);
        const int need_i = TRANS ? nblk - 1 - j : j;               // updates block i has received when it is this item's turn
        // ---- input: diag reads its own, fully updated block; an off-diagonal item reads the finished x_j ----
        if (t == 0) { if (r == 0) t3_wait(&cnt[j], need_i, err); else t3_wait(&ready[j], 1, err); }
        __syncthreads();
        if (t < HD_LEAF) {
#pragma unroll
            for (int v = 0; v < NRHS; ++v) vs[v][t] = __ldcg(&x[(long) v * ldx + (long) j * HD_LEAF + t]);
        }
        __syncthreads();
        if (!TRANS || r == 0) {
            // y = T v : 64 columns per thread, the two halves added through shared memory
#pragma unroll
            for (int v = 0; v < NRHS; ++v) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 64; ++q) acc = fma(tl[q], vs[v][half * 64 + q], acc);
                red[v][half][row] = acc;
            }
        } else {
            // y = T^T v : column sums over the 128 rows; over the 32 lanes of a warp by recursive halving (62 shuffles for the
            // 64 columns of a thread), over the 4 warps of a half through shared memory.  Lane l ends with columns 2 l, 2 l + 1.
#pragma unroll
            for (int v = 0; v < NRHS; ++v) {
                const double xv = vs[v][row];
                double p[32];
                const bool up16 = (lane & 16) != 0;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const double lo = tl[q] * xv, hi = tl[q + 32] * xv;
                    p[q] = (up16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, up16 ? lo : hi, 16);
                }
#pragma unroll
                for (int sd = 8; sd >= 1; sd >>= 1) {
                    const bool up = (lane & sd) != 0;
#pragma unroll
                    for (int q = 0; q < 2 * sd; ++q) p[q] = (up ? p[q + 2 * sd] : p[q]) + __shfl_xor_sync(0xffffffffu, up ? p[q] : p[q + 2 * sd], sd);
                }
                red[v][w4][half * 64 + 2 * lane] = p[0];
                red[v][w4][half * 64 + 2 * lane + 1] = p[1];
            }
        }
        // ---- the tile is consumed: take the next ticket and start its loads before waiting for this item's turn ----
        __syncthreads();
        if (t == 0) s_t = atomicAdd(&sync[0], 1);
        __syncthreads();
        const long tk2 = s_t;
        int c2 = 0, r2 = 0;
        const bool transposed_item = TRANS && r != 0;
        double res[NRHS];
        if (t < HD_LEAF) {
#pragma unroll
            for (int v = 0; v < NRHS; ++v)
                res[v] = transposed_item ? (red[v][0][t] + red[v][1][t]) + (red[v][2][t] + red[v][3][t]) : red[v][0][t] + red[v][1][t];
        }
        if (tk2 < nitems) {
            decode(tk2, c2, r2);
            const double *T = tile_ptr(c2, r2) + row + (long) (half * 64) * (r2 == 0 ? HD_LEAF : ldl);
            const long ld = r2 == 0 ? HD_LEAF : ldl;
#pragma unroll
            for (int q = 0; q < 64; ++q) tl[q] = __ldcs(&T[(long) q * ld]);
        }
        // ---- output ----
        if (r != 0 && t == 0) t3_wait(&cnt[i], need_i_of(TRANS, nblk, i, j), err);
        __syncthreads();
        if (t < HD_LEAF) {
#pragma unroll
            for (int v = 0; v < NRHS; ++v) {
                double *dst = &x[(long) v * ldx + (long) i * HD_LEAF + t];
                *dst = r == 0 ? res[v] : __ldcg(dst) - res[v];
            }
        }
        __threadfence();
        __syncthreads();
        if (t == 0) {
            if (r == 0) st_release(&ready[j], 1);
            else st_release(&cnt[i], need_i_of(TRANS, nblk, i, j) + 1);
        }
        tk = tk2; c = c2; r = r2;
    }
}

template <typename K> int set_smem_once(K kernel) {
    HD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    return HD_OK;
}
int g_trsv_version = 2;

} // namespace

void hd_trsv_set_version(int v) { g_trsv_version = v; }

// sync: device int buffer of at least 2 nblk + 2 entries
int hd_trsv(cudaStream_t st, bool transposed, const double *L, long ldl, const double *Dinv, const double *DinvT, int np,
            double *x, int nrhs, long ldx, int *sync) {
    const int nblk = np / HD_LEAF;
    int r0 = 0;
    while (r0 < nrhs) {
        int nb = nrhs - r0;
        if (transposed) nb = nb >= 2 ? 2 : 1; else nb = nb >= 4 ? 4 : (nb >= 2 ? 2 : 1);
        HD_CUDA(cudaMemsetAsync(sync, 0, sizeof(int) * (2 * nblk + 2), st));
        ++g_hd_launches;
        double *xr = x + (long) r0 * ldx;
        if (g_trsv_version == 3) {
            static int sms = 0;
            if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
            const long nitems = (long) nblk * (nblk + 1) / 2;
            const int grid = (int) (nitems < sms ? nitems : sms);
            if (!transposed) {
                if (nb == 4) trsv3_kernel<4, false><<<grid, T3_THREADS, 0, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
                else if (nb == 2) trsv3_kernel<2, false><<<grid, T3_THREADS, 0, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
                else trsv3_kernel<1, false><<<grid, T3_THREADS, 0, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
            } else {
                if (nb == 2) trsv3_kernel<2, true><<<grid, T3_THREADS, 0, st>>>(L, ldl, DinvT, xr, ldx, nblk, sync);
                else trsv3_kernel<1, true><<<grid, T3_THREADS, 0, st>>>(L, ldl, DinvT, xr, ldx, nblk, sync);
            }
        } else if (g_trsv_version == 2) {
            static unsigned long long attr = 0;
            int dev = 0;
            cudaGetDevice(&dev);
            if (!(attr >> (dev & 63) & 1ull)) {
                HD_CALL(set_smem_once(trsv_fwd2_kernel<1>)); HD_CALL(set_smem_once(trsv_fwd2_kernel<2>)); HD_CALL(set_smem_once(trsv_fwd2_kernel<4>));
                HD_CALL(set_smem_once(trsv_bwd2_kernel<1>)); HD_CALL(set_smem_once(trsv_bwd2_kernel<2>));
                attr |= 1ull << (dev & 63);
            }
            if (!transposed) {
                if (nb == 4) trsv_fwd2_kernel<4><<<nblk, T2_THREADS, T2_SMEM, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
                else if (nb == 2) trsv_fwd2_kernel<2><<<nblk, T2_THREADS, T2_SMEM, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
                else trsv_fwd2_kernel<1><<<nblk, T2_THREADS, T2_SMEM, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
            } else {
                if (nb == 2) trsv_bwd2_kernel<2><<<nblk, T2_THREADS, T2_SMEM, st>>>(L, ldl, DinvT, xr, ldx, nblk, sync);
                else trsv_bwd2_kernel<1><<<nblk, T2_THREADS, T2_SMEM, st>>>(L, ldl, DinvT, xr, ldx, nblk, sync);
            }
        } else if (!transposed) {
            if (nb == 4) trsv_fwd_kernel<4><<<nblk, 128, 0, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
            else if (nb == 2) trsv_fwd_kernel<2><<<nblk, 128, 0, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
            else trsv_fwd_kernel<1><<<nblk, 128, 0, st>>>(L, ldl, Dinv, xr, ldx, nblk, sync);
        } else {
            if (nb == 2) trsv_bwd_kernel<2><<<nblk, 128, 0, st>>>(L, ldl, DinvT, xr, ldx, nblk, sync);
            else trsv_bwd_kernel<1><<<nblk, 128, 0, st>>>(L, ldl, DinvT, xr, ldx, nblk, sync);
        }
        HD_CUDA(cudaGetLastError());
        r0 += nb;
    }
    return HD_OK;
}
