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

template <typename K> int set_smem_once(K kernel) {
    HD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    return HD_OK;
}
int g_trsv_version = 2;

} // namespace

void hd_trsv_set_version(int v) { g_trsv_version = v; }

// sync: device int buffer of at least nblk + 1 entries
int hd_trsv(cudaStream_t st, bool transposed, const double *L, long ldl, const double *Dinv, const double *DinvT, int np,
            double *x, int nrhs, long ldx, int *sync) {
    const int nblk = np / HD_LEAF;
    int r0 = 0;
    while (r0 < nrhs) {
        int nb = nrhs - r0;
        if (transposed) nb = nb >= 2 ? 2 : 1; else nb = nb >= 4 ? 4 : (nb >= 2 ? 2 : 1);
        HD_CUDA(cudaMemsetAsync(sync, 0, sizeof(int) * (nblk + 1), st));
        ++g_hd_launches;
        double *xr = x + (long) r0 * ldx;
        if (g_trsv_version == 2) {
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
