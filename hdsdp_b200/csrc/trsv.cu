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

} // namespace

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
        if (!transposed) {
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
