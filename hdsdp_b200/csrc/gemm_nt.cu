// hdsdp_b200/csrc/gemm_nt.cu -- FP64 tensor-core (DMMA) tile GEMM for sm_100a.
//
//   C (M x N) = alpha * A (M x K) * B (N x K)^T + beta * C          all column-major
//
// This one kernel carries every O(n^3) operation of the hot path:
//   * trailing updates / panel solves of the recursive Cholesky of M and S
//     (replaces dpotrf/dtrsm behind reference linalg/hdsdp_linsolver.c:1096,1158,1184),
//   * S^-1 = X X^T with X = L^-T  (replaces dpotri, hdsdp_linsolver.c:1250),
//   * the rank-one Schur products V^T = A^T S^-1 and G = A^T V with the Hadamard-square
//     epilogue M_ij += s_i s_j G_ij^2 (reference M2 column builder, hdsdp_conic_sdp.c:687-778).
//
// Design (B200): FP64 MMA exists only as mma.sync m8n8k4 (SASS DMMA.8x8x4; the sm_90 m16n8k*
// shapes lower to the same instruction on sm_100a, and tcgen05 has no f64 kind), so this is a
// register-accumulator kernel: 128x128x16 CTA tile, 8 warps of 64(m) x 32(n), 4-stage
// cp.async (LDGSTS 16 B) pipeline, shared tiles stored k-major with a row stride of 132
// doubles so the fragment loads (address = tig*132 + gid) are bank-conflict free.
// The MMA is issued "transposed" (MMA-M runs along n, MMA-N along m) so each thread owns two
// m-consecutive doubles per accumulator pair and the epilogue is 16-byte vector traffic on the
// column-major C.  Grid: 1-D, grouped 8x8 super-tiles for L2 reuse of the A/B panels.
#include "common.h"

int hd_num_sms();

namespace {

constexpr int BM = 128, BN = 128;
constexpr int LDS_ = 132;                    // padded smem row stride (doubles)
constexpr int GROUP = 8;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct TileMap {
    int tiles_m, tiles_n, lower;
};

// linear CTA index -> (tm, tn); returns false if this CTA has no tile.  Super-tiles are GROUP m-tiles x
// GROUP*(128/BN) n-tiles (a 1024 x 1024 block of C) so that consecutive CTAs share their A/B panels in L2.
template <int BN>
__device__ __forceinline__ bool map_tile(const TileMap &tmap, int bid, int &tm, int &tn) {
    constexpr int GN = GROUP * (128 / BN);
    const int gsq = GROUP * GN;
    int super = bid / gsq, within = bid % gsq;
    int sm_, sn_;
    if (tmap.lower) {
        // super-tiles of the lower triangle of a square super-grid, row by row
        int s = (int) ((sqrtf(8.0f * (float) super + 1.0f) - 1.0f) * 0.5f);
        while ((s + 1) * (s + 2) / 2 <= super) ++s;
        while (s * (s + 1) / 2 > super) --s;
        sm_ = s;
        sn_ = super - s * (s + 1) / 2;
    } else {
        int super_n = (tmap.tiles_n + GN - 1) / GN;
        sm_ = super / super_n;
        sn_ = super % super_n;
    }
    tm = sm_ * GROUP + within % GROUP;
    tn = sn_ * GN + within / GROUP;
    if (tm >= tmap.tiles_m || tn >= tmap.tiles_n) return false;
    if (tmap.lower && tn * BN > tm * BM + BM - 1) return false;
    return true;
}

// BN = 128: one CTA per SM.  BN = 64: two CTAs per SM, so one CTA's
// barrier stalls and C read-modify-write epilogue overlap with the other CTA's main loop.
template <int BN, int BK, int STAGES, int MINB, bool SIGNED, int WM = 64, bool PEER = false>
__global__ void __launch_bounds__((128 / WM) * (BN / 32) * 32, MINB) dgemm_nt_kernel(GemmArgs g, TileMap tmap) {
    constexpr int NT = (128 / WM) * (BN / 32) * 32; // warps: (128 / WM) along m x (BN / 32) along n
    constexpr int MJ = WM / 8;                      // 8-row MMA tiles per warp along m
    constexpr int LDB_ = BN + 4;                          // 132 / 68: both = 4 (mod 16) -> conflict-free fragment loads
    constexpr int STAGE_DOUBLES = BK * (LDS_ + LDB_);     // A tile + B tile
    extern __shared__ __align__(16) double smem[];
    int tm, tn;
    if (!map_tile<BN>(tmap, blockIdx.x, tm, tn)) return;
    const int m0 = tm * BM;
    int n0 = tn * BN;
    if (g.bc_nb > 0) {
        // block-cyclic N: local column block -> global column block (stride = nRanks * nb); tiles above the diagonal
        // (and past the matrix edge, which lies above it) have no work
        const int jb = n0 / g.bc_nb;
        n0 = jb * g.bc_stride + (n0 - jb * g.bc_nb);
        if ((g.flags & HD_GEMM_LOWER) && n0 > m0 + BM - 1) return;
    }

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = (warp % (128 / WM)) * WM; // warp offset along m
    const int wn = (warp / (128 / WM)) * 32; // warp offset along n

    int k_lo = 0;
    if (g.flags & HD_GEMM_KTRI_MAX) k_lo = max(m0, n0) & ~(BK - 1); // X[i,k] == 0 for k < i on both operands
    if (g.flags & HD_GEMM_KTRI_A) k_lo = m0 & ~(BK - 1);            // ... on A only (triangular inverse X11 times a full block)
    const int nk = (g.K - k_lo) / BK;

    const double *Ag = g.A + (long) k_lo * g.lda + m0;
    const double *Bg = g.B + (long) k_lo * g.ldb + n0;

    // cp.async mapping: A stage = BK k-rows x 128 doubles, B stage = BK k-rows x BN doubles, 16-byte chunks
    auto load_stage = [&](int stage, int kt) {
        double *As = smem + stage * STAGE_DOUBLES;
        double *Bs = As + BK * LDS_;
        const double *a = Ag + (long) kt * BK * g.lda;
        const double *b = Bg + (long) kt * BK * g.ldb;
#pragma unroll
        for (int i = 0; i < BK * 64 / NT; ++i) {
            int c = tid + i * NT;
            int kr = c >> 6, mc = (c & 63) * 2;
            cp_async16(As + kr * LDS_ + mc, a + (long) kr * g.lda + mc);
        }
#pragma unroll
        for (int i = 0; i < BK * (BN / 2) / NT; ++i) {
            int c = tid + i * NT;
            int kr = c / (BN / 2), nc = (c % (BN / 2)) * 2;
            cp_async16(Bs + kr * LDB_ + nc, b + (long) kr * g.ldb + nc);
        }
    };

    double acc[4][MJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < MJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    if (g.beta != 0.0 || (g.flags & HD_GEMM_EPI_HADSQ)) {
        // C is read in the epilogue: pull this thread's 64-byte segments of the tile into L2 now (4 lanes share a segment)
        if (tig == 0) {
            const double *Cp = g.C + (long) (n0 + wn + gid) * g.ldc + m0 + wm;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < MJ; ++j)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(Cp + (long) (8 * i) * g.ldc + 8 * j));
        }
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const double *As = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double *Bs = As + BK * LDS_;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double bf[4], af[MJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) bf[i] = Bs[(kk + tig) * LDB_ + wn + 8 * i + gid];
            if (SIGNED) {
                const double sg = __ldg(&g.ksign[k_lo + kt * BK + kk + tig]);
#pragma unroll
                for (int i = 0; i < 4; ++i) bf[i] *= sg;
            }
#pragma unroll
            for (int j = 0; j < MJ; ++j) af[j] = As[(kk + tig) * LDS_ + wm + 8 * j + gid];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < MJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], bf[i], af[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[m .. m+1][n] with n = n0+wn+8i+gid, m = m0+wm+8j+2*tig
    const bool lower = (g.flags & HD_GEMM_LOWER) != 0;
    const bool hadsq = (g.flags & HD_GEMM_EPI_HADSQ) != 0;
    const bool interior = !lower || (n0 + BN - 1 <= m0); // no entry of this tile lies above the diagonal
    double *Cw = g.C + (long) (n0 + wn + gid) * g.ldc + m0 + wm + 2 * tig;
    if (interior && !hadsq) {
        if (g.beta == 0.0) {
            const bool colscale = (g.flags & HD_GEMM_EPI_COLSCALE) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double cs = colscale ? g.alpha * g.sb[n0 + wn + 8 * i + gid] : g.alpha;
#pragma unroll
                for (int j = 0; j < MJ; ++j) {
                    double2 v = make_double2(cs * acc[i][j][0], cs * acc[i][j][1]);
                    *reinterpret_cast<double2 *>(Cw + (long) (8 * i) * g.ldc + 8 * j) = v;
                    if (PEER) // fused hand-off (own instantiation: keeps the pointer out of the register budget of the others): the same 16 bytes go to the peer GPU's copy of the panel
                        *reinterpret_cast<double2 *>(g.peerC + (Cw - g.C) + (long) (8 * i) * g.ldc + 8 * j) = v;
                }
            }
        } else {
            // read-modify-write in batches of 8 independent 16-byte loads (the C tile was prefetched into L2 above)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double2 old[MJ];
#pragma unroll
                for (int j = 0; j < MJ; ++j) old[j] = *reinterpret_cast<const double2 *>(Cw + (long) (8 * i) * g.ldc + 8 * j);
#pragma unroll
                for (int j = 0; j < MJ; ++j) {
                    double2 v = make_double2(g.alpha * acc[i][j][0] + g.beta * old[j].x, g.alpha * acc[i][j][1] + g.beta * old[j].y);
                    *reinterpret_cast<double2 *>(Cw + (long) (8 * i) * g.ldc + 8 * j) = v;
                }
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + wn + 8 * i + gid;
        double sbn = hadsq ? g.sb[n] : 0.0;
#pragma unroll
        for (int j = 0; j < MJ; ++j) {
            const int m = m0 + wm + 8 * j + 2 * tig;
            if (lower && m + 1 < n) continue;
            double2 *cp = reinterpret_cast<double2 *>(g.C + (long) n * g.ldc + m);
            double2 v;
            if (hadsq) {
                double2 old = *cp;
                v.x = old.x + g.sa[m] * sbn * acc[i][j][0] * acc[i][j][0];
                v.y = old.y + g.sa[m + 1] * sbn * acc[i][j][1] * acc[i][j][1];
            } else if (g.beta == 0.0) {
                v.x = g.alpha * acc[i][j][0];
                v.y = g.alpha * acc[i][j][1];
            } else {
                double2 old = *cp;
                v.x = g.alpha * acc[i][j][0] + g.beta * old.x;
                v.y = g.alpha * acc[i][j][1] + g.beta * old.y;
            }
            if (lower && m < n) {
                // m == n-1: only the second element (m+1 == n) is in the lower triangle
                g.C[(long) n * g.ldc + m + 1] = v.y;
            } else {
                *cp = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Thin products.  One SM delivers ~0.23 TFLOP/s of DMMA, so a 128 x 64 x 256 tile keeps its CTA busy for 18 us: the latency of
// the small GEMMs on the critical path of the factorisations (leaf triangular solves, 128 / 256-wide syrk and strip updates,
// chol.cu) is set by the tile size, not by the launch.  This kernel cuts the same products into 32 x 128 tiles (4 x more CTAs,
// 4 x shorter): 8 warps, warp w owns the 16 columns 16 w .. 16 w + 15 of all 32 rows, 3-stage cp.async pipeline over K chunks
// of 32.  C may alias A when N == 128 (the in-place leaf solves X <- X Dinv^T): a CTA reads only the rows it later writes.
// ------------------------------------------------------------------------------------------------------------
constexpr int TH_M = 32, TH_N = 128, TH_K = 32, TH_STAGES = 3, TH_LDA = TH_M + 4, TH_LDB = TH_N + 4, TH_THREADS = 256;
constexpr int TH_STAGE_DOUBLES = TH_K * (TH_LDA + TH_LDB);
constexpr int TH_SMEM = TH_STAGES * TH_STAGE_DOUBLES * 8;

// blockIdx.z > 0 or gridDim.z > 1: split-K.  Slice z covers k in [z * kslice, (z + 1) * kslice) and writes its partial product
// (alpha = 1, beta = 0) to part + z * M * N (column-major M x N, ld = M); splitk_reduce_kernel then adds the slices in order.
__global__ void __launch_bounds__(TH_THREADS, 1) dgemm_nt_thin_kernel(GemmArgs g, double *part, int kslice) {
    extern __shared__ __align__(16) double smem[];
    const int m0 = blockIdx.x * TH_M, n0 = blockIdx.y * TH_N;
    if (part) {
        g.A += (long) blockIdx.z * kslice * g.lda;
        g.B += (long) blockIdx.z * kslice * g.ldb;
        g.K = min(kslice, g.K - (int) blockIdx.z * kslice);
        g.C = part + (long) blockIdx.z * g.M * g.N;
        g.ldc = g.M;
        g.alpha = 1.0; g.beta = 0.0;
    }
    const bool lower = (g.flags & HD_GEMM_LOWER) != 0;
    if (lower && n0 > m0 + TH_M - 1) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int wn = warp * 16;
    const int nk = g.K / TH_K;
    const double *Ag = g.A + m0;
    const double *Bg = g.B + n0;
    auto load_stage = [&](int stage, int kt) {
        double *As = smem + stage * TH_STAGE_DOUBLES;
        double *Bs = As + TH_K * TH_LDA;
        const double *a = Ag + (long) kt * TH_K * g.lda;
        const double *b = Bg + (long) kt * TH_K * g.ldb;
#pragma unroll
        for (int i = 0; i < TH_K * (TH_M / 2) / TH_THREADS; ++i) {
            const int c = tid + i * TH_THREADS, kr = c / (TH_M / 2), mc = (c % (TH_M / 2)) * 2;
            cp_async16(As + kr * TH_LDA + mc, a + (long) kr * g.lda + mc);
        }
#pragma unroll
        for (int i = 0; i < TH_K * (TH_N / 2) / TH_THREADS; ++i) {
            const int c = tid + i * TH_THREADS, kr = c / (TH_N / 2), nc = (c % (TH_N / 2)) * 2;
            cp_async16(Bs + kr * TH_LDB + nc, b + (long) kr * g.ldb + nc);
        }
    };
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < TH_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<TH_STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + TH_STAGES - 1;
            if (nxt < nk) load_stage(nxt % TH_STAGES, nxt);
            cp_async_commit();
        }
        const double *As = smem + (kt % TH_STAGES) * TH_STAGE_DOUBLES;
        const double *Bs = As + TH_K * TH_LDA;
#pragma unroll
        for (int kk = 0; kk < TH_K; kk += 4) {
            double bf[2], af[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) bf[i] = Bs[(kk + tig) * TH_LDB + wn + 8 * i + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) af[j] = As[(kk + tig) * TH_LDA + 8 * j + gid];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], bf[i], af[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads(); // in-place products: every warp has finished reading the A rows before any of them is overwritten
    // thread owns C[m .. m+1][n], n = n0 + wn + 8 i + gid, m = m0 + 8 j + 2 tig
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int n = n0 + wn + 8 * i + gid;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + 8 * j + 2 * tig;
            if (lower && m + 1 < n) continue;
            double2 *cp = reinterpret_cast<double2 *>(g.C + (long) n * g.ldc + m);
            double2 v;
            if (g.beta == 0.0) {
                v.x = g.alpha * acc[i][j][0];
                v.y = g.alpha * acc[i][j][1];
            } else {
                const double2 old = *cp;
                v.x = g.alpha * acc[i][j][0] + g.beta * old.x;
                v.y = g.alpha * acc[i][j][1] + g.beta * old.y;
            }
            if (lower && m < n) g.C[(long) n * g.ldc + m + 1] = v.y;
            else *cp = v;
        }
    }
}

// C = alpha * sum_z part_z + beta * C over the (lower part of the) M x N result, slices added in a fixed order (deterministic)
__global__ void splitk_reduce_kernel(const double *__restrict__ part, int nsplit, int M, int N, double alpha, double beta, double *C,
                                     long ldc, int lower) {
    const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long) M * N) return;
    const int m = (int) (idx % M), n = (int) (idx / M);
    if (lower && m < n) return;
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z) s += part[(long) z * M * N + idx];
    double *cp = C + (long) n * ldc + m;
    *cp = (beta == 0.0) ? alpha * s : alpha * s + beta * (*cp);
}

double *g_splitk_ws = nullptr;
size_t g_splitk_bytes = 0;

int g_thin_max_tiles = 96; // products with at most this many 128 x 64 tiles take the thin kernel (0 disables it)

int launch_thin(cudaStream_t st, const GemmArgs &g) {
    static unsigned long long attr = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr >> (dev & 63) & 1ull)) {
        HD_CUDA(cudaFuncSetAttribute(dgemm_nt_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TH_SMEM));
        attr |= 1ull << (dev & 63);
    }
    // long-K products with a small result (S assembly from thousands of dense rank-one rows: 128 x 128 x 20000) would run on
    // a handful of CTAs: cut K into slices so that ~2 CTAs per SM are busy, then add the partial products in slice order
    const long ctas = (long) (g.M / TH_M) * (g.N / TH_N);
    if (g.K >= 4096 && ctas * 4 <= hd_num_sms() && g.A != g.C) {
        int nsplit = (int) ((2L * hd_num_sms() + ctas - 1) / ctas);
        int kslice = ((g.K / nsplit + TH_K - 1) / TH_K) * TH_K;
        if (kslice < 256) kslice = 256;
        nsplit = (g.K + kslice - 1) / kslice;
        const size_t need = sizeof(double) * (size_t) nsplit * g.M * g.N;
        if (need > g_splitk_bytes) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(st, &cap);
            if (cap == cudaStreamCaptureStatusNone) {
                if (g_splitk_ws) cudaFree(g_splitk_ws);
                g_splitk_ws = nullptr; g_splitk_bytes = 0;
                if (cudaMalloc(&g_splitk_ws, need) == cudaSuccess) g_splitk_bytes = need; else cudaGetLastError();
            }
        }
        if (need <= g_splitk_bytes) {
            g_hd_launches += 2;
            dgemm_nt_thin_kernel<<<dim3(g.M / TH_M, g.N / TH_N, nsplit), TH_THREADS, TH_SMEM, st>>>(g, g_splitk_ws, kslice);
            const long total = (long) g.M * g.N;
            splitk_reduce_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, st>>>(g_splitk_ws, nsplit, g.M, g.N, g.alpha, g.beta, g.C, g.ldc,
                                                                                  (g.flags & HD_GEMM_LOWER) ? 1 : 0);
            HD_CUDA(cudaGetLastError());
            return HD_OK;
        }
    }
    ++g_hd_launches;
    dgemm_nt_thin_kernel<<<dim3(g.M / TH_M, g.N / TH_N), TH_THREADS, TH_SMEM, st>>>(g, nullptr, 0);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int g_num_sms = 0;
int g_variant = 4; // 0: 128x128x16 4 stages; 1: 128x128x32 3 stages; 2: 128x64x16 4 stages x2 CTAs; 3: 128x64x32 2 stages x2 CTAs;
                   // 4: as 3 with 8 warps of 32x32 per CTA (16 warps / SM, default); 5: 128x128x32 with 16 warps of 32x32

template <int BN, int BK, int STAGES> constexpr int smem_bytes() { return STAGES * BK * (LDS_ + BN + 4) * 8; }

template <int BN, int BK, int STAGES, int MINB, bool SIGNED = false, int WM = 64, bool PEER = false>
int launch_variant(cudaStream_t st, const GemmArgs &g) {
    static unsigned long long attr = 0; // one bit per device (function attributes are per context)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr >> (dev & 63) & 1ull)) {
        HD_CUDA(cudaFuncSetAttribute(dgemm_nt_kernel<BN, BK, STAGES, MINB, SIGNED, WM, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem_bytes<BN, BK, STAGES>()));
        attr |= 1ull << (dev & 63);
    }
    constexpr int GN = GROUP * (128 / BN);
    TileMap tmap;
    tmap.tiles_m = g.M / BM;
    tmap.tiles_n = g.N / BN;
    tmap.lower = (g.flags & HD_GEMM_LOWER) ? 1 : 0;
    long nsuper;
    if (g.bc_nb > 0) {
        if (g.bc_nb % BM || g.bc_stride % g.bc_nb) return HD_FAILED;
        tmap.lower = 0; // rectangular enumeration of the owned blocks; the kernel drops tiles above the diagonal
    }
    if (tmap.lower) {
        if (g.M != g.N) return HD_FAILED;
        long s = (tmap.tiles_m + GROUP - 1) / GROUP;
        nsuper = s * (s + 1) / 2;
    } else {
        nsuper = (long) ((tmap.tiles_m + GROUP - 1) / GROUP) * ((tmap.tiles_n + GN - 1) / GN);
    }
    long nblocks = nsuper * GROUP * GN;
    ++g_hd_launches;
    dgemm_nt_kernel<BN, BK, STAGES, MINB, SIGNED, WM, PEER><<<(unsigned) nblocks, (128 / WM) * (BN / 32) * 32, smem_bytes<BN, BK, STAGES>(), st>>>(g, tmap);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

} // namespace

void hd_gemm_set_variant(int v) { g_variant = v; }
void hd_gemm_set_thin(int max_tiles) { g_thin_max_tiles = max_tiles; }
int hd_gemm_get_variant() { return g_variant; }

int hd_num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

int hd_gemm_nt(cudaStream_t st, const GemmArgs &g) {
    if (g.M <= 0 || g.N <= 0) return HD_OK;
    if (g.M % BM || g.N % BN || g.K % 32) {
        fprintf(stderr, "[hdsdpcu] gemm_nt: unpadded shape %d %d %d\n", g.M, g.N, g.K);
        return HD_FAILED;
    }
    // thin products (few tiles): 32 x 128 tiles on 4 x as many SMs; plain alpha / beta / LOWER semantics only
    if (g_thin_max_tiles > 0 && !g.peerC && !g.ksign && g.bc_nb == 0 && g.B != g.C && (g.A != g.C || g.N == TH_N) &&
        !(g.flags & (HD_GEMM_KTRI_MAX | HD_GEMM_KTRI_A | HD_GEMM_EPI_HADSQ | HD_GEMM_EPI_COLSCALE))) {
        long tiles = (long) (g.M / BM) * (g.N / 64);
        if (g.flags & HD_GEMM_LOWER) tiles = tiles / 2 + g.M / BM;
        if (tiles <= g_thin_max_tiles) {
            if ((g.flags & HD_GEMM_LOWER) && g.M != g.N) return HD_FAILED;
            return launch_thin(st, g);
        }
    }
    // In-place products (C aliases A, used by the leaf triangular solves with N == 128) are only safe when ONE CTA
    // owns all columns of its row block: the 64-wide tiles would let a sibling CTA overwrite rows still being read.
    if (g.A == g.C || g.B == g.C) {
        if (g.N != BN) return HD_FAILED;
        if (g.peerC) return launch_variant<128, 32, 3, 1, false, 64, true>(st, g);
        return launch_variant<128, 32, 3, 1>(st, g);
    }
    if (g.peerC) return HD_FAILED; // only the in-place leaf products hand their tiles to a peer
    if (g.peerC && (g.beta != 0.0 || (g.flags & (HD_GEMM_LOWER | HD_GEMM_EPI_HADSQ)))) return HD_FAILED;
    if ((g.flags & HD_GEMM_EPI_COLSCALE) && (g.beta != 0.0 || (g.flags & (HD_GEMM_LOWER | HD_GEMM_EPI_HADSQ)))) return HD_FAILED;
    if (g.ksign) return launch_variant<64, 32, 2, 2, true>(st, g); // LDL^T fallback path: one instantiation is enough
    switch (g_variant) {
        case 0: return launch_variant<128, 16, 4, 1>(st, g);
        case 1: return launch_variant<128, 32, 3, 1>(st, g);
        case 2: return launch_variant<64, 16, 4, 2>(st, g);
        case 4: return launch_variant<64, 32, 2, 2, false, 32>(st, g);  // 8 warps of 32x32, 16 warps / SM
        case 5: return launch_variant<128, 32, 2, 1, false, 32>(st, g); // 16 warps of 32x32, one CTA / SM
        case 6: return launch_variant<64, 16, 4, 2, false, 32>(st, g);  // as 4 with BK = 16, 4 stages
        case 7: return launch_variant<64, 16, 3, 2, false, 32>(st, g);  // as 4 with BK = 16, 3 stages
        default: return launch_variant<64, 32, 2, 2>(st, g);
    }
}
