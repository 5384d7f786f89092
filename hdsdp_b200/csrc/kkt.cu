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

// LP cone: for LP column c with rows r_0..r_k (CSC by LP column) and weight w_c = s_c^-2:
//   M[r_a, r_b] += a_a a_b w_c (r_a >= r_b);  asinv[r] += a / s_c ; asinvrd[r] += rd * a * w_c
__global__ void lp_schur_kernel(const int *__restrict__ colptr, const int *__restrict__ rowidx, const double *__restrict__ val,
                                const double *__restrict__ sinv, int ncol, double rd, int build_matrix, double *M, long ldm,
                                double *asinv, double *asinvrd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const double si = sinv[c], w = si * si;
    for (int a = colptr[c]; a < colptr[c + 1]; ++a) {
        const int ra = rowidx[a];
        const double va = val[a];
        atomicAdd(&asinv[ra], va * si);
        if (rd != 0.0) atomicAdd(&asinvrd[ra], rd * va * w);
        if (!build_matrix) continue;
        for (int b = colptr[c]; b <= a; ++b) {
            const int rb = rowidx[b];
            const int r = max(ra, rb), q = min(ra, rb);
            atomicAdd(&M[(long) q * ldm + r], va * val[b] * w);
        }
    }
}

inline unsigned nblk(long total, int threads) { return (unsigned) ((total + threads - 1) / threads); }

} // namespace

int hd_scale_vec(cudaStream_t st, double *x, int m, double a) {
    HDK(scale_vec_kernel)<<<nblk(m, 256), 256, 0, st>>>(x, m, a);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int kkt_create(KktCU **pk, int nRow) {
    KktCU *k = new KktCU();
    k->m = nRow;
    k->mp = hd_pad(nRow);
    size_t bytes = sizeof(double) * (size_t) k->mp * k->mp;
    if (cudaMalloc(&k->d_M, bytes) != cudaSuccess) { cudaGetLastError(); delete k; return HD_MEMORY; }
    HD_CUDA(cudaMemset(k->d_M, 0, bytes));
    int rc = chol_create(&k->chol, nRow);
    if (rc != HD_OK) { cudaFree(k->d_M); delete k; return rc; }
    HD_CUDA(cudaMalloc(&k->d_asinv, sizeof(double) * k->mp));
    HD_CUDA(cudaMalloc(&k->d_asinvrd, sizeof(double) * k->mp));
    HD_CUDA(cudaMalloc(&k->d_asinvc, sizeof(double) * k->mp));
    HD_CUDA(cudaMemset(k->d_asinv, 0, sizeof(double) * k->mp));
    HD_CUDA(cudaMemset(k->d_asinvrd, 0, sizeof(double) * k->mp));
    HD_CUDA(cudaMemset(k->d_asinvc, 0, sizeof(double) * k->mp));
    HD_CUDA(cudaMalloc(&k->d_scal, sizeof(double) * 8));
    HD_CUDA(cudaMemset(k->d_scal, 0, sizeof(double) * 8));
    HD_CUDA(cudaMallocHost(&k->h_scal, sizeof(double) * 8));
    HD_CUDA(cudaMalloc(&k->d_rhs, sizeof(double) * (size_t) k->mp * 8));
    HD_CUDA(cudaMemset(k->d_rhs, 0, sizeof(double) * (size_t) k->mp * 8));
    HD_CUDA(cudaMallocHost(&k->h_vec, sizeof(double) * (size_t) k->mp * 8));
    *pk = k;
    return HD_OK;
}

void kkt_destroy(KktCU *k) {
    if (!k) return;
    cudaFree(k->d_M); cudaFree(k->d_asinv); cudaFree(k->d_asinvrd); cudaFree(k->d_asinvc);
    cudaFree(k->d_scal); cudaFree(k->d_rhs);
    cudaFreeHost(k->h_scal); cudaFreeHost(k->h_vec);
    if (k->dist) dist_destroy(k->dist);
    if (k->d_gather) cudaFree(k->d_gather);
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
int kkt_add_lp(KktCU *k, int nLpCol, const int *d_colptr, const int *d_rowidx, const double *d_val, const double *sInvHost,
               double *d_sinv_stage, double rd, int typeKKT) {
    cudaStream_t st = hd_stream();
    k->fresh = false;
    HD_CUDA(cudaMemcpyAsync(d_sinv_stage, sInvHost, sizeof(double) * nLpCol, cudaMemcpyHostToDevice, st));
    HDK(lp_schur_kernel)<<<nblk(nLpCol, 128), 128, 0, st>>>(d_colptr, d_rowidx, d_val, d_sinv_stage, nLpCol, rd,
                                                        typeKKT != KKT_CORRECTOR, k->d_M, k->mp, k->d_asinv, k->d_asinvrd);
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

// nRhs device vectors of stride mp (padded with zeros) solved in place
int kkt_solve_dev(KktCU *k, double *d_x, int nRhs) {
    if (!k->factored) return HD_FAILED;
    cudaStream_t st = hd_stream();
    HD_CALL(chol_fsolve(st, k->chol, d_x, nRhs, k->mp));
    HD_CALL(chol_dsolve(st, k->chol, d_x, nRhs, k->mp));
    HD_CALL(chol_bsolve(st, k->chol, d_x, nRhs, k->mp));
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
        HD_CALL(kkt_solve_dev(k, k->d_rhs, nb));
        HD_CUDA(cudaMemcpyAsync(k->h_vec, k->d_rhs, sizeof(double) * (size_t) k->mp * nb, cudaMemcpyDeviceToHost, st));
        HD_CUDA(cudaStreamSynchronize(st));
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
