// hdsdp_b200/csrc/lanczos.cu -- dual ratio test on the device.
//
// Reference: coneRatioTest = sdpDenseConeRatioTestImpl (interface/hdsdp_conic_sdp.c:1642-1686): assemble
// dS = dAdaRatio Rd I - A' dy + dTau C into BUFFER_DUALSTEP, then HLanczosSolve (linalg/hdsdp_lanczos.c:161-299) on the
// operator  w -> -L^-1 dS L^-T w  (sdpDenseConeILanczosMultiply, hdsdp_conic_sdp.c:462-505): the largest alpha with
// S + alpha dS >= 0 is 1 / lambda_max.  3..8 ratio tests per IPM iteration, <= 30 Lanczos steps each.
//
// Here every n-vector (the Krylov basis V, w, z1, z2, the warm start) lives in HBM next to L and dS, so a ratio test
// moves O(1) scalars per step across PCIe instead of dragging L back to the host.  The control flow, the start vector
// (srand(n) sequence), the check frequency, the residual tests and the step-size formula are the reference's; the
// (k+1) x (k+1) projected eigenproblem (k <= 30) is solved on the host by cyclic Jacobi (the reference calls dsyevr).
#include "cone.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr int LZ_DIM = 30;       // HLanczosInit(cone->Lanczos, nCol, 30), hdsdp_conic_sdp.c:1393
constexpr int LZ_THREADS = 1024;

__device__ double block_sum(double v, double *red) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// y = -A x with A full symmetric n x n (ld): CTA = 128 rows x 8 column slices
__global__ void __launch_bounds__(1024) neg_symv_kernel(const double *__restrict__ A, long ld, int n, const double *__restrict__ x, double *y) {
    __shared__ double part[8][128];
    const int r = blockIdx.x * 128 + (threadIdx.x & 127), sl = threadIdx.x >> 7;
    double s = 0.0;
    if (r < n) {
        const int chunk = (n + 7) / 8, j0 = sl * chunk, j1 = min(n, j0 + chunk);
#pragma unroll 8
        for (int j = j0; j < j1; ++j) s += A[(long) j * ld + r] * x[j];
    }
    part[sl][threadIdx.x & 127] = s;
    __syncthreads();
    if (threadIdx.x < 128 && r < n) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += part[q][threadIdx.x];
        y[r] = -t;
    }
}

// one Lanczos step after w = M v_k (hdsdp_lanczos.c:199-221); scal[0] = vAlp, scal[1] = normPres
__global__ void __launch_bounds__(LZ_THREADS) lz_step_kernel(int n, int k, double *V, long ldv, double *w, double *v, double *H, int nH,
                                                            double *scal) {
    __shared__ double red[33];
    const double hprev = (k > 0) ? H[(long) nH * (k - 1) + k] : 0.0;
    const double *vk = V + (long) k * ldv, *vkm = V + (long) (k > 0 ? k - 1 : 0) * ldv;
    double d = 0.0;
    for (int i = threadIdx.x; i < n; i += LZ_THREADS) {
        double wi = w[i];
        if (k > 0) wi -= hprev * vkm[i];
        w[i] = wi;
        d += wi * vk[i];
    }
    const double alp = -block_sum(d, red);
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += LZ_THREADS) {
        const double wi = w[i] + alp * vk[i];
        w[i] = wi;
        q += wi * wi;
    }
    const double nrm = sqrt(block_sum(q, red));
    double *vn = V + (long) (k + 1) * ldv;
    for (int i = threadIdx.x; i < n; i += LZ_THREADS) {
        const double vi = (nrm > 0.0) ? w[i] / nrm : 0.0;
        v[i] = vi;
        if (nrm > 0.0) vn[i] = vi;
    }
    if (threadIdx.x == 0) {
        H[(long) nH * k + k] = -alp;
        if (nrm > 0.0) { H[(long) nH * k + k + 1] = nrm; H[(long) nH * (k + 1) + k] = nrm; }
        scal[0] = alp; scal[1] = nrm;
    }
}

// z = V[:, 0:kp] * yv  (fds_gemv, :246/:257)
__global__ void lz_combine_kernel(int n, int kp, const double *__restrict__ V, long ldv, const double *__restrict__ yv, double *z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int j = 0; j < kp; ++j) s += V[(long) j * ldv + i] * yv[j];
    z[i] = s;
}

// b += a * x ; out = ||b||   (axpy + nrm2, :254-256 / :259-263); optionally keep a copy of b before the update
__global__ void __launch_bounds__(LZ_THREADS) lz_axpy_nrm_kernel(int n, double a, const double *__restrict__ x, double *b, double *keep, double *out) {
    __shared__ double red[33];
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += LZ_THREADS) {
        double bi = b[i];
        if (keep) keep[i] = bi;
        bi += a * x[i];
        b[i] = bi;
        q += bi * bi;
    }
    const double s = block_sum(q, red);
    if (threadIdx.x == 0) *out = sqrt(s);
}

// v = (base + pert) / ||base + pert|| (HLanczosIPerturb + normalize, :179-187), also written to V[:, 0]
__global__ void __launch_bounds__(LZ_THREADS) lz_start_kernel(int n, const double *base, const double *__restrict__ pert, double *v, double *V0) {
    __shared__ double red[33];
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += LZ_THREADS) {
        const double t = (base ? base[i] : 0.0) + pert[i];
        v[i] = t;
        q += t * t;
    }
    const double nrm = sqrt(block_sum(q, red));
    for (int i = threadIdx.x; i < n; i += LZ_THREADS) {
        const double t = (nrm > 0.0) ? v[i] / nrm : 0.0;
        v[i] = t; V0[i] = t;
    }
}

// eigen-decomposition of a small symmetric matrix (cyclic Jacobi); eigenvalues ascending, vectors in columns of Q
void jacobi_eig(int n, std::vector<double> &A, std::vector<double> &evals, std::vector<double> &Q) {
    Q.assign((size_t) n * n, 0.0);
    for (int i = 0; i < n; ++i) Q[(size_t) i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += A[(size_t) q * n + p] * A[(size_t) q * n + p];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[(size_t) q * n + p];
                if (apq == 0.0) continue;
                const double app = A[(size_t) p * n + p], aqq = A[(size_t) q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[(size_t) p * n + k], akq = A[(size_t) q * n + k];
                    A[(size_t) p * n + k] = c * akp - s * akq;
                    A[(size_t) q * n + k] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[(size_t) k * n + p], aqk = A[(size_t) k * n + q];
                    A[(size_t) k * n + p] = c * apk - s * aqk;
                    A[(size_t) k * n + q] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double qkp = Q[(size_t) p * n + k], qkq = Q[(size_t) q * n + k];
                    Q[(size_t) p * n + k] = c * qkp - s * qkq;
                    Q[(size_t) q * n + k] = s * qkp + c * qkq;
                }
            }
    }
    std::vector<int> ord(n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return A[(size_t) a * n + a] < A[(size_t) b * n + b]; });
    evals.resize(n);
    std::vector<double> Qs((size_t) n * n);
    for (int i = 0; i < n; ++i) {
        evals[i] = A[(size_t) ord[i] * n + ord[i]];
        for (int k = 0; k < n; ++k) Qs[(size_t) i * n + k] = Q[(size_t) ord[i] * n + k];
    }
    Q.swap(Qs);
}

// the reference's start / perturbation vectors (hdsdp_lanczos.c:33-55), same libc rand() sequence
void start_vector(int n, double scale, std::vector<double> &v) {
    v.resize(n);
    srand((unsigned int) n);
    for (int i = 0; i < n; ++i) {
        srand((unsigned int) rand());
        v[i] = scale * sqrt(sqrt((double) (rand() % 1627))) * (rand() % 2 - 0.5);
    }
}

} // namespace

struct LanczosCU {
    int n = 0, np = 0;
    double *V = nullptr;      // np x (LZ_DIM + 1)
    double *H = nullptr;      // (LZ_DIM + 1)^2, device
    double *v = nullptr, *w = nullptr, *z1 = nullptr, *z2 = nullptr, *warm = nullptr, *tmp = nullptr, *pert = nullptr;
    double *yv = nullptr;     // small device vector (eigenvector of the projected problem)
    double *scal = nullptr;   // device scalars
    double *symv_ws = nullptr; // partial vectors of the ordered symv reduction
    double *h_scal = nullptr; // pinned
    double *h_H = nullptr;    // pinned
    int nComputed = 0;
    int lastSteps = 0;
};

static int lz_create(LanczosCU **pl, int n, int np) {
    LanczosCU *l = new LanczosCU();
    l->n = n; l->np = np;
    HD_CUDA(cudaMalloc(&l->V, sizeof(double) * (size_t) np * (LZ_DIM + 1)));
    HD_CUDA(cudaMalloc(&l->H, sizeof(double) * (LZ_DIM + 1) * (LZ_DIM + 1)));
    double **vecs[] = {&l->v, &l->w, &l->z1, &l->z2, &l->warm, &l->tmp, &l->pert};
    for (double **p : vecs) {
        HD_CUDA(cudaMalloc(p, sizeof(double) * np));
        HD_CUDA(cudaMemset(*p, 0, sizeof(double) * np));
    }
    HD_CUDA(cudaMalloc(&l->symv_ws, sizeof(double) * (size_t) hd_symv_ws_doubles(np)));
    HD_CUDA(cudaMalloc(&l->yv, sizeof(double) * (LZ_DIM + 1)));
    HD_CUDA(cudaMalloc(&l->scal, sizeof(double) * 8));
    HD_CUDA(cudaMallocHost(&l->h_scal, sizeof(double) * 8));
    HD_CUDA(cudaMallocHost(&l->h_H, sizeof(double) * (LZ_DIM + 1) * (LZ_DIM + 1)));
    *pl = l;
    return HD_OK;
}

void lz_destroy(LanczosCU *l) {
    if (!l) return;
    cudaFree(l->V); cudaFree(l->H); cudaFree(l->v); cudaFree(l->w); cudaFree(l->z1); cudaFree(l->z2); cudaFree(l->warm);
    cudaFree(l->tmp); cudaFree(l->pert); cudaFree(l->yv); cudaFree(l->scal); cudaFree(l->symv_ws);
    cudaFreeHost(l->h_scal); cudaFreeHost(l->h_H);
    delete l;
}

// out = -L^-1 dS L^-T in  (device vectors of length np, padding zero); dS must be full symmetric in d_buf[DUALSTEP]
static int lz_matvec(ConeCU *c, DenseChol *f, cudaStream_t st, const double *in, double *out, double *tmp) {
    HD_CUDA(cudaMemcpyAsync(tmp, in, sizeof(double) * c->np, cudaMemcpyDeviceToDevice, st));
    HD_CALL(chol_bsolve(st, f, tmp, 1, c->np));                                   // L^T x = w
    // y = -dS x from the lower triangle of dS (read once, two CTAs per SM): the full-matrix kernel below ran on n / 128 CTAs only
    HD_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * c->np, st));
    HD_CALL(hd_symv_lower(st, c->d_buf[BUF_DUALSTEP], c->np, c->np, tmp, c->np, out, c->np, 1, -1.0, c->lanczos ? c->lanczos->symv_ws : nullptr));
    return chol_fsolve(st, f, out, 1, c->np);                                     // L z = y
}

int cone_lanczos_multiply(ConeCU *c, int which, const double *x, double *y) {
    cudaStream_t st = hd_stream();
    DenseChol *f = (which == BUF_DUALVAR) ? c->factor : c->checker;
    if (!f->factored) return HD_FAILED;
    if (!c->lanczos) HD_CALL(lz_create(&c->lanczos, c->n, c->np));
    LanczosCU *l = c->lanczos;
    HD_CALL(hd_symmetrize_lower(st, c->d_buf[BUF_DUALSTEP], c->np, c->np));
    HD_CUDA(cudaMemsetAsync(l->z1, 0, sizeof(double) * c->np, st));
    HD_CUDA(cudaMemcpyAsync(l->z1, x, sizeof(double) * c->n, cudaMemcpyHostToDevice, st));
    HD_CALL(lz_matvec(c, f, st, l->z1, l->z2, l->tmp));
    HD_CUDA(cudaMemcpyAsync(y, l->z2, sizeof(double) * c->n, cudaMemcpyDeviceToHost, st));
    HD_CUDA(cudaStreamSynchronize(st));
    return HD_OK;
}

int cone_ratio_test(ConeCU *c, double dTauStep, const double *dyHost, double dAdaRatio, int which, double *maxStep) {
    cudaStream_t st = hd_stream();
    const int n = c->n, np = c->np;
    DenseChol *f = (which == BUF_DUALVAR) ? c->factor : c->checker;
    if (!f->factored) return HD_FAILED;
    // dS = dAdaRatio * Rd * I - A' dy + dTau * C   (hdsdp_conic_sdp.c:1663)
    HD_CALL(cone_update_buffer(c, dTauStep, -1.0, dyHost, nullptr, dAdaRatio * c->dualResidual, BUF_DUALSTEP));
    if (n == 1) { // :1674-1680
        double s = 0.0, ds = 0.0;
        HD_CUDA(cudaMemcpyAsync(&s, c->d_buf[which], sizeof(double), cudaMemcpyDeviceToHost, st));
        HD_CUDA(cudaMemcpyAsync(&ds, c->d_buf[BUF_DUALSTEP], sizeof(double), cudaMemcpyDeviceToHost, st));
        HD_CUDA(cudaStreamSynchronize(st));
        *maxStep = (ds > 0.0) ? INFINITY : -s / ds;
        return HD_OK;
    }
    HD_CALL(hd_symmetrize_lower(st, c->d_buf[BUF_DUALSTEP], np, np));
    if (!c->lanczos) HD_CALL(lz_create(&c->lanczos, n, np));
    LanczosCU *l = c->lanczos;
    const int nH = LZ_DIM + 1;
    // ---- start vector (:164-187) ---------------------------------------------------------------------------------
    std::vector<double> hv;
    start_vector(n, l->nComputed == 0 ? 1.0 : 1e-03, hv);
    HD_CUDA(cudaMemcpyAsync(l->pert, hv.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    HD_CUDA(cudaMemsetAsync(l->V, 0, sizeof(double) * (size_t) np * (LZ_DIM + 1), st));
    HD_CUDA(cudaMemsetAsync(l->H, 0, sizeof(double) * nH * nH, st));
    HDK(lz_start_kernel)<<<1, LZ_THREADS, 0, st>>>(n, l->nComputed == 0 ? nullptr : l->warm, l->pert, l->v, l->V);
    HD_CUDA(cudaStreamSynchronize(st)); // hv is a stack-owned staging buffer
    int checkFreq = LZ_DIM / 5;
    if (checkFreq > 3) checkFreq = 3;
    int rc = HD_OK;
    int k = 0;
    for (k = 0; k < LZ_DIM; ++k) {
        HD_CALL(lz_matvec(c, f, st, l->v, l->w, l->tmp));
        HDK(lz_step_kernel)<<<1, LZ_THREADS, 0, st>>>(n, k, l->V, np, l->w, l->v, l->H, nH, l->scal);
        HD_CUDA(cudaMemcpyAsync(l->h_scal, l->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        HD_CUDA(cudaStreamSynchronize(st));
        const double normPres = l->h_scal[1];
        if ((k + 1) % checkFreq == 0 || k > LZ_DIM - 1 || normPres == 0.0) {
            const int kp = k + 1;
            HD_CUDA(cudaMemcpyAsync(l->h_H, l->H, sizeof(double) * nH * nH, cudaMemcpyDeviceToHost, st));
            HD_CUDA(cudaStreamSynchronize(st));
            std::vector<double> U((size_t) kp * kp), ev, Q;
            for (int i = 0; i < kp; ++i)
                for (int j = 0; j < kp; ++j) U[(size_t) i * kp + j] = 0.5 * (l->h_H[(size_t) nH * i + j] + l->h_H[(size_t) nH * j + i]);
            jacobi_eig(kp, U, ev, Q);
            const double *yMax = &Q[(size_t) (kp - 1) * kp];                       // largest eigenvalue (YMat + kPlus1)
            const double *ySec = &Q[(size_t) (kp >= 2 ? kp - 2 : 0) * kp];         // second largest (YMat)
            const double resiVal = fabs(l->h_H[(size_t) nH * k + kp] * yMax[k]);   // |H(k+1, k) * y_k|
            if (resiVal < 1e-04 || k >= LZ_DIM - 1) {
                const double eigMin1 = ev[kp - 1], eigMin2 = (kp >= 2) ? ev[kp - 2] : ev[kp - 1];
                // z1 = V yMax ; z2 = M z1 ; warm = z2 ; z2 -= eig z1 ; resiVal1 = ||z2||
                HD_CUDA(cudaMemcpyAsync(l->yv, yMax, sizeof(double) * kp, cudaMemcpyHostToDevice, st));
                HDK(lz_combine_kernel)<<<(n + 255) / 256, 256, 0, st>>>(n, kp, l->V, np, l->yv, l->z1);
                HD_CUDA(cudaStreamSynchronize(st));
                HD_CALL(lz_matvec(c, f, st, l->z1, l->z2, l->tmp));
                HDK(lz_axpy_nrm_kernel)<<<1, LZ_THREADS, 0, st>>>(n, -eigMin1, l->z1, l->z2, l->warm, l->scal + 2);
                // z2 = V ySec ; z1 = M z2 ; z1 -= eig z2 ; resiVal2 = ||z1||
                HD_CUDA(cudaMemcpyAsync(l->yv, ySec, sizeof(double) * kp, cudaMemcpyHostToDevice, st));
                HDK(lz_combine_kernel)<<<(n + 255) / 256, 256, 0, st>>>(n, kp, l->V, np, l->yv, l->z2);
                HD_CUDA(cudaStreamSynchronize(st));
                HD_CALL(lz_matvec(c, f, st, l->z2, l->z1, l->tmp));
                HDK(lz_axpy_nrm_kernel)<<<1, LZ_THREADS, 0, st>>>(n, -eigMin1, l->z2, l->z1, nullptr, l->scal + 3);
                HD_CUDA(cudaMemcpyAsync(l->h_scal + 2, l->scal + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
                HD_CUDA(cudaStreamSynchronize(st));
                const double resiVal1 = l->h_scal[2], resiVal2 = l->h_scal[3];
                const double resiDiff = eigMin1 - eigMin2 - resiVal2;
                double valGamma = (resiDiff > 0) ? resiDiff : 1e-16;
                const double resiVal1sqr = resiVal1 * resiVal1 / valGamma;
                valGamma = resiVal1 < resiVal1sqr ? resiVal1 : resiVal1sqr;
                if (valGamma < 1e-03 || valGamma + eigMin1 <= 0.5) {
                    *maxStep = (valGamma + eigMin1 <= 0.0) ? INFINITY : 1.0 / (valGamma + eigMin1);
                    break;
                } else {
                    if (normPres == 0.0) { rc = HD_FAILED; break; }
                    *maxStep = 1.0 / (valGamma + eigMin1);
                }
            }
        }
    }
    l->lastSteps = k + 1;
    if (rc == HD_OK) l->nComputed += 1;
    return rc;
}

int cone_lanczos_steps(ConeCU *c) { return c->lanczos ? c->lanczos->lastSteps : 0; }


// ------------------------------------------------------------------------------------------------------------------
// Extreme eigenvalue of a symmetric matrix (the DIMACS check of HDSDPCheckSolution, interface/hdsdp.c:852-861,
// runs dsyevr on every primal block X -- O(n^3) on the host -- to get ONE eigenvalue: fds_syev(n, X, d, Y, 1, ...) asks for
// index n, i.e. the LARGEST one, which the reference then records as "dMinPrimalEVal").  Here: Lanczos on the device
// (symv + the step kernel above), Ritz values of the tridiagonal by Jacobi on the host every 10 steps, stop when the
// residual bound |beta_k y_k| of the wanted Ritz pair is below 1e-12 ||X||.  which = 1: largest, which = 0: smallest.
// ------------------------------------------------------------------------------------------------------------------
int sym_extreme_eig(int n, const double *Xhost, int which, double *out, int *steps) {
    cudaStream_t st = hd_stream();
    if (n <= 0) return HD_FAILED;
    if (n == 1) { *out = Xhost[0]; if (steps) *steps = 0; return HD_OK; }
    const int maxit = n < 300 ? n : 300;
    const int nH = maxit + 1;
    double *X = nullptr, *V = nullptr, *H = nullptr, *v = nullptr, *w = nullptr, *scal = nullptr;
    HD_CUDA(cudaMalloc(&X, sizeof(double) * (size_t) n * n));
    HD_CUDA(cudaMalloc(&V, sizeof(double) * (size_t) n * (maxit + 1)));
    HD_CUDA(cudaMalloc(&H, sizeof(double) * (size_t) nH * nH));
    HD_CUDA(cudaMalloc(&v, sizeof(double) * n)); HD_CUDA(cudaMalloc(&w, sizeof(double) * n)); HD_CUDA(cudaMalloc(&scal, sizeof(double) * 8));
    HD_CUDA(cudaMemcpyAsync(X, Xhost, sizeof(double) * (size_t) n * n, cudaMemcpyHostToDevice, st));
    HD_CUDA(cudaMemsetAsync(H, 0, sizeof(double) * (size_t) nH * nH, st));
    std::vector<double> hv;
    start_vector(n, 1.0, hv);
    HD_CUDA(cudaMemcpyAsync(w, hv.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    HDK(lz_start_kernel)<<<1, LZ_THREADS, 0, st>>>(n, nullptr, w, v, V);
    std::vector<double> hH((size_t) nH * nH), hs(2);
    int rc = HD_OK, k = 0;
    double best = 0.0, scale = 0.0;
    for (k = 0; k < maxit; ++k) {
        HDK(neg_symv_kernel)<<<(n + 127) / 128, 1024, 0, st>>>(X, n, n, v, w);             // w = -X v_k
        HDK(lz_step_kernel)<<<1, LZ_THREADS, 0, st>>>(n, k, V, n, w, v, H, nH, scal);
        if ((k + 1) % 10 != 0 && k + 1 != maxit) continue;
        if (cudaMemcpyAsync(hH.data(), H, sizeof(double) * (size_t) nH * nH, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { rc = HD_FAILED; break; }
        const int kp = k + 1;
        std::vector<double> T((size_t) kp * kp), ev, Q;
        for (int i = 0; i < kp; ++i)
            for (int j = 0; j < kp; ++j) T[(size_t) i * kp + j] = 0.5 * (hH[(size_t) nH * i + j] + hH[(size_t) nH * j + i]);
        jacobi_eig(kp, T, ev, Q);
        // operator is -X: lambda_max(X) = -theta_min, lambda_min(X) = -theta_max
        const int idx = which ? 0 : kp - 1;
        best = -ev[idx];
        scale = fmax(fabs(ev[0]), fabs(ev[kp - 1]));
        const double beta = hH[(size_t) nH * k + kp];                                       // H(k+1, k)
        const double resid = fabs(beta * Q[(size_t) idx * kp + k]);
        if (resid <= 1e-12 * fmax(scale, 1e-300) || beta == 0.0) { ++k; break; }
    }
    if (steps) *steps = k;
    *out = best;
    cudaFree(X); cudaFree(V); cudaFree(H); cudaFree(v); cudaFree(w); cudaFree(scal);
    return rc;
}
