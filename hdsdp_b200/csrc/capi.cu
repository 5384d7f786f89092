// hdsdp_b200/csrc/capi.cu -- extern "C" boundary of libhdsdp_cuda.so (declared in include/hdsdpcu.h).
// Thin: argument checks, host<->device staging, and dispatch into chol.cu / cone.cu / kkt.cu.
#include "../../include/hdsdpcu.h"
#include "cone.h"
#include <cstring>
#include <vector>

long g_hd_launches = 0;

namespace {
cudaStream_t g_stream = nullptr;
bool g_ready = false;

struct LinsysCU {
    DenseChol *c;
    double *d_vec;  // np x 8 staging for solves
    double *d_inv;  // np x np, lazily
};

struct LpCU {
    int m, ncol;
    int *d_colptr, *d_rowidx;
    double *d_val, *d_sinv, *d_obj;
};

int ensure_ready() {
    if (g_ready) return HD_OK;
    return hdsdpcu_init(-1);
}
} // namespace

cudaStream_t hd_stream() { return g_stream; }

extern "C" {

int hdsdpcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hdsdpcu_init(int device) {
    if (hdsdpcu_device_count() <= 0) {
        fprintf(stderr, "[hdsdpcu] no CUDA device: the hot path has no CPU fallback\n");
        return HD_FAILED;
    }
    // one device per process (one process per GPU is the multi-GPU model): the library stream, the side
    // stream of the factorisations and their events belong to the first device selected
    static int g_device = -1;
    if (device >= 0) {
        if (g_device >= 0 && device != g_device) {
            fprintf(stderr, "[hdsdpcu] hdsdpcu_init(%d): this process is already bound to device %d (one process per GPU)\n", device, g_device);
            return HD_FAILED;
        }
        HD_CUDA(cudaSetDevice(device));
    }
    if (g_device < 0) HD_CUDA(cudaGetDevice(&g_device));
    if (!g_stream) HD_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    if (!g_ready) {
        // measurement knobs can also come from the environment (HDSDPCU_GEMM_VARIANT, HDSDPCU_CHOL_BLOCK, HDSDPCU_CHOL_LEAF)
        if (const char *e = getenv("HDSDPCU_GEMM_VARIANT")) hd_gemm_set_variant(atoi(e));
        if (const char *e = getenv("HDSDPCU_CHOL_BLOCK")) hd_chol_set_block(atoi(e));
        if (const char *e = getenv("HDSDPCU_CHOL_LEAF")) hd_chol_set_leaf(atoi(e));
        if (const char *e = getenv("HDSDPCU_CHOL_GRAPH")) hd_chol_set_graph(atoi(e));
        if (const char *e = getenv("HDSDPCU_DIST_DELAY")) hd_dist_set_delay(atoi(e));
        if (const char *e = getenv("HDSDPCU_INVERT_FORK")) hd_chol_set_invert_fork(atoi(e));
        if (const char *e = getenv("HDSDPCU_TRSV_VERSION")) hd_trsv_set_version(atoi(e));
        if (const char *e = getenv("HDSDPCU_CHOL_TAIL")) hd_chol_set_tail(atoi(e));
    }
    g_ready = true;
    return HD_OK;
}

void *hdsdpcu_stream(void) { return (void *) g_stream; }
int hdsdpcu_sync(void) {
    if (!g_ready) return HD_FAILED;
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}
const char *hdsdpcu_version(void) { return "hdsdp-b200 0.1 (sm_100a)"; }
long hdsdpcu_launch_count(int reset) {
    long v = g_hd_launches;
    if (reset) g_hd_launches = 0;
    return v;
}

// Device-time bracket for the host shims (integration/): two events on the library stream around one call of the hot path.
// The side stream of the factorisations forks from and joins the library stream, so the bracket covers it.
static cudaEvent_t g_tm_ev[2] = {nullptr, nullptr};
int hdsdpcu_timer_start(void) {
    HD_CALL(ensure_ready());
    if (!g_tm_ev[0]) { HD_CUDA(cudaEventCreate(&g_tm_ev[0])); HD_CUDA(cudaEventCreate(&g_tm_ev[1])); }
    HD_CUDA(cudaEventRecord(g_tm_ev[0], g_stream));
    return HD_OK;
}
int hdsdpcu_timer_stop(double *ms) {
    if (!g_tm_ev[0]) return HD_FAILED;
    HD_CUDA(cudaEventRecord(g_tm_ev[1], g_stream));
    HD_CUDA(cudaEventSynchronize(g_tm_ev[1]));
    float f = 0.f;
    HD_CUDA(cudaEventElapsedTime(&f, g_tm_ev[0], g_tm_ev[1]));
    if (ms) *ms = (double) f;
    return HD_OK;
}

int hdsdpcu_set_option(const char *name, int value) {
    if (name && strcmp(name, "gemm_variant") == 0) { hd_gemm_set_variant(value); return HD_OK; }
    if (name && strcmp(name, "gemm_thin") == 0) { hd_gemm_set_thin(value); return HD_OK; }
    if (name && strcmp(name, "chol_block") == 0) { hd_chol_set_block(value); return HD_OK; }
    if (name && strcmp(name, "dist_delay") == 0) { hd_dist_set_delay(value); return HD_OK; }
    if (name && strcmp(name, "chol_sched") == 0) { hd_chol_set_sched(value); return HD_OK; }
    if (name && strcmp(name, "chol_leaf") == 0) { hd_chol_set_leaf(value); return HD_OK; }
    if (name && strcmp(name, "ldl_pivot") == 0) { hd_chol_set_ldl_pivot(value); return HD_OK; }
    if (name && strcmp(name, "invert_fork") == 0) { hd_chol_set_invert_fork(value); return HD_OK; }
    if (name && strcmp(name, "chol_tail") == 0) { hd_chol_set_tail(value); return HD_OK; }
    if (name && strcmp(name, "chol_partition") == 0) { hd_chol_set_partition(value); return HD_OK; }
    if (name && strcmp(name, "trsv_version") == 0) { hd_trsv_set_version(value); return HD_OK; }
    if (name && strcmp(name, "chol_graph") == 0) { hd_chol_set_graph(value); return HD_OK; }
    return HD_FAILED;
}

int hdsdpcu_debug_leafclk(long long *out) { return hd_leaf_clocks(out); }

int hdsdpcu_copy_dev(void *d_dst, const void *d_src, long bytes) {
    HD_CALL(ensure_ready());
    HD_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t) bytes, cudaMemcpyDeviceToDevice, g_stream));
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------
// B1: dense linear system back-end
// ------------------------------------------------------------------------------------------------
int hdsdpcu_linsys_create(void **pchol, int nCol) {
    if (!pchol) return HD_FAILED;
    HD_CALL(ensure_ready());
    LinsysCU *l = (LinsysCU *) calloc(1, sizeof(LinsysCU));
    if (!l) return HD_MEMORY;
    int rc = chol_create(&l->c, nCol);
    if (rc != HD_OK) { free(l); return rc; }
    if (cudaMalloc(&l->d_vec, sizeof(double) * (size_t) l->c->np * 8) != cudaSuccess) {
        cudaGetLastError(); chol_destroy(l->c); free(l); return HD_MEMORY;
    }
    *pchol = l;
    return HD_OK;
}

void hdsdpcu_linsys_setparam(void *chol, void *param) { (void) chol; (void) param; }
int hdsdpcu_linsys_symbolic(void *chol, int *colBeg, int *colIdx) { (void) chol; (void) colBeg; (void) colIdx; return HD_OK; }

static int linsys_load_host(LinsysCU *l, const double *elem) {
    DenseChol *c = l->c;
    HD_CALL(chol_ensure_work(c));
    HD_CALL(hd_h2d_matrix(g_stream, c->L, c->np, elem, c->n, c->work));
    HD_CALL(hd_pad_identity(g_stream, c->L, c->np, c->n, c->np));
    c->factored = false;
    return HD_OK;
}

int hdsdpcu_linsys_numeric(void *chol, int *colBeg, int *colIdx, double *elem) {
    (void) colBeg; (void) colIdx;
    LinsysCU *l = (LinsysCU *) chol;
    HD_CALL(linsys_load_host(l, elem));
    int info = 0;
    HD_CALL(chol_factor(g_stream, l->c, &info));
    return info == 0 ? HD_OK : HD_FAILED;
}

int hdsdpcu_linsys_psdcheck(void *chol, int *colBeg, int *colIdx, double *elem, int *isPsd) {
    (void) colBeg; (void) colIdx;
    LinsysCU *l = (LinsysCU *) chol;
    HD_CALL(linsys_load_host(l, elem));
    int info = 0;
    HD_CALL(chol_factor(g_stream, l->c, &info));
    *isPsd = (info == 0);
    return HD_OK;
}

// mode 0: L \ rhs, 1: L' \ rhs, 2: both
static int linsys_solve_host(LinsysCU *l, int nRhs, double *rhs, double *sol, int mode) {
    DenseChol *c = l->c;
    if (!c->factored) return HD_FAILED;
    const int n = c->n, np = c->np;
    double *out = sol ? sol : rhs;
    for (int r0 = 0; r0 < nRhs; r0 += 8) {
        int nb = (nRhs - r0 < 8) ? nRhs - r0 : 8;
        HD_CUDA(cudaMemsetAsync(l->d_vec, 0, sizeof(double) * (size_t) np * nb, g_stream));
        HD_CUDA(cudaMemcpy2DAsync(l->d_vec, (size_t) np * 8, rhs + (size_t) r0 * n, (size_t) n * 8, (size_t) n * 8, nb,
                                  cudaMemcpyHostToDevice, g_stream));
        if (mode == 0 || mode == 2) HD_CALL(chol_fsolve(g_stream, c, l->d_vec, nb, np));
        if (mode == 2) HD_CALL(chol_dsolve(g_stream, c, l->d_vec, nb, np));
        if (mode == 1 || mode == 2) HD_CALL(chol_bsolve(g_stream, c, l->d_vec, nb, np));
        HD_CUDA(cudaMemcpy2DAsync(out + (size_t) r0 * n, (size_t) n * 8, l->d_vec, (size_t) np * 8, (size_t) n * 8, nb,
                                  cudaMemcpyDeviceToHost, g_stream));
        HD_CUDA(cudaStreamSynchronize(g_stream));
    }
    return HD_OK;
}

void hdsdpcu_linsys_fsolve(void *chol, int nRhs, double *rhs, double *sol) { linsys_solve_host((LinsysCU *) chol, nRhs, rhs, sol, 0); }
void hdsdpcu_linsys_bsolve(void *chol, int nRhs, double *rhs, double *sol) { linsys_solve_host((LinsysCU *) chol, nRhs, rhs, sol, 1); }
int hdsdpcu_linsys_solve(void *chol, int nRhs, double *rhs, double *sol) { return linsys_solve_host((LinsysCU *) chol, nRhs, rhs, sol, 2); }

int hdsdpcu_linsys_getdiag(void *chol, double *diag) {
    LinsysCU *l = (LinsysCU *) chol;
    DenseChol *c = l->c;
    HD_CUDA(cudaMemcpy2DAsync(diag, 8, c->L, (size_t) (c->np + 1) * 8, 8, c->n, cudaMemcpyDeviceToHost, g_stream));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}

void hdsdpcu_linsys_invert(void *chol, double *fullInv, double *aux) {
    (void) aux;
    LinsysCU *l = (LinsysCU *) chol;
    DenseChol *c = l->c;
    if (!l->d_inv && cudaMalloc(&l->d_inv, sizeof(double) * (size_t) c->np * c->np) != cudaSuccess) {
        fprintf(stderr, "[hdsdpcu] linsys_invert: out of device memory\n");
        return;
    }
    if (chol_invert(g_stream, c, l->d_inv) != HD_OK) return;
    if (hd_d2h_matrix(g_stream, fullInv, l->d_inv, c->np, c->n, c->work) != HD_OK) return;   // work (X = L^-T) is dead after the product
    HD_CUDA_VOID(cudaStreamSynchronize(g_stream));
}

void hdsdpcu_linsys_destroy(void **pchol) {
    if (!pchol || !*pchol) return;
    LinsysCU *l = (LinsysCU *) *pchol;
    chol_destroy(l->c);
    if (l->d_vec) cudaFree(l->d_vec);
    if (l->d_inv) cudaFree(l->d_inv);
    free(l);
    *pchol = nullptr;
}

int hdsdpcu_linsys_set_indefinite(void *chol, int on) {
    ((LinsysCU *) chol)->c->ldl = (on != 0);
    return HD_OK;
}
int hdsdpcu_linsys_inertia(void *chol, int *nNegative, int *nPerturbed) {
    DenseChol *c = ((LinsysCU *) chol)->c;
    if (!c->ldl || !c->factored) return HD_FAILED;
    if (nNegative) *nNegative = c->nnegative;
    if (nPerturbed) *nPerturbed = c->nperturbed;
    return HD_OK;
}
int hdsdpcu_linsys_padded_dim(void *chol) { return ((LinsysCU *) chol)->c->np; }
int hdsdpcu_linsys_numeric_dev(void *chol, const double *d_elem, long ld, int *info) {
    LinsysCU *l = (LinsysCU *) chol;
    HD_CALL(chol_load_dev(g_stream, l->c, d_elem, ld));
    int inf = 0;
    HD_CALL(chol_factor(g_stream, l->c, &inf));
    if (info) *info = inf;
    return HD_OK;
}
int hdsdpcu_linsys_solve_dev(void *chol, int nRhs, double *d_x, long ldx) {
    LinsysCU *l = (LinsysCU *) chol;
    if (!l->c->factored) return HD_FAILED;
    HD_CALL(chol_fsolve(g_stream, l->c, d_x, nRhs, ldx));
    HD_CALL(chol_dsolve(g_stream, l->c, d_x, nRhs, ldx));
    return chol_bsolve(g_stream, l->c, d_x, nRhs, ldx);
}
int hdsdpcu_linsys_invert_dev(void *chol, double *d_inv) { return chol_invert(g_stream, ((LinsysCU *) chol)->c, d_inv); }
double *hdsdpcu_linsys_factor_dev(void *chol) { return ((LinsysCU *) chol)->c->L; }

// ------------------------------------------------------------------------------------------------
// B2: cone
// ------------------------------------------------------------------------------------------------
int hdsdpcu_cone_create(void **pcone, int nRow, int nCol, const int *beg, const int *idx, const double *elem) {
    if (!pcone) return HD_FAILED;
    HD_CALL(ensure_ready());
    ConeCU *c = nullptr;
    int rc = cone_create(&c, nRow, nCol, beg, idx, elem);
    if (rc != HD_OK) return rc;
    *pcone = c;
    return HD_OK;
}
void hdsdpcu_cone_destroy(void **pcone) {
    if (!pcone || !*pcone) return;
    cone_destroy((ConeCU *) *pcone);
    *pcone = nullptr;
}
int hdsdpcu_cone_getdim(void *cone) { return ((ConeCU *) cone)->n; }
int hdsdpcu_cone_gettypes(void *cone, int *types) {
    ConeCU *c = (ConeCU *) cone;
    for (int i = 0; i <= c->m; ++i) types[i] = c->types[i];
    return HD_OK;
}
void hdsdpcu_cone_setstart(void *cone, double r) { ((ConeCU *) cone)->dualResidual = r; }
void hdsdpcu_cone_reduceresi(void *cone, double r) { ((ConeCU *) cone)->dualResidual = r; }
void hdsdpcu_cone_setperturb(void *cone, double p) { ((ConeCU *) cone)->dualPerturb = p; }

int hdsdpcu_cone_update(void *cone, double tau, const double *y) {
    ConeCU *c = (ConeCU *) cone;
    return cone_update_buffer(c, tau, -1.0, y, nullptr, -c->dualResidual, BUF_DUALVAR); // hdsdp_conic_sdp.c:1630
}
int hdsdpcu_cone_update_dev(void *cone, double tau, const double *d_y) {
    ConeCU *c = (ConeCU *) cone;
    return cone_update_buffer(c, tau, -1.0, nullptr, d_y, -c->dualResidual, BUF_DUALVAR);
}
int hdsdpcu_cone_updatebuffer(void *cone, double cC, double aS, const double *a, double eye, int which) {
    return cone_update_buffer((ConeCU *) cone, cC, aS, a, nullptr, eye, which);
}
int hdsdpcu_cone_factorize(void *cone, int which, int *isPsd) { return cone_factorize((ConeCU *) cone, which, isPsd); }
int hdsdpcu_cone_interiorcheck(void *cone, double tau, const double *y, int *isInterior) {
    HD_CALL(hdsdpcu_cone_update(cone, tau, y));
    return cone_factorize((ConeCU *) cone, BUF_DUALVAR, isInterior);
}
int hdsdpcu_cone_interiorcheckexpert(void *cone, double cC, double aS, const double *a, double eye, int which, int *isInterior) {
    HD_CALL(cone_update_buffer((ConeCU *) cone, cC, aS, a, nullptr, eye, which));
    return cone_factorize((ConeCU *) cone, which == BUF_DUALVAR ? BUF_DUALVAR : BUF_DUALCHECK, isInterior);
}
int hdsdpcu_cone_getbarrier(void *cone, double tau, const double *y, int which, double *logdet) {
    ConeCU *c = (ConeCU *) cone;
    if (y) {
        if (which != BUF_DUALVAR) return HD_FAILED;
        HD_CALL(hdsdpcu_cone_update(cone, tau, y));
        int psd = 0;
        HD_CALL(cone_factorize(c, BUF_DUALVAR, &psd));
        if (!psd) return HD_FAILED; // reference: HFpLinsysNumeric fails when dpotrf does
    }
    DenseChol *f = (which == BUF_DUALVAR) ? c->factor : c->checker;
    HD_CALL(chol_logdet(g_stream, f, c->d_scal + 7, nullptr));
    HD_CUDA(cudaMemcpyAsync(c->h_scal + 7, c->d_scal + 7, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    *logdet = c->h_scal[7];
    return HD_OK;
}
}

namespace {
__global__ void axpy_full_kernel(double *dst, const double *src, long total, double alpha) {
    long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < total) dst[idx] += alpha * src[idx];
}
__global__ void scal_kernel(double *x, long total, double a) {
    long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < total) x[idx] *= a;
}
} // namespace

extern "C" {

int hdsdpcu_cone_addstepandcheck(void *cone, double dStep, int which, int *isInterior) {
    // reference hdsdp_conic_sdp.c:2333-2360: (checker <- S if not DUALVAR); target += dStep * dS; PSD check
    ConeCU *c = (ConeCU *) cone;
    const long total = (long) c->np * c->np;
    int tgt = (which == BUF_DUALVAR) ? BUF_DUALVAR : BUF_DUALCHECK;
    if (tgt == BUF_DUALCHECK)
        HD_CUDA(cudaMemcpyAsync(c->d_buf[BUF_DUALCHECK], c->d_buf[BUF_DUALVAR], sizeof(double) * total, cudaMemcpyDeviceToDevice, g_stream));
    HDK(axpy_full_kernel)<<<(unsigned) ((total + 255) / 256), 256, 0, g_stream>>>(c->d_buf[tgt], c->d_buf[BUF_DUALSTEP], total, dStep);
    HD_CUDA(cudaGetLastError());
    return cone_factorize(c, tgt, isInterior);
}

int hdsdpcu_cone_scal(void *cone, double dScal) {
    // sdpDataMatScal on the objective (hdsdp_conic_sdp.c:1604-1614): scale every representation of C
    ConeCU *c = (ConeCU *) cone;
    const int m = c->m;
    HostCoeff &h = c->coeff[m];
    for (double &v : h.val) v *= dScal;
    for (double &v : h.packed) v *= dScal;
    h.sign *= dScal;
    // S-assembly lists: entries with con == m ; simplest is to rescale through a tiny kernel over all entries
    // (done on the host copy and re-uploaded: objective scaling happens once per solve)
    std::vector<int> con;
    std::vector<double> val;
    int nent = 0;
    if (c->npos > 0) {
        HD_CUDA(cudaMemcpy(&nent, c->d_pos_ptr + c->npos, sizeof(int), cudaMemcpyDeviceToHost));
        con.resize(nent); val.resize(nent);
        HD_CUDA(cudaMemcpy(con.data(), c->d_ent_con, sizeof(int) * nent, cudaMemcpyDeviceToHost));
        HD_CUDA(cudaMemcpy(val.data(), c->d_ent_val, sizeof(double) * nent, cudaMemcpyDeviceToHost));
        for (int e = 0; e < nent; ++e) if (con[e] == m) val[e] *= dScal;
        HD_CUDA(cudaMemcpy(c->d_ent_val, val.data(), sizeof(double) * nent, cudaMemcpyHostToDevice));
    }
    if (h.type == COEFF_DENSE) {
        std::vector<int> dcon(c->nds);
        HD_CUDA(cudaMemcpy(dcon.data(), c->d_dense_con, sizeof(int) * c->nds, cudaMemcpyDeviceToHost));
        for (int d = 0; d < c->nds; ++d)
            if (dcon[d] == m) HDK(scal_kernel)<<<(unsigned) ((c->npack + 255) / 256), 256, 0, g_stream>>>(c->d_dense_packed + (long) d * c->npack, c->npack, dScal);
    }
    if (h.type == COEFF_DSR1) {
        std::vector<int> dcon(c->ndr1p);
        HD_CUDA(cudaMemcpy(dcon.data(), c->d_dr1_con, sizeof(int) * c->ndr1p, cudaMemcpyDeviceToHost));
        for (int d = 0; d < c->ndr1; ++d)
            if (dcon[d] == m) HDK(scal_kernel)<<<1, 32, 0, g_stream>>>(c->d_dr1_sign + d, 1, dScal);
    }
    if (c->d_obj_val) HDK(scal_kernel)<<<(unsigned) ((c->obj_nent + 255) / 256), 256, 0, g_stream>>>(c->d_obj_val, c->obj_nent, dScal);
    if (c->d_obj_full) {
        long total = (long) c->np * c->np;
        HDK(scal_kernel)<<<(unsigned) ((total + 255) / 256), 256, 0, g_stream>>>(c->d_obj_full, total, dScal);
    }
    HD_CUDA(cudaGetLastError());
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}

int hdsdpcu_cone_buildprimalxsx(void *cone, const double *dPrimalScalMatrix, double *dPrimalXSXBuffer, int iDualMat) {
    return cone_build_xsx((ConeCU *) cone, dPrimalScalMatrix, dPrimalXSXBuffer, iDualMat);
}
int hdsdpcu_sym_extreme_eig(int n, const double *X, int largest, double *eig, int *lanczosSteps) {
    HD_CALL(ensure_ready());
    return sym_extreme_eig(n, X, largest, eig, lanczosSteps);
}
int hdsdpcu_cone_xdots(void *cone, const double *dConePrimal, double *xDotS) {
    return cone_xdots((ConeCU *) cone, dConePrimal, xDotS);
}
int hdsdpcu_cone_getdual(void *cone, double *dConeDual) {
    // coneDRecover (sdpDenseConeGetDual, hdsdp_conic_sdp.c:2497-2508): the dual matrix, full symmetric
    ConeCU *c = (ConeCU *) cone;
    HD_CALL(hd_symmetrize_lower(g_stream, c->d_buf[BUF_DUALVAR], c->np, c->np));
    return hdsdpcu_cone_getbuffer(cone, BUF_DUALVAR, dConeDual);
}
int hdsdpcu_cone_getprimal(void *cone, double dBarrierMu, const double *dRowDual, const double *dRowDualStep, double *dConePrimal, int *isFeasible) {
    return cone_get_primal((ConeCU *) cone, dBarrierMu, dRowDual, dRowDualStep, dConePrimal, isFeasible);
}
int hdsdpcu_cone_ratiotest(void *cone, double barHsdTauStep, const double *rowDualStep, double dAdaRatio, int whichBuffer, double *maxStep) {
    return cone_ratio_test((ConeCU *) cone, barHsdTauStep, rowDualStep, dAdaRatio, whichBuffer, maxStep);
}
int hdsdpcu_cone_lanczosmultiply(void *cone, int whichBuffer, const double *x, double *y) {
    return cone_lanczos_multiply((ConeCU *) cone, whichBuffer, x, y);
}
int hdsdpcu_cone_lanczossteps(void *cone) { return cone_lanczos_steps((ConeCU *) cone); }
int hdsdpcu_cone_buildschur(void *cone, int iCone, void *kkt, int typeKKT) {
    return cone_build_schur((ConeCU *) cone, iCone, (KktCU *) kkt, typeKKT);
}

int hdsdpcu_cone_setsinv(void *cone, const double *fullInv) {
    ConeCU *c = (ConeCU *) cone;
    HD_CUDA(cudaMemsetAsync(c->d_sinv, 0, sizeof(double) * (size_t) c->np * c->np, g_stream));
    HD_CUDA(cudaMemcpy2DAsync(c->d_sinv, (size_t) c->np * 8, fullInv, (size_t) c->n * 8, (size_t) c->n * 8, c->n,
                              cudaMemcpyHostToDevice, g_stream));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    c->sinv_valid = true;
    return HD_OK;
}
int hdsdpcu_cone_setsinv_linsys(void *cone, void *chol) {
    ConeCU *c = (ConeCU *) cone;
    LinsysCU *l = (LinsysCU *) chol;
    if (!l || l->c->np != c->np || !l->c->factored) return HD_FAILED;
    HD_CALL(chol_invert(g_stream, l->c, c->d_sinv));
    c->sinv_valid = true;
    return HD_OK;
}

int hdsdpcu_cone_getbuffer(void *cone, int which, double *out) {
    ConeCU *c = (ConeCU *) cone;
    double *stage = cone_scratch(c);
    if (!stage) return HD_MEMORY;
    HD_CALL(hd_d2h_matrix(g_stream, out, c->d_buf[which], c->np, c->n, stage));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}
int hdsdpcu_cone_getsinv(void *cone, double *out) {
    ConeCU *c = (ConeCU *) cone;
    double *stage = cone_scratch(c);
    if (!stage) return HD_MEMORY;
    HD_CALL(hd_d2h_matrix(g_stream, out, c->d_sinv, c->np, c->n, stage));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}
int hdsdpcu_cone_getfactordiag(void *cone, int which, double *diag) {
    ConeCU *c = (ConeCU *) cone;
    DenseChol *f = (which == BUF_DUALVAR) ? c->factor : c->checker;
    HD_CUDA(cudaMemcpy2DAsync(diag, 8, f->L, (size_t) (f->np + 1) * 8, 8, f->n, cudaMemcpyDeviceToHost, g_stream));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------
// B2: KKT
// ------------------------------------------------------------------------------------------------
int hdsdpcu_kkt_create(void **pkkt, int nRow) {
    if (!pkkt) return HD_FAILED;
    HD_CALL(ensure_ready());
    KktCU *k = nullptr;
    int rc = kkt_create(&k, nRow);
    if (rc != HD_OK) return rc;
    *pkkt = k;
    return HD_OK;
}
int hdsdpcu_kkt_addcone(void *kkt, void *cone) {
    KktCU *k = (KktCU *) kkt;
    ConeCU *c = (ConeCU *) cone;
    if (c->m != k->m) return HD_FAILED;
    k->cones.push_back(c);
    return HD_OK;
}
void hdsdpcu_kkt_destroy(void **pkkt) {
    if (!pkkt || !*pkkt) return;
    kkt_destroy((KktCU *) *pkkt);
    *pkkt = nullptr;
}
int hdsdpcu_kkt_buildup(void *kkt, int typeKKT) { return kkt_build_up((KktCU *) kkt, typeKKT); }
int hdsdpcu_kkt_clean(void *kkt, int typeKKT) { return kkt_clean((KktCU *) kkt, typeKKT); }
int hdsdpcu_kkt_buildupextra_bound(void *kkt, const double *diagAdd, const double *asinvAdd, const double *asinvRdAdd, int typeKKT) {
    if (typeKKT == KKT_PRIMAL) return HD_FAILED; // hdsdp_conic_bound.c:207-209
    return kkt_add_host((KktCU *) kkt, typeKKT == KKT_CORRECTOR ? nullptr : diagAdd, asinvAdd, asinvRdAdd, nullptr, nullptr);
}

int hdsdpcu_lp_create(void **plp, int nRow, int nLpCol, const int *beg, const int *idx, const double *elem) {
    // user data: CSC [nLpCol x (nRow + 1)], column 0 = objective, column k+1 = constraint k, row index = LP column
    HD_CALL(ensure_ready());
    std::vector<int> cnt(nLpCol + 1, 0);
    for (int k = 0; k < nRow; ++k)
        for (int e = beg[k + 1]; e < beg[k + 2]; ++e) cnt[idx[e] + 1] += 1;
    for (int c = 0; c < nLpCol; ++c) cnt[c + 1] += cnt[c];
    std::vector<int> ptr(cnt), rowidx(cnt[nLpCol]);
    std::vector<double> val(cnt[nLpCol]);
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int k = 0; k < nRow; ++k)
        for (int e = beg[k + 1]; e < beg[k + 2]; ++e) {
            int p = fill[idx[e]]++;
            rowidx[p] = k; val[p] = elem[e];
        }
    LpCU *lp = (LpCU *) calloc(1, sizeof(LpCU));
    lp->m = nRow; lp->ncol = nLpCol;
    HD_CUDA(cudaMalloc(&lp->d_colptr, sizeof(int) * (nLpCol + 1)));
    HD_CUDA(cudaMalloc(&lp->d_rowidx, sizeof(int) * (rowidx.size() + 1)));
    HD_CUDA(cudaMalloc(&lp->d_val, sizeof(double) * (val.size() + 1)));
    HD_CUDA(cudaMalloc(&lp->d_sinv, sizeof(double) * (nLpCol + 1)));
    HD_CUDA(cudaMemcpy(lp->d_colptr, ptr.data(), sizeof(int) * (nLpCol + 1), cudaMemcpyHostToDevice));
    HD_CUDA(cudaMemcpy(lp->d_rowidx, rowidx.data(), sizeof(int) * rowidx.size(), cudaMemcpyHostToDevice));
    HD_CUDA(cudaMemcpy(lp->d_val, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice));
    // objective (column 0 of the user data): needed by the HOMOGENEOUS terms
    std::vector<double> obj(nLpCol + 1, 0.0);
    for (int e = beg[0]; e < beg[1]; ++e) obj[idx[e]] = elem[e];
    HD_CUDA(cudaMalloc(&lp->d_obj, sizeof(double) * (nLpCol + 1)));
    HD_CUDA(cudaMemcpy(lp->d_obj, obj.data(), sizeof(double) * (nLpCol + 1), cudaMemcpyHostToDevice));
    *plp = lp;
    return HD_OK;
}
int hdsdpcu_lp_setobjective(void *plp, const double *colObj) {
    // the host solver may rescale the objective after the image was created (LPConeScal, hdsdp_conic_lp.c); ncol doubles
    LpCU *lp = (LpCU *) plp;
    HD_CUDA(cudaMemcpyAsync(lp->d_obj, colObj, sizeof(double) * lp->ncol, cudaMemcpyHostToDevice, g_stream));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    return HD_OK;
}
void hdsdpcu_lp_destroy(void **plp) {
    if (!plp || !*plp) return;
    LpCU *lp = (LpCU *) *plp;
    cudaFree(lp->d_colptr); cudaFree(lp->d_rowidx); cudaFree(lp->d_val); cudaFree(lp->d_sinv); cudaFree(lp->d_obj);
    free(lp);
    *plp = nullptr;
}
int hdsdpcu_kkt_buildupextra_lp(void *kkt, void *plp, const double *colDualInverse, double dualResidual, int typeKKT) {
    LpCU *lp = (LpCU *) plp;
    KktCU *k = (KktCU *) kkt;
    if (lp->m != k->m) return HD_FAILED;
    // M += A D^2 A^T, A s^-1, R_d A s^-2, dTraceSinv and (HOMOGENEOUS) dCSinv, dCSinvCSinv, A C s^-2: all on the device
    return kkt_add_lp(k, lp->ncol, lp->d_colptr, lp->d_rowidx, lp->d_val, lp->d_obj, colDualInverse, lp->d_sinv, dualResidual, typeKKT);
}
int hdsdpcu_kkt_regularize(void *kkt, double reg) { return kkt_regularize((KktCU *) kkt, reg); }
int hdsdpcu_kkt_export(void *kkt, double *a, double *ard, double *ac, double *cscs, double *cs, double *csrd, double *tr) {
    return kkt_export((KktCU *) kkt, a, ard, ac, cscs, cs, csrd, tr);
}
int hdsdpcu_kkt_factorize(void *kkt) { return kkt_factorize((KktCU *) kkt, nullptr); }
int hdsdpcu_kkt_ldl_status(void *kkt, int *isLdl, int *nNegative, int *nPerturbed) {
    KktCU *k = (KktCU *) kkt;
    if (isLdl) *isLdl = k->chol->ldl ? 1 : 0;
    if (nNegative) *nNegative = k->chol->ldl ? k->chol->nnegative : 0;
    if (nPerturbed) *nPerturbed = k->chol->ldl ? k->chol->nperturbed : 0;
    return HD_OK;
}
int hdsdpcu_kkt_symv(void *kkt, const double *x, double *y) {
    // y = M x with the assembled Schur matrix (lower triangle in HBM): the dsymv of the reference's PCG loop; test / PCG building block
    KktCU *k = (KktCU *) kkt;
    const size_t bytes = sizeof(double) * (size_t) k->mp;
    memset(k->h_vec, 0, 2 * bytes);
    memcpy(k->h_vec, x, sizeof(double) * k->m);
    HD_CUDA(cudaMemcpyAsync(k->d_rhs, k->h_vec, 2 * bytes, cudaMemcpyHostToDevice, g_stream));
    HD_CALL(kkt_symv_dev(k, k->d_rhs, k->d_rhs + k->mp, 1));
    HD_CUDA(cudaMemcpyAsync(k->h_vec, k->d_rhs + k->mp, bytes, cudaMemcpyDeviceToHost, g_stream));
    HD_CUDA(cudaStreamSynchronize(g_stream));
    memcpy(y, k->h_vec, sizeof(double) * k->m);
    return HD_OK;
}
int hdsdpcu_kkt_set_solver(void *kkt, int mode) {
    KktCU *k = (KktCU *) kkt;
    if (mode != 0 && mode != 1) return HD_FAILED;
    k->solver_mode = mode;
    k->use_jacobi = true;
    k->factored = false;
    return HD_OK;
}
int hdsdpcu_kkt_pcg_status(void *kkt, int *useJacobi, int *lastIterations, int *nSolves, int *nFallbacks) {
    KktCU *k = (KktCU *) kkt;
    if (useJacobi) *useJacobi = (k->solver_mode == 1 && k->use_jacobi) ? 1 : 0;
    if (lastIterations) *lastIterations = k->last_cg_iters;
    if (nSolves) *nSolves = k->cg_solves;
    if (nFallbacks) *nFallbacks = k->cg_fallbacks;
    return HD_OK;
}
int hdsdpcu_kkt_solve_status(void *kkt, double *relResidual, int *refineSteps) {
    KktCU *k = (KktCU *) kkt;
    if (relResidual) *relResidual = k->last_residual;
    if (refineSteps) *refineSteps = k->last_refine_steps;
    return HD_OK;
}
int hdsdpcu_kkt_solve(void *kkt, const double *rhs, double *lhs) { return kkt_solve((KktCU *) kkt, 1, rhs, lhs); }
int hdsdpcu_kkt_solve_many(void *kkt, int nRhs, const double *rhs, double *lhs) { return kkt_solve((KktCU *) kkt, nRhs, rhs, lhs); }
void hdsdpcu_kkt_registerpsdp(void *kkt, int nCones, double **X) {
    KktCU *k = (KktCU *) kkt;
    k->primalX.assign(nCones > 0 ? nCones : 0, nullptr);
    if (X) for (int i = 0; i < nCones; ++i) k->primalX[i] = X[i];
}
int hdsdpcu_kkt_addhost(void *kkt, const double *d, const double *a, const double *ard, const double *ac, const double *s4) {
    return kkt_add_host((KktCU *) kkt, d, a, ard, ac, s4);
}
int hdsdpcu_kkt_getmatrix(void *kkt, double *M) { return kkt_get_matrix((KktCU *) kkt, M); }
int hdsdpcu_kkt_padded_dim(void *kkt) { return ((KktCU *) kkt)->mp; }
double *hdsdpcu_kkt_matrix_dev(void *kkt) { return ((KktCU *) kkt)->d_M; }
double *hdsdpcu_kkt_asinv_dev(void *kkt) { return ((KktCU *) kkt)->d_asinv; }
int hdsdpcu_kkt_solve_dev(void *kkt, int nRhs, double *d_x) { return kkt_solve_dev((KktCU *) kkt, d_x, nRhs); }
int hdsdpcu_kkt_setshard(void *kkt, int rank, int nRanks) {
    KktCU *k = (KktCU *) kkt;
    if (nRanks < 1 || rank < 0 || rank >= nRanks) return HD_FAILED;
    k->rank = rank; k->nranks = nRanks; k->shard_nb = HD_LEAF;
    return HD_OK;
}

// ---- multi-GPU factorisation of M (dist.cu) ----------------------------------------------------
int hdsdpcu_dist_blob_bytes(void) { return 3 * (int) sizeof(cudaIpcMemHandle_t); }
int hdsdpcu_dist_owner(int col, int blockSize, int nRanks) { return (blockSize > 0 && nRanks > 0) ? (col / blockSize) % nRanks : -1; }

int hdsdpcu_kkt_dist_init(void *kkt, int rank, int nRanks, int blockSize) {
    HD_CALL(ensure_ready());
    KktCU *k = (KktCU *) kkt;
    if (nRanks < 1 || rank < 0 || rank >= nRanks || blockSize < HD_LEAF || blockSize % HD_LEAF || k->dist) return HD_FAILED;
    k->rank = rank; k->nranks = nRanks; k->shard_nb = blockSize;
    if (nRanks == 1) return HD_OK;
    DenseChol *chols[1] = {k->chol};
    HD_CALL(dist_create(&k->dist, k->m, blockSize, nRanks, 1, &rank, chols, g_stream));
    HD_CUDA(cudaMalloc(&k->d_gather, sizeof(double) * 16 * 8));
    return HD_OK;
}
int hdsdpcu_kkt_dist_export(void *kkt, void *blob) {
    KktCU *k = (KktCU *) kkt;
    if (!k->dist) return HD_FAILED;
    return dist_export(k->dist, 0, blob);
}
int hdsdpcu_kkt_dist_connect(void *kkt, const void *blobs) {
    KktCU *k = (KktCU *) kkt;
    if (!k->dist) return HD_FAILED;
    return dist_connect(k->dist, blobs);
}

// The whole distributed schedule with nRanks ranks living in this process on the current device (CUDA events
// instead of peer flags): rank r gets only the block columns it owns (the rest of its buffer is poisoned with NaN),
// every rank must end up with the complete factor.  outL[r] (n x n, lower meaningful) for r = 0 and nRanks-1.
int hdsdpcu_distchol_selftest(int n, int blockSize, int nRanks, const double *A, double *outL0, double *outLlast, int *info,
                              int indefinite, double *outSign0) {
    HD_CALL(ensure_ready());
    if (nRanks < 1 || nRanks > 16) return HD_FAILED;
    DistChol *d = nullptr;
    int ranks[16];
    for (int r = 0; r < nRanks; ++r) ranks[r] = r;
    HD_CALL(dist_create(&d, n, blockSize, nRanks, nRanks, ranks, nullptr, nullptr));
    const int np = hd_pad(n);
    std::vector<double> stage((size_t) np * np);
    int rc = HD_OK;
    for (int r = 0; r < nRanks && rc == HD_OK; ++r) {
        DenseChol *c = dist_local_chol(d, r);
        for (int j = 0; j < np; ++j) {
            const bool mine = (j / blockSize) % nRanks == r;
            for (int i = 0; i < np; ++i) {
                double v = __builtin_nan("");
                if (mine && i >= (j / blockSize) * blockSize) v = (i < n && j < n) ? A[(size_t) j * n + i] : (i == j ? 1.0 : 0.0);
                stage[(size_t) j * np + i] = v;
            }
        }
        if (cudaMemcpy(c->L, stage.data(), sizeof(double) * (size_t) np * np, cudaMemcpyHostToDevice) != cudaSuccess) rc = HD_FAILED;
    }
    if (rc == HD_OK) rc = dist_factor(d, info, indefinite != 0);
    if (rc == HD_OK && indefinite && outSign0 &&
        cudaMemcpy(outSign0, dist_local_chol(d, 0)->sgn, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = HD_FAILED;
    for (int pass = 0; pass < 2 && rc == HD_OK; ++pass) {
        double *out = pass == 0 ? outL0 : outLlast;
        if (!out) continue;
        DenseChol *c = dist_local_chol(d, pass == 0 ? 0 : nRanks - 1);
        if (cudaMemcpy2D(out, (size_t) n * 8, c->L, (size_t) np * 8, (size_t) n * 8, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = HD_FAILED;
    }
    dist_destroy(d);
    return rc;
}

int hdsdpcu_dgemm_nt_dev(int M, int N, int K, double alpha, const double *dA, long lda, const double *dB, long ldb,
                         double beta, double *dC, long ldc, int lowerOnly) {
    HD_CALL(ensure_ready());
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K; g.A = dA; g.lda = lda; g.B = dB; g.ldb = ldb; g.C = dC; g.ldc = ldc;
    g.alpha = alpha; g.beta = beta; g.flags = lowerOnly ? HD_GEMM_LOWER : 0;
    return hd_gemm_nt(g_stream, g);
}

} // extern "C"
