// hdsdp_b200/csrc/common.h -- shared declarations of the sm_100a hot-path library.
//
// Everything here is FP64, column-major, and sized in "leaves" of HD_LEAF = 128 rows/cols:
// every dense matrix the library owns is padded to a multiple of 128 in both dimensions
// (identity / zero padding) so the DMMA tile kernels never need edge predicates.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define HD_LEAF 128

// every kernel launch of the library goes through HDK(kernel)<<<...>>> so launches can be counted
extern long g_hd_launches;
#define HDK(k) (++g_hd_launches, k)

// hdsdp_retcode values (reference interface/hdsdp.h:42-48)
#define HD_OK 0
#define HD_FAILED 1
#define HD_MEMORY 2

#define HD_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "[hdsdpcu] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_),    \
                    __FILE__, __LINE__, cudaGetErrorString(e_));                               \
            return (e_ == cudaErrorMemoryAllocation) ? HD_MEMORY : HD_FAILED;                  \
        }                                                                                      \
    } while (0)

#define HD_CUDA_VOID(call)                                                                     \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "[hdsdpcu] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_),    \
                    __FILE__, __LINE__, cudaGetErrorString(e_));                               \
        }                                                                                      \
    } while (0)

#define HD_CALL(call)                     \
    do {                                  \
        int rc_ = (call);                 \
        if (rc_ != HD_OK) return rc_;     \
    } while (0)

static inline int hd_pad(int n) { return ((n + HD_LEAF - 1) / HD_LEAF) * HD_LEAF; }

// ---------------------------------------------------------------------------------------------
// DMMA GEMM:  C (M x N) = alpha * A (M x K) * B (N x K)^T + beta * C      (all column-major)
// M, N multiples of 128; K multiple of 16; pointers 16-byte aligned; ld* even.
// ---------------------------------------------------------------------------------------------
enum : int {
    HD_GEMM_LOWER = 1,      // only tiles that intersect the lower triangle (m >= n); strict-upper entries untouched
    HD_GEMM_KTRI_MAX = 2,   // operands are "upper triangular in (row,k)": start k at min(m0,n0)... see gemm_nt.cu
    HD_GEMM_EPI_HADSQ = 4,  // epilogue C += sa[m]*sb[n]*acc^2 (rank-one Schur, M2)
    HD_GEMM_KTRI_A = 16,    // A is upper triangular in (row, k): A[i, k] == 0 for k < i, so the k-range of a tile starts at its first row
    HD_GEMM_EPI_COLSCALE = 8, // beta == 0 only: C[:, n] = alpha * acc * sb[n]   (LDL^T panel solves: X = (A Dinv^T) J)
};

struct DenseChol;
struct GemmArgs {
    int M, N, K;
    const double *A; long lda;
    const double *B; long ldb;
    double *C; long ldc;
    double alpha, beta;
    int flags;
    const double *sa;   // HADSQ: scale over m
    const double *sb;   // HADSQ: scale over n
    double *peerC;       // optional second destination with the layout of C (beta == 0, full tiles): the tile is ALSO stored
                         // there -- a peer GPU's buffer over NVLink (dist.cu: panel solve fused with the hand-off to the next owner)
    const double *ksign; // optional +-1 per k: C = alpha * A diag(ksign) B^T + beta C   (LDL^T updates: A22 -= L21 J L21^T)
    // block-cyclic N (distributed Cholesky, dist.cu): the N dimension enumerates only the column blocks this rank owns.
    // Local column c lives at global column (c / bc_nb) * bc_stride + c % bc_nb (relative to B / C); 0 = contiguous.
    int bc_nb, bc_stride;
};

int hd_gemm_nt(cudaStream_t st, const GemmArgs &g);
// recursive kernels of chol.cu, enqueue-only (used by the single-GPU driver and by dist.cu)
int hd_potrf_rec(cudaStream_t st, double *A, long lda, int n, double *dinv, int *info, int base);
int hd_trsm_rec(cudaStream_t st, double *B, long ldb, int rows, const double *L, long ldl, int n, const double *dinv);
int hd_chol_finish(cudaStream_t st, DenseChol *c);
// while set, the leaf products of hd_trsm_rec also store their result at peer + (C - local)  (enqueue-time state)
void hd_trsm_set_peer(const double *local, double *peer);
void hd_ldl_scope(DenseChol *c);
int chol_ldl_prepare(cudaStream_t st, DenseChol *c, int nb, int rank, int nranks); // transposed inverse leaves for the L^T solve
int hd_num_sms();
void hd_gemm_set_variant(int v);
int hd_gemm_get_variant();
void hd_gemm_set_thin(int max_tiles);
void hd_chol_set_graph(int on);
void hd_chol_set_block(int nb);
void hd_chol_set_sched(int v);
void hd_dist_set_delay(int on);
void hd_chol_set_leaf(int v);
void hd_chol_set_ldl_pivot(int v);
void hd_chol_set_invert_fork(int v);
void hd_chol_set_tail(int v);
void hd_chol_set_partition(int v);
int hd_chol_partition_sms(int *chain, int *bulk);
void hd_trsv_set_version(int v);
int hd_leaf_clocks(long long *out);

// ---------------------------------------------------------------------------------------------
// Dense SPD factorisation object (device resident).
// ---------------------------------------------------------------------------------------------
struct DenseChol {
    int n;        // logical dimension
    int np;       // padded dimension (multiple of 128) == leading dimension
    double *L;    // np x np factor (lower); strict upper of diagonal leaves is garbage
    double *Dinv; // (np/128) inverse leaves, each 128 x 128 column-major lower triangular
    double *DinvT; // their transposes (for the L^T solve)
    int *sync;    // ticket + ready flags of the one-launch triangular solves
    int *dinfo;   // device: 0 ok, else 1-based index of the first non-positive pivot
    int *hinfo;   // pinned host mirror
    double *work; // np x np workspace (inverse / staging), allocated lazily
    bool factored;
    // LDL^T fallback (reference dsytrf path): A = L J L^T, J = diag(sgn); allocated on first use
    bool ldl;
    double *sgn;     // np entries, +-1 (stored behind the inverse leaves in the Dinv allocation)
    double *dfloor;  // device scalar: static-pivoting floor = 1e-13 max|diag A|
    int *dperturb;   // device: number of pivots replaced by the floor
    int nperturbed, nnegative;
    // CUDA graph of the factorisation (launch-bound sizes), see chol_factor
    void *graph_exec;
    unsigned long long graph_key, eager_key;
    long graph_kernels;
};

int chol_create(DenseChol **pc, int n);
void chol_destroy(DenseChol *c);
int chol_ensure_work(DenseChol *c);
// copy an n x n column-major device matrix (leading dim lds) into the padded factor buffer
int chol_load_dev(cudaStream_t st, DenseChol *c, const double *dS, long lds);
// factor whatever is in c->L ; *info: 0 = SPD, >0 = not positive definite (LAPACK dpotrf semantics)
int chol_factor(cudaStream_t st, DenseChol *c, int *info);
// inv (np x np, ld = np, full symmetric) = (L L^T)^-1 ; uses c->work
int chol_invert(cudaStream_t st, DenseChol *c, double *inv);
// in-place triangular solves on nrhs device vectors (each of length >= n, stride ldx)
int chol_fsolve(cudaStream_t st, DenseChol *c, double *x, int nrhs, long ldx);
int chol_bsolve(cudaStream_t st, DenseChol *c, double *x, int nrhs, long ldx);
// x <- J x between the two triangular solves of an LDL^T factor (no-op for Cholesky)
int chol_dsolve(cudaStream_t st, DenseChol *c, double *x, int nrhs, long ldx);
// sum_i log(L_ii) * 2 written to *dlogdet (device) ; diag(L) to ddiag (device, n) if non-null
int hd_trsv(cudaStream_t st, bool transposed, const double *L, long ldl, const double *Dinv, const double *DinvT, int np,
            double *x, int nrhs, long ldx, int *sync);
int chol_logdet(cudaStream_t st, DenseChol *c, double *dlogdet, double *ddiag);

// small utility kernels (util.cu)
int hd_set_identity(cudaStream_t st, double *A, long lda, int n);
int hd_pad_identity(cudaStream_t st, double *A, long lda, int n, int np);
int hd_symmetrize_lower(cudaStream_t st, double *A, long lda, int n);
int hd_scale_vec(cudaStream_t st, double *x, int m, double a);
int hd_h2d_matrix(cudaStream_t st, double *d_dst, long ldd, const double *h_src, int n, double *d_stage);
int hd_d2h_matrix(cudaStream_t st, double *h_dst, const double *d_src, long lds, int n, double *d_stage);
long hd_symv_ws_doubles(int np);
int hd_symv_lower(cudaStream_t st, const double *A, long ld, int np, const double *d_x, long ldx, double *d_y, long ldy, int nRhs, double alpha,
                  double *ws);
int hd_copy2d(cudaStream_t st, double *dst, long ldd, const double *src, long lds, int rows, int cols);
