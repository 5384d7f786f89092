// hdsdp_b200/csrc/cone.h -- device image of one SDP cone and of the KKT (Schur) object.
#pragma once
#include "common.h"
#include <vector>

// sdp_coeff_type (reference linalg/def_hdsdp_sdpdata.h:25-33)
enum : int { COEFF_ZERO = 0, COEFF_SPARSE = 1, COEFF_DENSE = 2, COEFF_SPR1 = 3, COEFF_DSR1 = 4 };
// KKT types (reference interface/hdsdp_conic.h:16-19)
enum : int { KKT_INFEASIBLE = 0, KKT_CORRECTOR = 1, KKT_HOMOGENEOUS = 2, KKT_PRIMAL = 3 };
// buffers (reference interface/hdsdp_conic.h:24-26)
enum : int { BUF_DUALVAR = 0, BUF_DUALCHECK = 1, BUF_DUALSTEP = 2 };

// Host-side classified coefficient (mirror of sdp_coeff after presolve)
struct HostCoeff {
    int type = COEFF_ZERO;
    // SPARSE: lower triplets
    std::vector<int> row, col;
    std::vector<double> val;
    // DENSE: packed lower column-major n(n+1)/2
    std::vector<double> packed;
    // SPR1 / DSR1: A = sign * a a^T (a normalised, sign carries the scale)
    double sign = 0.0;
    std::vector<int> idx;     // SPR1 only
    std::vector<double> fac;  // SPR1: values at idx ; DSR1: dense n-vector
};

// Group of "small sparse" constraints whose S^-1 columns are staged together in shared memory
struct SsGroup { int first, count, ent_first, ent_count; };

struct KktCU;
struct DistChol;
struct LanczosCU;

struct ConeCU {
    int m = 0;   // constraints
    int n = 0;   // cone dimension
    int np = 0;  // padded dimension / leading dimension of every n x n device matrix
    double dualResidual = 0.0, dualPerturb = 0.0;

    std::vector<HostCoeff> coeff; // m + 1 entries, [m] is the objective C
    std::vector<int> types;       // m + 1

    // ---- S assembly -------------------------------------------------------------------------
    // position-sorted scatter list (SPARSE entries and SPR1 outer products, constraints + C)
    int npos = 0;
    int *d_pos = nullptr;       // [npos] linear index row + col*np (row >= col)
    int *d_pos_ptr = nullptr;   // [npos+1]
    int *d_ent_con = nullptr;   // [nent] coefficient index (0..m-1 constraint, m = C)
    double *d_ent_val = nullptr;
    // DENSE coefficients, packed, one column per matrix: [npack x nds]
    int nds = 0; long npack = 0;
    double *d_dense_packed = nullptr;
    int *d_dense_con = nullptr; // [nds] coefficient index
    double *d_dense_part = nullptr; // [chunks x npack] partial sums of the chunked packed axpy (many dense coefficients)
    double *d_dd_part = nullptr;    // [32 x ndp] partial traces of the batched dense rows
    // DSR1 coefficients: factor matrix F [np x nr1dp] (column = a_i), con index, sign
    int ndr1 = 0, ndr1p = 0;
    double *d_dr1_F = nullptr; double *d_dr1_W = nullptr; // W = F diag(coef*sign) workspace
    int *d_dr1_con = nullptr; double *d_dr1_sign = nullptr;

    double *d_coef = nullptr;   // [m+1] scaled coefficients of the current update
    double *h_coef = nullptr;   // pinned staging
    double *d_buf[3] = {nullptr, nullptr, nullptr}; // S, checker, dS : np x np
    DenseChol *factor = nullptr, *checker = nullptr;
    double *d_sinv = nullptr;   // np x np full symmetric
    double *d_scal = nullptr;   // device scalars [8]
    double *h_scal = nullptr;   // pinned

    // ---- Schur classes (constraints only; sorted by original index inside each class) ---------
    // R: rank-one (SPR1 + DSR1)
    int nr = 0, nrp = 0;
    int *d_r_con = nullptr; double *d_r_sign = nullptr;
    double *d_r_At = nullptr;   // [nrp x np] row i = a_i^T  (column-major, ld nrp)
    double *d_r_Vt = nullptr;   // [nrp x np] workspace V^T = A^T S^-1
    double *d_r_G = nullptr;    // [nrp x nrp] Gram workspace when R does not map 1:1 onto the Schur rows
    bool r_all_unit = false;    // every factor is a unit vector e_k
    int *d_r_unit = nullptr;    // [nr] k_i when r_all_unit
    bool r_identity_map = false; // R covers constraints 0..m-1 in order
    // sparse view of SPR1 factors (CSR by rank-one constraint) for dots against explicit matrices
    int *d_r_sp_ptr = nullptr; int *d_r_sp_idx = nullptr; double *d_r_sp_val = nullptr; bool r_has_sparse_view = false;
    // SS: small sparse (nnz <= SS_MAX), SB: big sparse
    int nss = 0; int ss_nent = 0;
    int *d_ss_con = nullptr; int *d_ss_ptr = nullptr; int *d_ss_row = nullptr; int *d_ss_col = nullptr;
    double *d_ss_val = nullptr; // pre-scaled: 0.5 * x on the diagonal
    std::vector<SsGroup> ss_groups; SsGroup *d_ss_groups = nullptr; int ss_stage_pairs = 0;
    int nsb = 0;
    std::vector<int> sb_con, sb_ptr; // host copies (we loop over big rows on the host)
    int *d_sb_ptr = nullptr; int *d_sb_row = nullptr; int *d_sb_col = nullptr; double *d_sb_val = nullptr; int *d_sb_con = nullptr;
    // D: dense
    int nd = 0;
    std::vector<int> d_con_host;
    double *d_dn_full = nullptr; // [np*np x nd] each matrix full symmetric np x np (zero padded)
    // batched dense x dense block (nd >= DD_MIN): "vec" layout, constraint index fastest:
    //   d_dn_vec[i + (c + k*np) * ndp] = A_i[c, k];  U / Ut workspaces of the same shape, G ndp x ndp
    int ndp = 0;
    double *d_dn_vec = nullptr, *d_dn_U = nullptr, *d_dn_Ut = nullptr, *d_dn_G = nullptr;
    int *d_dn_con = nullptr;
    // objective in Schur-usable form
    int obj_type = COEFF_ZERO;
    int obj_nent = 0; int *d_obj_row = nullptr; int *d_obj_col = nullptr; double *d_obj_val = nullptr; // sparse C (pre-scaled)
    double *d_obj_full = nullptr; // dense / rank-one C expanded to full np x np

    double *d_U = nullptr, *d_B = nullptr; // np x np workspaces for explicit S^-1 A S^-1
    double *d_sbv_part = nullptr;          // partial sums of the many-nnz side vectors
    double *d_prim = nullptr;              // np x np: S^-1 of the checker buffer (primal recovery)
    bool sinv_valid = false;
    LanczosCU *lanczos = nullptr; // ratio-test state (lanczos.cu), created on first use
};

struct KktCU {
    int m = 0, mp = 0;
    double *d_M = nullptr;       // mp x mp column-major, lower triangle meaningful
    DenseChol *chol = nullptr;   // factor of M (separate buffer: M stays intact, reference CG back-end semantics)
    double *d_asinv = nullptr, *d_asinvrd = nullptr, *d_asinvc = nullptr; // [mp]
    double *d_scal = nullptr;    // [8]: 0 dCSinvCSinv, 1 dCSinv, 2 dCSinvRdSinv, 3 dTraceSinv, 4 scratch min-diag
    double *h_scal = nullptr;    // pinned [8]
    double *d_rhs = nullptr;     // [mp x 8] solve staging
    double *h_vec = nullptr;     // pinned [mp x 8]
    std::vector<ConeCU *> cones;
    std::vector<double *> primalX; // host pointers registered by HKKTRegisterPSDP
    bool factored = false;
    bool fresh = false; // M was zeroed by HKKTClean and nothing has been accumulated yet
    // multi-GPU column sharding of the Schur assembly (rank r builds columns j with (j/128) % nranks == r)
    int rank = 0, nranks = 1, shard_nb = HD_LEAF;
    struct DistChol *dist = nullptr; // distributed factorisation of M (dist.cu); null on one GPU
    double *d_gather = nullptr;      // [16 x 8] all-gather scratch
    double *d_ref = nullptr;         // [mp x 8] right-hand sides + residuals of the refined LDL^T solves
    double last_residual = 0.0;      // max |b - M x| / max |b| of the last refined solve
    int last_refine_steps = 0;
    // reference solver policy for M (HDSDP_LINSYS_DENSE_ITERATIVE): 0 = direct Cholesky (default), 1 = Jacobi-PCG first, Cholesky after its first failure
    int solver_mode = 0;
    bool use_jacobi = true;
    double *d_pcg = nullptr;        // [8 x mp + 8] CG vectors and scalars
    double *d_symv_ws = nullptr;    // partial vectors of the ordered symv reduction (hd_symv_ws_doubles(mp))
    int last_cg_iters = 0, cg_solves = 0, cg_fallbacks = 0;
};

cudaStream_t hd_stream();

// dist.cu
int dist_create(DistChol **pd, int n, int nb, int P, int nlocal, const int *ranks, DenseChol **chols, cudaStream_t main_stream);
void dist_destroy(DistChol *d);
DenseChol *dist_local_chol(DistChol *d, int li);
cudaStream_t dist_local_stream(DistChol *d, int li);
int dist_block(const DistChol *d);
int dist_export(DistChol *d, int li, void *blob);   // 192 bytes
int dist_connect(DistChol *d, const void *blobs);    // P x 192 bytes, rank order
int dist_factor(DistChol *d, int *info_out, bool ldl = false);
int dist_allgather_small(DistChol *d, const double *d_val, int cnt, double *d_out);

int cone_create(ConeCU **pc, int nRow, int nCol, const int *beg, const int *idx, const double *elem);
void cone_destroy(ConeCU *c);
int cone_update_buffer(ConeCU *c, double cCoef, double aScal, const double *aCoefHost, const double *aCoefDev,
                       double eyeCoef, int which);
int cone_factorize(ConeCU *c, int which, int *isPsd);
int cone_build_schur(ConeCU *c, int iCone, KktCU *k, int typeKKT);
int cone_build_xsx(ConeCU *c, const double *Xhost, double *XSXhost, int iDualMat);
int cone_xdots(ConeCU *c, const double *Xhost, double *out);
double *cone_scratch(ConeCU *c);   // np x np device scratch (allocated on first use)
int cone_get_primal(ConeCU *c, double mu, const double *yHost, const double *dyHost, double *Xhost, int *isFeasible);
// lanczos.cu
int cone_ratio_test(ConeCU *c, double dTauStep, const double *dyHost, double dAdaRatio, int which, double *maxStep);
int cone_lanczos_multiply(ConeCU *c, int which, const double *x, double *y);
int cone_lanczos_steps(ConeCU *c);
void lz_destroy(LanczosCU *l);
int sym_extreme_eig(int n, const double *Xhost, int which, double *out, int *steps);

int kkt_create(KktCU **pk, int nRow);
void kkt_destroy(KktCU *k);
int kkt_clean(KktCU *k, int typeKKT);
int kkt_build_up(KktCU *k, int typeKKT);
int kkt_regularize(KktCU *k, double reg);
int kkt_add_host(KktCU *k, const double *diagAdd, const double *asinvAdd, const double *asinvrdAdd, const double *asinvcAdd,
                 const double *scalarsAdd4);
int kkt_add_lp(KktCU *k, int nLpCol, const int *d_colptr, const int *d_rowidx, const double *d_val, const double *d_obj,
               const double *sInvHost, double *d_sinv_stage, double rd, int typeKKT);
int kkt_export(KktCU *k, double *asinv, double *asinvrd, double *asinvc, double *csinvcsinv, double *csinv, double *csinvrd,
               double *tracesinv);
int kkt_factorize(KktCU *k, int *info_out);
int kkt_solve_dev(KktCU *k, double *d_x, int nRhs);
int kkt_symv_dev(KktCU *k, const double *d_x, double *d_y, int nRhs);
int kkt_solve(KktCU *k, int nRhs, const double *rhs, double *lhs);
int kkt_get_matrix(KktCU *k, double *Mhost);
