/*
 * oracle/ref_driver.c -- ctypes-callable driver around the UNMODIFIED reference objects.
 *
 * TEST INFRASTRUCTURE ONLY (checker / CPU baseline).  Never on the product path.
 *
 * This file is our own code; it only calls the reference's public cone/KKT/solver API
 * (interface/hdsdp_conic.h:27-63, interface/hdsdp_schur.h:10-22, interface/hdsdp.h:108-120),
 * following the call sequence of the reference's cone+KKT harness tests/test_file_io.c:356-467
 * minus the two HConeRatioTest calls (they abort on mcp100).
 * It is compiled together with the reference objects into oracle/_ref/libhdsdp_ref.so by
 * oracle/build_ref.sh and is used to
 *   - dump (S, diag L, S^-1, M, dASinvVec, dASinvRdSinvVec, dASinvCSinvVec, scalars) at a given
 *     operating point (y, tau, R_d, typeKKT) -> golden fixtures and live parity checks,
 *   - run full HDSDPOptimize solves (objective / iteration-count parity, CPU baseline timing),
 *   - time the reference's HKKTBuildUp + HKKTFactorize + HKKTSolve on host cores (cpu_baseline).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "interface/hdsdp.h"
#include "interface/def_hdsdp.h"
#include "interface/hdsdp_utils.h"
#include "interface/hdsdp_conic.h"
#include "interface/hdsdp_schur.h"
#include "interface/hdsdp_user_data.h"
#include "interface/hdsdp_file_io.h"
#include "linalg/hdsdp_sdpdata.h"
#include "linalg/def_hdsdp_sdpdata.h"
#include "linalg/hdsdp_linsolver.h"

#define REFDRV_MAXCONES 64

typedef struct {
    int nRows;
    int nCones;
    int coneKind[REFDRV_MAXCONES];   /* 0 = SDP (dense or sparse cone chosen by the reference), 1 = LP */
    int coneDim[REFDRV_MAXCONES];
    user_data *usr[REFDRV_MAXCONES];
    hdsdp_cone *cones[REFDRV_MAXCONES];
    int *beg[REFDRV_MAXCONES];
    int *idx[REFDRV_MAXCONES];
    double *elem[REFDRV_MAXCONES];
    hdsdp_kkt *kkt;
    int finalized;
} refdrv;

static double now_sec(void) {
    struct timeval t;
    gettimeofday(&t, NULL);
    return (double) t.tv_sec + 1e-6 * (double) t.tv_usec;
}

refdrv *refdrv_create(int nRows) {
    refdrv *h = (refdrv *) calloc(1, sizeof(refdrv));
    if (h) h->nRows = nRows;
    return h;
}

/* Data is copied: the reference keeps pointers into user_data until ProcData. */
int refdrv_add_cone(refdrv *h, int kind, int dim, const int *beg, const int *idx, const double *elem) {
    if (h->nCones >= REFDRV_MAXCONES) return -1;
    int k = h->nCones;
    int nCols = (kind == 0) ? h->nRows + 1 : h->nRows + 1;
    int nnz = beg[nCols];
    h->beg[k] = (int *) malloc(sizeof(int) * (nCols + 1));
    h->idx[k] = (int *) malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    h->elem[k] = (double *) malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    memcpy(h->beg[k], beg, sizeof(int) * (nCols + 1));
    memcpy(h->idx[k], idx, sizeof(int) * nnz);
    memcpy(h->elem[k], elem, sizeof(double) * nnz);
    h->coneKind[k] = kind;
    h->coneDim[k] = dim;
    h->nCones += 1;
    return k;
}

int refdrv_finalize(refdrv *h) {
    for (int k = 0; k < h->nCones; ++k) {
        if (HUserDataCreate(&h->usr[k]) != HDSDP_RETCODE_OK) return 1;
        HUserDataSetConeData(h->usr[k], h->coneKind[k] == 0 ? HDSDP_CONETYPE_DENSE_SDP : HDSDP_CONETYPE_LP,
                             h->nRows, h->coneDim[k], h->beg[k], h->idx[k], h->elem[k]);
        if (HConeCreate(&h->cones[k], k) != HDSDP_RETCODE_OK) return 2;
        if (HConeSetData(h->cones[k], h->usr[k]) != HDSDP_RETCODE_OK) return 3;
        if (HConeProcData(h->cones[k]) != HDSDP_RETCODE_OK) return 4;
        if (HConePresolveData(h->cones[k]) != HDSDP_RETCODE_OK) return 5;
    }
    if (HKKTCreate(&h->kkt) != HDSDP_RETCODE_OK) return 6;
    if (HKKTInit(h->kkt, h->nRows, h->nCones, h->cones) != HDSDP_RETCODE_OK) return 7;
    h->finalized = 1;
    return 0;
}

/* cone kind actually chosen by the reference: cone_type enum value */
int refdrv_cone_type(refdrv *h, int k) { return (int) h->cones[k]->cone; }
int refdrv_is_kkt_sparse(refdrv *h) { return h->kkt->isKKTSparse; }

/* S = -Rd*I - A'y + tau*C, then factorize (HConeGetLogBarrier with rowDual != NULL does update +
 * HFpLinsysNumeric, hdsdp_conic_sdp.c:2252-2290).  Returns logdet via *logdet; nonzero on failure. */
int refdrv_set_point(refdrv *h, const double *y, double tau, double rd, double *logdet) {
    double total = 0.0;
    for (int k = 0; k < h->nCones; ++k) {
        double ld = 0.0;
        HConeSetStart(h->cones[k], rd);
        HConeUpdate(h->cones[k], tau, (double *) y);
        if (HConeGetLogBarrier(h->cones[k], tau, (double *) y, BUFFER_DUALVAR, &ld) != HDSDP_RETCODE_OK) return 1 + k;
        total += ld;
    }
    if (logdet) *logdet = total;
    return 0;
}

/* HConeRatioTest (interface/hdsdp_conic.c:270 -> sdpDenseConeRatioTestImpl hdsdp_conic_sdp.c:1642): largest step alpha with
 * S + alpha dS >= 0 for cone k, dS = dAdaRatio*Rd*I - A'dy + dTau*C, by the reference's Lanczos (linalg/hdsdp_lanczos.c:161).
 * Requires refdrv_set_point before (S factorised). */
int refdrv_ratio_test(refdrv *h, int k, double dTau, const double *dy, double dAdaRatio, int whichBuffer, double *maxStep) {
    return (int) HConeRatioTest(h->cones[k], dTau, (double *) dy, dAdaRatio, whichBuffer, maxStep);
}

/* HConeGetPrimal (interface/hdsdp_conic.c:389 -> sdpDenseConeGetPrimal hdsdp_conic_sdp.c:2395): X = mu (S^-1 + S^-1 dS S^-1),
 * S = C - A'y, dS = A'dy; X and aux are dim*dim */
void refdrv_get_primal(refdrv *h, int k, double mu, const double *y, const double *dy, double *X, double *aux) {
    HConeGetPrimal(h->cones[k], mu, (double *) y, (double *) dy, X, aux);
}

/* HConeBuildPrimalXSXDirection (interface/hdsdp_conic.c:335 -> sdpDenseConeBuildPrimalXSXDirection hdsdp_conic_sdp.c:2021):
 * XSX += X * S * X with S = the dual matrix of cone k (iDualMat = 1, after refdrv_set_point) */
void refdrv_build_xsx(refdrv *h, int k, double *X, double *XSX, int iDualMat) {
    HConeBuildPrimalXSXDirection(h->cones[k], h->kkt, X, XSX, iDualMat);
}

int refdrv_interior_check(refdrv *h, const double *y, double tau, int *isInterior) {
    int all = 1;
    for (int k = 0; k < h->nCones; ++k) {
        int in = 0;
        if (HConeCheckIsInterior(h->cones[k], tau, (double *) y, &in) != HDSDP_RETCODE_OK) return 1;
        all = all && in;
    }
    *isInterior = all;
    return 0;
}

/* strategy < 0: the reference's own (auto) strategy; otherwise KKT_M2..KKT_M5 forced on every row */
int refdrv_build(refdrv *h, int typeKKT, int strategy) {
    hdsdp_retcode rc;
    if (strategy < 0) rc = HKKTBuildUp(h->kkt, typeKKT);
    else rc = HKKTBuildUpFixed(h->kkt, typeKKT, strategy);
    return (int) rc;
}

void refdrv_register_primal(refdrv *h, double **X) { HKKTRegisterPSDP(h->kkt, X); }

void refdrv_regularize(refdrv *h, double reg) { HKKTRegularize(h->kkt, reg); }

int refdrv_get_M(refdrv *h, double *M) {
    if (h->kkt->isKKTSparse) {
        /* expand CSC lower into dense column-major */
        int m = h->nRows;
        memset(M, 0, sizeof(double) * (size_t) m * m);
        for (int c = 0; c < m; ++c)
            for (int p = h->kkt->kktMatBeg[c]; p < h->kkt->kktMatBeg[c + 1]; ++p)
                M[(size_t) c * m + h->kkt->kktMatIdx[p]] = h->kkt->kktMatElem[p];
        return 0;
    }
    memcpy(M, h->kkt->kktMatElem, sizeof(double) * (size_t) h->nRows * h->nRows);
    return 0;
}

void refdrv_get_vectors(refdrv *h, double *asinv, double *asinvrd, double *asinvc, double *scalars4) {
    HKKTExport(h->kkt, asinv, asinvrd, asinvc, &scalars4[0], &scalars4[1], &scalars4[2], &scalars4[3]);
    /* scalars4 = { dCSinvCSinv, dCSinv, dCSinvRdSinv, dTraceSinv } */
}

/* S^-1 as left in kkt->invBuffer by the LAST cone's build; dim*dim doubles */
void refdrv_get_sinv(refdrv *h, int dim, double *out) {
    memcpy(out, h->kkt->invBuffer, sizeof(double) * (size_t) dim * dim);
}

/* dense dual matrix of SDP cone k (lower triangle meaningful); returns 0 if dense, 1 if sparse storage */
int refdrv_get_S(refdrv *h, int k, double *out) {
    if (h->cones[k]->cone == HDSDP_CONETYPE_DENSE_SDP) {
        hdsdp_cone_sdp_dense *c = (hdsdp_cone_sdp_dense *) h->cones[k]->coneData;
        int n = c->nCol;
        if (!c->isDualSparse) { memcpy(out, c->dualMatElem, sizeof(double) * (size_t) n * n); return 0; }
        memset(out, 0, sizeof(double) * (size_t) n * n);
        for (int j = 0; j < n; ++j)
            for (int p = c->dualMatBeg[j]; p < c->dualMatBeg[j + 1]; ++p)
                out[(size_t) j * n + c->dualMatIdx[p]] = c->dualMatElem[p];
        return 1;
    } else if (h->cones[k]->cone == HDSDP_CONETYPE_SPARSE_SDP) {
        hdsdp_cone_sdp_sparse *c = (hdsdp_cone_sdp_sparse *) h->cones[k]->coneData;
        int n = c->nCol;
        if (!c->isDualSparse) { memcpy(out, c->dualMatElem, sizeof(double) * (size_t) n * n); return 0; }
        memset(out, 0, sizeof(double) * (size_t) n * n);
        for (int j = 0; j < n; ++j)
            for (int p = c->dualMatBeg[j]; p < c->dualMatBeg[j + 1]; ++p)
                out[(size_t) j * n + c->dualMatIdx[p]] = c->dualMatElem[p];
        return 1;
    }
    return -1;
}

/* diag(L) of the dual factor of cone k */
int refdrv_get_Ldiag(refdrv *h, int k, double *out) {
    hdsdp_linsys_fp *f = NULL;
    if (h->cones[k]->cone == HDSDP_CONETYPE_DENSE_SDP) f = ((hdsdp_cone_sdp_dense *) h->cones[k]->coneData)->dualFactor;
    else if (h->cones[k]->cone == HDSDP_CONETYPE_SPARSE_SDP) f = ((hdsdp_cone_sdp_sparse *) h->cones[k]->coneData)->dualFactor;
    else return -1;
    return (int) HFpLinsysGetDiag(f, out);
}

/* Per-row classification of a dense-type SDP cone: types[m+1] (last = objective C), perm[m], strat[m].
 * For a sparse-type SDP cone: types[] for every row (ZERO where absent), perm = rowIdx order, strat = -1. */
int refdrv_get_classification(refdrv *h, int k, int *types, int *perm, int *strat) {
    int m = h->nRows;
    if (h->cones[k]->cone == HDSDP_CONETYPE_DENSE_SDP) {
        hdsdp_cone_sdp_dense *c = (hdsdp_cone_sdp_dense *) h->cones[k]->coneData;
        for (int i = 0; i < m; ++i) {
            types[i] = (int) sdpDataMatGetType(c->sdpRow[i]);
            perm[i] = c->sdpConePerm[i];
            strat[i] = c->KKTStrategies[i];
        }
        types[m] = (int) sdpDataMatGetType(c->sdpObj);
        return 0;
    } else if (h->cones[k]->cone == HDSDP_CONETYPE_SPARSE_SDP) {
        hdsdp_cone_sdp_sparse *c = (hdsdp_cone_sdp_sparse *) h->cones[k]->coneData;
        for (int i = 0; i < m; ++i) { types[i] = 0; perm[i] = -1; strat[i] = -1; }
        for (int e = 0; e < c->nRowElem; ++e) {
            types[c->rowIdx[e]] = (int) sdpDataMatGetType(c->sdpRow[e]);
            perm[e] = c->rowIdx[e];
        }
        types[m] = (int) sdpDataMatGetType(c->sdpObj);
        return 1;
    }
    return -1;
}

/* rank-one sign (scale) of row i of dense-type cone k after presolve; 0 if not rank one */
double refdrv_get_r1_sign(refdrv *h, int k, int i) {
    sdp_coeff *a = NULL;
    if (h->cones[k]->cone == HDSDP_CONETYPE_DENSE_SDP) {
        hdsdp_cone_sdp_dense *c = (hdsdp_cone_sdp_dense *) h->cones[k]->coneData;
        a = (i < h->nRows) ? c->sdpRow[i] : c->sdpObj;
    } else if (h->cones[k]->cone == HDSDP_CONETYPE_SPARSE_SDP) {
        hdsdp_cone_sdp_sparse *c = (hdsdp_cone_sdp_sparse *) h->cones[k]->coneData;
        if (i >= h->nRows) a = c->sdpObj;
        else for (int e = 0; e < c->nRowElem; ++e) if (c->rowIdx[e] == i) a = c->sdpRow[e];
    }
    if (!a) return 0.0;
    if (a->dataType == SDP_COEFF_SPR1) return ((sdp_coeff_spr1 *) a->dataMat)->spR1FactorSign;
    if (a->dataType == SDP_COEFF_DSR1) return ((sdp_coeff_dsr1 *) a->dataMat)->r1FactorSign;
    return 0.0;
}

/* HKKTFactorize + HKKTSolve (the reference's default PCG -> Cholesky path) */
int refdrv_factorize(refdrv *h) { return (int) HKKTFactorize(h->kkt); }
int refdrv_solve_rhs(refdrv *h, const double *rhs, double *sol) {
    return (int) HKKTSolve(h->kkt, (double *) rhs, sol);
}

/* State of the reference's default solver of M after the last HKKTSolve (conjGradLinSolver, linalg/hdsdp_linsolver.c:1289-1660):
 * out[0] = useJacobi (1: diagonal preconditioner still in use; 0: PCG failed once and a dpotrf of a copy of M is the
 * preconditioner from then on, :1558-1567), out[1] = CG iterations of the last solve, out[2] = number of solves,
 * out[3] = status of the last solve (iter_status).  Returns nonzero if M is not held by the dense iterative back-end. */
int refdrv_cg_status(refdrv *h, int *out) {
    if (!h->kkt || h->kkt->isKKTSparse || !h->kkt->kktM || h->kkt->kktM->LinType != HDSDP_LINSYS_DENSE_ITERATIVE) return 1;
    iterative_linsys *it = (iterative_linsys *) h->kkt->kktM->chol;
    if (!it) return 1;
    out[0] = it->useJacobi; out[1] = it->nIters; out[2] = it->nSolves; out[3] = (int) it->solStatus;
    return 0;
}

/* Time nRep x { HKKTBuildUp, [Regularize], HKKTFactorize, nSolve x HKKTSolve } on the host.
 * times[0..3] = seconds in build / factorize / solves / total (averaged over nRep). */
int refdrv_time_iteration(refdrv *h, int typeKKT, double reg, int nSolve, int nRep, double *times) {
    int m = h->nRows;
    double *rhs = (double *) malloc(sizeof(double) * m);
    double *sol = (double *) malloc(sizeof(double) * m);
    double tb = 0, tf = 0, ts = 0;
    int rc = 0;
    for (int r = 0; r < nRep && !rc; ++r) {
        double t0 = now_sec();
        rc = (int) HKKTBuildUp(h->kkt, typeKKT);
        if (reg > 0) HKKTRegularize(h->kkt, reg);
        double t1 = now_sec();
        if (!rc) rc = (int) HKKTFactorize(h->kkt);
        double t2 = now_sec();
        for (int s = 0; s < nSolve && !rc; ++s) {
            for (int i = 0; i < m; ++i) rhs[i] = h->kkt->dASinvVec[i] + (double) s;
            rc = (int) HKKTSolve(h->kkt, rhs, sol);
        }
        double t3 = now_sec();
        tb += t1 - t0; tf += t2 - t1; ts += t3 - t2;
    }
    times[0] = tb / nRep; times[1] = tf / nRep; times[2] = ts / nRep; times[3] = (tb + tf + ts) / nRep;
    free(rhs); free(sol);
    return rc;
}

void refdrv_destroy(refdrv *h) {
    if (!h) return;
    HKKTDestroy(&h->kkt);
    for (int k = 0; k < h->nCones; ++k) {
        HConeDestroy(&h->cones[k]);
        HUserDataDestroy(&h->usr[k]);
        free(h->beg[k]); free(h->idx[k]); free(h->elem[k]);
    }
    free(h);
}

/* ---------------------------------------------------------------------------------------------
 * Full solve through the public API (interface/hdsdp.h:108-120), as tests/test_file_io.c:185-278.
 * out[0]=pObj out[1]=dObj out[2]=iterations out[3]=status out[4]=seconds out[5..10]=DIMACS errors
 * ------------------------------------------------------------------------------------------- */
int refdrv_optimize(int nRows, int nCones, const int *kinds, const int *dims,
                    int **begs, int **idxs, double **elems, const double *rhs,
                    int maxIter, double *out, double *yOut) {
    hdsdp *solver = NULL;
    user_data *datas[REFDRV_MAXCONES] = {0};
    int rc = 0;
    double t0 = now_sec();
    if (HDSDPCreate(&solver) != HDSDP_RETCODE_OK) return 1;
    if (HDSDPInit(solver, nRows, nCones) != HDSDP_RETCODE_OK) { rc = 2; goto done; }
    for (int k = 0; k < nCones; ++k) {
        if (HUserDataCreate(&datas[k]) != HDSDP_RETCODE_OK) { rc = 3; goto done; }
        HUserDataSetConeData(datas[k], kinds[k] == 0 ? HDSDP_CONETYPE_DENSE_SDP : HDSDP_CONETYPE_LP,
                             nRows, dims[k], begs[k], idxs[k], elems[k]);
        if (HDSDPSetCone(solver, k, datas[k]) != HDSDP_RETCODE_OK) { rc = 4; goto done; }
    }
    HDSDPSetDualObjective(solver, (double *) rhs);
    if (maxIter > 0) HDSDPSetIntParam(solver, INT_PARAM_MAXITER, maxIter);
    rc = (int) HDSDPOptimize(solver, 1);
    out[0] = solver->pObjVal; out[1] = solver->dObjVal;
    out[2] = (double) solver->nIterCount; out[3] = (double) solver->HStatus;
    out[4] = now_sec() - t0;
    for (int i = 0; i < 6; ++i) out[5 + i] = solver->dErrs[i];
    if (yOut) {
        double p, d;
        HDSDPGetRowDual(solver, &p, &d, yOut);
    }
done:
    for (int k = 0; k < nCones; ++k) HUserDataDestroy(&datas[k]);
    HDSDPDestroy(&solver);
    return rc;
}

/* SDPA reader passthrough (interface/hdsdp_file_io.c:34): fills caller-visible pointers that stay
 * owned by this library until refdrv_free_sdpa. Only SDP blocks + one optional LP block. */
typedef struct {
    int nConstrs, nBlks, nLpCols, nCols, nElem;
    int *blkDims; double *rowRHS;
    int **beg; int **idx; double **elem;
    int *lpBeg; int *lpIdx; double *lpElem;
} refdrv_sdpa;

refdrv_sdpa *refdrv_read_sdpa(const char *fname) {
    refdrv_sdpa *s = (refdrv_sdpa *) calloc(1, sizeof(refdrv_sdpa));
    if (HReadSDPA((char *) fname, &s->nConstrs, &s->nBlks, &s->blkDims, &s->rowRHS, &s->beg, &s->idx, &s->elem,
                  &s->nCols, &s->nLpCols, &s->lpBeg, &s->lpIdx, &s->lpElem, &s->nElem) != HDSDP_RETCODE_OK) {
        free(s);
        return NULL;
    }
    return s;
}
int refdrv_sdpa_nconstrs(refdrv_sdpa *s) { return s->nConstrs; }
int refdrv_sdpa_nblks(refdrv_sdpa *s) { return s->nBlks; }
int refdrv_sdpa_nlp(refdrv_sdpa *s) { return s->nLpCols; }
int refdrv_sdpa_blkdim(refdrv_sdpa *s, int k) { return s->blkDims[k]; }
double *refdrv_sdpa_rhs(refdrv_sdpa *s) { return s->rowRHS; }
int *refdrv_sdpa_beg(refdrv_sdpa *s, int k) { return k < s->nBlks ? s->beg[k] : s->lpBeg; }
int *refdrv_sdpa_idx(refdrv_sdpa *s, int k) { return k < s->nBlks ? s->idx[k] : s->lpIdx; }
double *refdrv_sdpa_elem(refdrv_sdpa *s, int k) { return k < s->nBlks ? s->elem[k] : s->lpElem; }
