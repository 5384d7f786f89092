/*
 * integration/hdsdp_schur_cuda.c -- drop-in replacement of the reference's interface/hdsdp_schur.c.
 *
 * Same public functions (interface/hdsdp_schur.h:10-22), same hdsdp_kkt struct
 * (interface/def_hdsdp_schur.h:32-68); the Schur matrix M lives in HBM inside libhdsdp_cuda.so and
 * kkt->kktMatElem stays NULL.  A maintainer compiles this file INSTEAD of interface/hdsdp_schur.c and
 * links libhdsdp_cuda.so; nothing else in interface/ changes.
 *
 * What moves to the GPU (include/hdsdpcu.h):
 *   HKKTInit            -> hdsdpcu_kkt_create, one hdsdpcu_cone_create per SDP cone (from HCone->usrData,
 *                          the user_data CSC the cone was built from), hdsdpcu_lp_create per LP cone
 *   HKKTBuildUp         -> hdsdpcu_kkt_clean + per cone hdsdpcu_cone_buildschur / hdsdpcu_kkt_buildupextra_lp
 *   HKKTBuildUpExtraCone-> bound cone: the O(m) host arithmetic of sBoundConeGetKKT
 *                          (interface/hdsdp_conic_bound.c:201-249) is kept, its writes go through hdsdpcu_kkt_addhost
 *   HKKTRegularize / HKKTFactorize / HKKTSolve / HKKTExport -> hdsdpcu_kkt_*
 * S^-1 comes from the cone's own dual factor: device-to-device when that factor is the CUDA linsys back-end
 * (integration/hdsdp_linsys_cuda.c), otherwise HFpLinsysInvert on the host followed by one upload.
 *
 * This is original code written against the reference's headers; it is not derived from hdsdp_schur.c's body.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "interface/hdsdp_schur.h"
#include "interface/hdsdp_utils.h"
#include "interface/hdsdp_conic.h"
#include "interface/def_hdsdp_user_data.h"
#include "hdsdpcu.h"

int hdsdpcu_linsys_is_cuda(hdsdp_linsys_fp *lin); /* integration/hdsdp_linsys_cuda.c */

typedef struct {
    void *dkkt;              /* hdsdpcu kkt handle */
    int nCones;
    void **dcone;            /* per cone: hdsdpcu cone handle (SDP), lp handle (LP) or NULL */
    double *objScal;         /* per cone: scale applied to the objective by HConeScalByConstant (detected lazily) */
    int *scaled;
    double *hostInv;         /* maxConeDim^2 staging for host-side inverses */
    double *vecA, *vecB, *vecC, *vecD;
} kkt_cuda;

/* one side table keyed by the hdsdp_kkt pointer (the struct layout must stay the reference's) */
#define MAX_KKT 16
static hdsdp_kkt *g_keys[MAX_KKT];
static kkt_cuda *g_vals[MAX_KKT];

static kkt_cuda *lookup(hdsdp_kkt *k) {
    for (int i = 0; i < MAX_KKT; ++i) if (g_keys[i] == k) return g_vals[i];
    return NULL;
}

static void refresh_host_vectors(hdsdp_kkt *HKKT, kkt_cuda *kc, int typeKKT) {
    /* hdsdp_algo.c reads kkt->dASinvVec / dASinvRdSinvVec / dASinvCSinvVec in place (:283,:560,:600,:605,:1345) */
    hdsdpcu_kkt_export(kc->dkkt, HKKT->dASinvVec, HKKT->dASinvRdSinvVec,
                       typeKKT == KKT_TYPE_HOMOGENEOUS ? HKKT->dASinvCSinvVec : NULL,
                       &HKKT->dCSinvCSinv, &HKKT->dCSinv, &HKKT->dCSinvRdSinv, &HKKT->dTraceSinv);
}

extern hdsdp_retcode HKKTCreate( hdsdp_kkt **pHKKT ) {
    if ( !pHKKT ) return HDSDP_RETCODE_FAILED;
    hdsdp_kkt *HKKT = (hdsdp_kkt *) calloc(1, sizeof(hdsdp_kkt));
    if ( !HKKT ) return HDSDP_RETCODE_MEMORY;
    *pHKKT = HKKT;
    return HDSDP_RETCODE_OK;
}

extern hdsdp_retcode HKKTInit( hdsdp_kkt *HKKT, int nRow, int nCones, hdsdp_cone **cones ) {
    HKKT->nRow = nRow; HKKT->nCones = nCones; HKKT->cones = cones;
    int maxConeDim = 0;
    for ( int iCone = 0; iCone < nCones; ++iCone ) {
        int d = HConeGetDim(cones[iCone]);
        if ( d > maxConeDim ) maxConeDim = d;
    }
    HKKT->maxConeDim = maxConeDim;
    /* host buffers other reference code still scribbles on (HConeBuildPrimalXSXDirection, hdsdp_psdp.c:236) */
    size_t nsq = (size_t) maxConeDim * maxConeDim;
    HKKT->invBuffer = (double *) calloc(nsq, sizeof(double));
    HKKT->kktBuffer = (double *) calloc(nsq, sizeof(double));
    HKKT->kktBuffer2 = (double *) calloc(nsq, sizeof(double));
    HKKT->dASinvVec = (double *) calloc(nRow, sizeof(double));
    HKKT->dASinvCSinvVec = (double *) calloc(nRow, sizeof(double));
    HKKT->dASinvRdSinvVec = (double *) calloc(nRow, sizeof(double));
    if ( !HKKT->invBuffer || !HKKT->kktBuffer || !HKKT->kktBuffer2 || !HKKT->dASinvVec || !HKKT->dASinvCSinvVec || !HKKT->dASinvRdSinvVec )
        return HDSDP_RETCODE_MEMORY;
    HKKT->isKKTSparse = 0;   /* the GPU build always uses the dense M */
    HKKT->kktMatElem = NULL; /* M never exists on the host */
    HKKT->kktDiag = NULL;
    HKKT->kktM = NULL;
    HKKT->dPrimalX = NULL;

    kkt_cuda *kc = (kkt_cuda *) calloc(1, sizeof(kkt_cuda));
    kc->nCones = nCones;
    kc->dcone = (void **) calloc(nCones > 0 ? nCones : 1, sizeof(void *));
    kc->objScal = (double *) calloc(nCones > 0 ? nCones : 1, sizeof(double));
    kc->scaled = (int *) calloc(nCones > 0 ? nCones : 1, sizeof(int));
    kc->hostInv = (double *) calloc(nsq > 0 ? nsq : 1, sizeof(double));
    kc->vecA = (double *) calloc(nRow, sizeof(double)); kc->vecB = (double *) calloc(nRow, sizeof(double));
    kc->vecC = (double *) calloc(nRow, sizeof(double)); kc->vecD = (double *) calloc(nRow, sizeof(double));
    if ( hdsdpcu_kkt_create(&kc->dkkt, nRow) != 0 ) return HDSDP_RETCODE_FAILED;
    for ( int iCone = 0; iCone < nCones; ++iCone ) {
        hdsdp_cone *c = cones[iCone];
        user_data *u = (user_data *) c->usrData;
        if ( c->cone == HDSDP_CONETYPE_DENSE_SDP || c->cone == HDSDP_CONETYPE_SPARSE_SDP ) {
            if ( hdsdpcu_cone_create(&kc->dcone[iCone], nRow, u->nConicCol, u->coneMatBeg, u->coneMatIdx, u->coneMatElem) != 0 )
                return HDSDP_RETCODE_FAILED;
            /* every SDP cone is registered so that cone index == position (LP slots get a NULL image) */
        } else if ( c->cone == HDSDP_CONETYPE_LP ) {
            if ( hdsdpcu_lp_create(&kc->dcone[iCone], nRow, u->nConicCol, u->coneMatBeg, u->coneMatIdx, u->coneMatElem) != 0 )
                return HDSDP_RETCODE_FAILED;
        } else {
            printf("[hdsdpcu] unsupported cone type %d in HKKTInit\n", (int) c->cone);
            return HDSDP_RETCODE_FAILED;
        }
    }
    for ( int i = 0; i < MAX_KKT; ++i ) if ( !g_keys[i] ) { g_keys[i] = HKKT; g_vals[i] = kc; break; }
    printf("    Using dense Schur complement on the GPU (libhdsdp_cuda)\n");
    return HDSDP_RETCODE_OK;
}

/* ||C||_F of user column 0, to detect the objective scaling applied after the image was created */
static double raw_obj_fro(user_data *u) {
    int n = u->nConicCol; double s = 0.0;
    for ( int e = u->coneMatBeg[0]; e < u->coneMatBeg[1]; ++e ) {
        int p = u->coneMatIdx[e], col = 0; long start = 0;
        while ( col < n - 1 && start + (n - col) <= p ) { start += n - col; col++; }
        int row = (int) (p - start) + col;
        double v = u->coneMatElem[e];
        s += ( row == col ) ? v * v : 2.0 * v * v;
    }
    return sqrt(s);
}

static hdsdp_retcode build_sdp_cone( hdsdp_kkt *HKKT, kkt_cuda *kc, int iCone, int typeKKT ) {
    hdsdp_cone *c = HKKT->cones[iCone];
    void *dc = kc->dcone[iCone];
    double rd; int n; hdsdp_linsys_fp *factor;
    if ( c->cone == HDSDP_CONETYPE_DENSE_SDP ) {
        hdsdp_cone_sdp_dense *d = (hdsdp_cone_sdp_dense *) c->coneData; rd = d->dualResidual; n = d->nCol; factor = d->dualFactor;
    } else {
        hdsdp_cone_sdp_sparse *d = (hdsdp_cone_sdp_sparse *) c->coneData; rd = d->dualResidual; n = d->nCol; factor = d->dualFactor;
    }
    hdsdpcu_cone_setstart(dc, rd);
    if ( typeKKT == KKT_TYPE_HOMOGENEOUS && !kc->scaled[iCone] ) {
        /* HConeScalByConstant (hdsdp.c:315) scaled the reference's copy of C after our image was made */
        double raw = raw_obj_fro((user_data *) c->usrData), cur = HConeGetObjNorm(c, FRO_NORM);
        if ( raw > 0.0 && cur > 0.0 && fabs(cur / raw - 1.0) > 1e-14 ) hdsdpcu_cone_scal(dc, cur / raw);
        kc->scaled[iCone] = 1;
    }
    if ( typeKKT != KKT_TYPE_PRIMAL ) {
        if ( hdsdpcu_linsys_is_cuda(factor) ) {
            if ( hdsdpcu_cone_setsinv_linsys(dc, factor->chol) != 0 ) return HDSDP_RETCODE_FAILED;
        } else {
            HFpLinsysInvert(factor, kc->hostInv, HKKT->kktBuffer);
            if ( hdsdpcu_cone_setsinv(dc, kc->hostInv) != 0 ) return HDSDP_RETCODE_FAILED;
        }
    }
    (void) n;
    return (hdsdp_retcode) hdsdpcu_cone_buildschur(dc, iCone, kc->dkkt, typeKKT);
}

/* LP cone: the O(nnz) slack inversion stays on the host (hdsdp_conic_lp.c:262-271); M += A D^2 A' on the device */
static hdsdp_retcode build_lp_cone( hdsdp_kkt *HKKT, kkt_cuda *kc, int iCone, int typeKKT ) {
    hdsdp_cone_lp *lp = (hdsdp_cone_lp *) HKKT->cones[iCone]->coneData;
    if ( typeKKT == KKT_TYPE_PRIMAL ) {
        if ( !HKKT->dPrimalX || !HKKT->dPrimalX[iCone] ) return HDSDP_RETCODE_FAILED;
        for ( int i = 0; i < lp->nCol; ++i ) lp->colDualInverse[i] = HKKT->dPrimalX[iCone][i];
    } else {
        for ( int i = 0; i < lp->nCol; ++i ) lp->colDualInverse[i] = 1.0 / lp->colDual[i];
    }
    if ( hdsdpcu_kkt_buildupextra_lp(kc->dkkt, kc->dcone[iCone], lp->colDualInverse, lp->dualResidual, typeKKT) != 0 )
        return HDSDP_RETCODE_FAILED;
    if ( typeKKT == KKT_TYPE_HOMOGENEOUS ) { /* hdsdp_conic_lp.c:316-327: O(nnz) host arithmetic, added through addhost */
        double add[4] = {0, 0, 0, 0};
        memset(kc->vecC, 0, sizeof(double) * HKKT->nRow);
        for ( int i = 0; i < lp->nCol; ++i ) {
            double cs = lp->colObj[i] * lp->colDualInverse[i];
            add[1] += cs; add[0] += cs * cs;
            lp->colBuffer[i] = lp->colObj[i] * lp->colDualInverse[i] * lp->colDualInverse[i];
        }
        for ( int r = 0; r < lp->nRow; ++r )
            for ( int e = lp->rowMatBeg[r]; e < lp->rowMatBeg[r + 1]; ++e ) kc->vecC[r] += lp->rowMatElem[e] * lp->colBuffer[lp->rowMatIdx[e]];
        if ( hdsdpcu_kkt_addhost(kc->dkkt, NULL, NULL, NULL, kc->vecC, add) != 0 ) return HDSDP_RETCODE_FAILED;
    }
    return HDSDP_RETCODE_OK;
}

/* bound cone: interface/hdsdp_conic_bound.c:201-249 restated on host vectors, written through addhost */
static hdsdp_retcode build_bound_cone( hdsdp_kkt *HKKT, kkt_cuda *kc, hdsdp_cone *cone, int typeKKT ) {
    hdsdp_cone_bound_scalar *b = (hdsdp_cone_bound_scalar *) cone->coneData;
    if ( typeKKT == KKT_TYPE_PRIMAL ) return HDSDP_RETCODE_FAILED;
    int m = HKKT->nRow;
    double add[4] = {0, 0, 0, 0};
    int hsd = ( typeKKT == KKT_TYPE_HOMOGENEOUS );
    for ( int i = 0; i < b->nRow; ++i ) {
        b->dualLowerInverse[i] = 1.0 / b->dualLower[i];
        b->dualUpperInverse[i] = 1.0 / b->dualUpper[i];
    }
    for ( int i = 0; i < m; ++i ) {
        double li = b->dualLowerInverse[i], ui = b->dualUpperInverse[i];
        kc->vecA[i] = ui - li;                 /* dASinvVec increment */
        kc->vecB[i] = li * li + ui * ui;       /* diag(M) increment */
        if ( hsd ) {
            kc->vecC[i] = b->dBoundUp * ui * ui + b->dBoundLow * li * li;
            add[1] += b->dBoundUp * ui - b->dBoundLow * li;
            add[0] += b->dBoundUp * b->dBoundUp * ui * ui + b->dBoundLow * b->dBoundLow * li * li;
        }
    }
    int rc = hdsdpcu_kkt_addhost(kc->dkkt, typeKKT == KKT_TYPE_CORRECTOR ? NULL : kc->vecB, kc->vecA, NULL,
                                 hsd ? kc->vecC : NULL, hsd ? add : NULL);
    return rc == 0 ? HDSDP_RETCODE_OK : HDSDP_RETCODE_FAILED;
}

static hdsdp_retcode build_one( hdsdp_kkt *HKKT, kkt_cuda *kc, hdsdp_cone *cone, int iCone, int typeKKT ) {
    switch ( cone->cone ) {
        case HDSDP_CONETYPE_DENSE_SDP:
        case HDSDP_CONETYPE_SPARSE_SDP: return build_sdp_cone(HKKT, kc, iCone, typeKKT);
        case HDSDP_CONETYPE_LP:         return build_lp_cone(HKKT, kc, iCone, typeKKT);
        case HDSDP_CONETYPE_SCALAR_BOUND: return build_bound_cone(HKKT, kc, cone, typeKKT);
        default: return HDSDP_RETCODE_FAILED;
    }
}

extern hdsdp_retcode HKKTBuildUp( hdsdp_kkt *HKKT, int typeKKT ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return HDSDP_RETCODE_FAILED;
    if ( hdsdpcu_kkt_clean(kc->dkkt, typeKKT) != 0 ) return HDSDP_RETCODE_FAILED;
    for ( int iCone = 0; iCone < HKKT->nCones; ++iCone ) {
        hdsdp_retcode rc = build_one(HKKT, kc, HKKT->cones[iCone], iCone, typeKKT);
        if ( rc != HDSDP_RETCODE_OK ) return rc;
    }
    refresh_host_vectors(HKKT, kc, typeKKT);
    return HDSDP_RETCODE_OK;
}

extern hdsdp_retcode HKKTBuildUpExtraCone( hdsdp_kkt *HKKT, hdsdp_cone *cone, int typeKKT ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return HDSDP_RETCODE_FAILED;
    hdsdp_retcode rc = build_one(HKKT, kc, cone, cone->iCone, typeKKT);
    if ( rc != HDSDP_RETCODE_OK ) return rc;
    refresh_host_vectors(HKKT, kc, typeKKT);
    return HDSDP_RETCODE_OK;
}

extern hdsdp_retcode HKKTBuildUpFixed( hdsdp_kkt *HKKT, int typeKKT, int kktStrategy ) {
    (void) kktStrategy; /* M2..M5 are algebraically identical; the GPU build has one formula per class pair */
    return HKKTBuildUp(HKKT, typeKKT);
}

extern void HKKTExport( hdsdp_kkt *HKKT, double *dKKTASinvVec, double *dKKTASinvRdSinvVec, double *dKKTASinvCSinvVec,
                        double *dCSinvCSinv, double *dCSinv, double *dCSinvRdCSinv, double *dTraceSinv ) {
    if ( dKKTASinvVec ) memcpy(dKKTASinvVec, HKKT->dASinvVec, sizeof(double) * HKKT->nRow);
    if ( dKKTASinvRdSinvVec ) memcpy(dKKTASinvRdSinvVec, HKKT->dASinvRdSinvVec, sizeof(double) * HKKT->nRow);
    if ( dKKTASinvCSinvVec ) memcpy(dKKTASinvCSinvVec, HKKT->dASinvCSinvVec, sizeof(double) * HKKT->nRow);
    if ( dCSinvCSinv ) *dCSinvCSinv = HKKT->dCSinvCSinv;
    if ( dCSinv ) *dCSinv = HKKT->dCSinv;
    if ( dCSinvRdCSinv ) *dCSinvRdCSinv = HKKT->dCSinvRdSinv;
    if ( dTraceSinv ) *dTraceSinv = HKKT->dTraceSinv;
}

extern hdsdp_retcode HKKTFactorize( hdsdp_kkt *HKKT ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return HDSDP_RETCODE_FAILED;
    return hdsdpcu_kkt_factorize(kc->dkkt) == 0 ? HDSDP_RETCODE_OK : HDSDP_RETCODE_FAILED;
}

extern hdsdp_retcode HKKTSolve( hdsdp_kkt *HKKT, double *dRhsVec, double *dLhsVec ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return HDSDP_RETCODE_FAILED;
    return hdsdpcu_kkt_solve(kc->dkkt, dRhsVec, dLhsVec) == 0 ? HDSDP_RETCODE_OK : HDSDP_RETCODE_FAILED;
}

extern void HKKTRegularize( hdsdp_kkt *HKKT, double dKKTReg ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( kc ) hdsdpcu_kkt_regularize(kc->dkkt, dKKTReg);
}

extern void HKKTRegisterPSDP( hdsdp_kkt *HKKT, double **dPrimalX ) {
    kkt_cuda *kc = lookup(HKKT);
    HKKT->dPrimalX = dPrimalX;
    if ( kc ) hdsdpcu_kkt_registerpsdp(kc->dkkt, HKKT->nCones, dPrimalX);
}

extern void HKKTClear( hdsdp_kkt *HKKT ) {
    if ( !HKKT ) return;
    kkt_cuda *kc = lookup(HKKT);
    if ( kc ) {
        for ( int i = 0; i < kc->nCones; ++i ) {
            if ( !kc->dcone[i] ) continue;
            if ( HKKT->cones && HKKT->cones[i] && HKKT->cones[i]->cone == HDSDP_CONETYPE_LP ) hdsdpcu_lp_destroy(&kc->dcone[i]);
            else hdsdpcu_cone_destroy(&kc->dcone[i]);
        }
        hdsdpcu_kkt_destroy(&kc->dkkt);
        free(kc->dcone); free(kc->objScal); free(kc->scaled); free(kc->hostInv);
        free(kc->vecA); free(kc->vecB); free(kc->vecC); free(kc->vecD); free(kc);
        for ( int i = 0; i < MAX_KKT; ++i ) if ( g_keys[i] == HKKT ) { g_keys[i] = NULL; g_vals[i] = NULL; }
    }
    free(HKKT->dASinvVec); free(HKKT->dASinvCSinvVec); free(HKKT->dASinvRdSinvVec);
    free(HKKT->invBuffer); free(HKKT->kktBuffer); free(HKKT->kktBuffer2);
    memset(HKKT, 0, sizeof(hdsdp_kkt));
}

extern void HKKTDestroy( hdsdp_kkt **pHKKT ) {
    if ( !pHKKT || !*pHKKT ) return;
    HKKTClear(*pHKKT);
    free(*pHKKT);
    *pHKKT = NULL;
}
