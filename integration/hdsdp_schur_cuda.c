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
 * The SDP cones themselves live on the device (integration/hdsdp_conic_cuda.c binds the cone vtable), so S^-1 is computed
 * from the device-resident Cholesky factor of S inside hdsdpcu_cone_buildschur: no dual matrix, factor or inverse ever crosses
 * PCIe.  A cone that was not bound by the cone hook is an error (there is no host path for S).
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
#include "hdsdpcu_shim.h"

typedef struct {
    void *dkkt;              /* hdsdpcu kkt handle */
    int nCones;
    void **dcone;            /* per cone: hdsdpcu cone handle (SDP, owned by the cone hook), lp handle (LP, owned here) or NULL */
    int *ownsLp;             /* per cone: dcone[i] is an LP image created (and to be destroyed) here */
    double *vecA, *vecB, *vecC, *vecD;
} kkt_cuda;

/* side table keyed by the hdsdp_kkt pointer (the struct layout must stay the reference's); grows on demand */
static hdsdp_kkt **g_keys = NULL;
static kkt_cuda **g_vals = NULL;
static int g_cap = 0;

static kkt_cuda *lookup(hdsdp_kkt *k) {
    for (int i = 0; i < g_cap; ++i) if (g_keys[i] == k) return g_vals[i];
    return NULL;
}

static int table_insert(hdsdp_kkt *k, kkt_cuda *v) {
    for (int i = 0; i < g_cap; ++i) if (!g_keys[i]) { g_keys[i] = k; g_vals[i] = v; return 0; }
    int ncap = g_cap ? 2 * g_cap : 8;
    hdsdp_kkt **nk = (hdsdp_kkt **) realloc(g_keys, sizeof(hdsdp_kkt *) * ncap);
    if ( !nk ) return 1;
    g_keys = nk;
    kkt_cuda **nv = (kkt_cuda **) realloc(g_vals, sizeof(kkt_cuda *) * ncap);
    if ( !nv ) return 1;
    g_vals = nv;
    for (int i = g_cap; i < ncap; ++i) { g_keys[i] = NULL; g_vals[i] = NULL; }
    g_keys[g_cap] = k; g_vals[g_cap] = v;
    g_cap = ncap;
    return 0;
}

void *hdsdpcu_shim_kkt_handle(void *hkkt) {
    kkt_cuda *kc = lookup((hdsdp_kkt *) hkkt);
    return kc ? kc->dkkt : NULL;
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
    if ( !kc ) return HDSDP_RETCODE_MEMORY;
    kc->nCones = nCones;
    kc->dcone = (void **) calloc(nCones > 0 ? nCones : 1, sizeof(void *));
    kc->ownsLp = (int *) calloc(nCones > 0 ? nCones : 1, sizeof(int));
    kc->vecA = (double *) calloc(nRow, sizeof(double)); kc->vecB = (double *) calloc(nRow, sizeof(double));
    kc->vecC = (double *) calloc(nRow, sizeof(double)); kc->vecD = (double *) calloc(nRow, sizeof(double));
    if ( !kc->dcone || !kc->ownsLp || !kc->vecA || !kc->vecB || !kc->vecC || !kc->vecD || table_insert(HKKT, kc) != 0 ) {
        free(kc->dcone); free(kc->ownsLp); free(kc->vecA); free(kc->vecB); free(kc->vecC); free(kc->vecD); free(kc);
        return HDSDP_RETCODE_MEMORY;
    }
    int drc = hdsdpcu_kkt_create(&kc->dkkt, nRow);
    if ( drc != 0 ) return drc == 2 ? HDSDP_RETCODE_MEMORY : HDSDP_RETCODE_FAILED;
    /* HDSDPCU_KKT_SOLVER=pcg: the reference's own policy for M (hdsdp_schur.c:19-35: Jacobi-PCG first, Cholesky after its first
       failure) on the device instead of the direct Cholesky factorisation */
    const char *solver = getenv("HDSDPCU_KKT_SOLVER");
    if ( solver && strcmp(solver, "pcg") == 0 && hdsdpcu_kkt_set_solver(kc->dkkt, 1) != 0 ) return HDSDP_RETCODE_FAILED;
    for ( int iCone = 0; iCone < nCones; ++iCone ) {
        hdsdp_cone *c = cones[iCone];
        user_data *u = (user_data *) c->usrData;
        if ( c->cone == HDSDP_CONETYPE_DENSE_SDP || c->cone == HDSDP_CONETYPE_SPARSE_SDP ) {
            kc->dcone[iCone] = hdsdpcu_shim_cone_handle(c);   /* created by the cone hook in coneProcData */
            if ( !kc->dcone[iCone] ) {
                printf("[hdsdpcu] SDP cone %d has no device image (cone hook not linked?); there is no host path for S\n", iCone);
                return HDSDP_RETCODE_FAILED;
            }
        } else if ( c->cone == HDSDP_CONETYPE_LP ) {
            if ( hdsdpcu_lp_create(&kc->dcone[iCone], nRow, u->nConicCol, u->coneMatBeg, u->coneMatIdx, u->coneMatElem) != 0 )
                return HDSDP_RETCODE_FAILED;
            kc->ownsLp[iCone] = 1;
        } else {
            printf("[hdsdpcu] unsupported cone type %d in HKKTInit\n", (int) c->cone);
            return HDSDP_RETCODE_FAILED;
        }
    }
    printf("    Using dense Schur complement on the GPU (libhdsdp_cuda); SDP cones are device resident\n");
    return HDSDP_RETCODE_OK;
}

static hdsdp_retcode build_sdp_cone( hdsdp_kkt *HKKT, kkt_cuda *kc, int iCone, int typeKKT ) {
    /* S^-1 = inverse of the device-resident dual factor (hdsdp_conic_sdp.c:1755 HFpLinsysInvert), computed inside
       hdsdpcu_cone_buildschur and cached until S is factorised again (the corrector builds of one iteration reuse it) */
    (void) HKKT;
    return (hdsdp_retcode) hdsdpcu_cone_buildschur(kc->dcone[iCone], iCone, kc->dkkt, typeKKT);
}

/* LP cone: the O(nLpCol) slack inversion stays on the host (hdsdp_conic_lp.c:262-271); everything that touches M or the
   side vectors runs on the device */
static hdsdp_retcode build_lp_cone( hdsdp_kkt *HKKT, kkt_cuda *kc, int iCone, int typeKKT ) {
    hdsdp_cone_lp *lp = (hdsdp_cone_lp *) HKKT->cones[iCone]->coneData;
    shim_prof_host_begin();
    if ( typeKKT == KKT_TYPE_PRIMAL ) {
        if ( !HKKT->dPrimalX || !HKKT->dPrimalX[iCone] ) return HDSDP_RETCODE_FAILED;
        for ( int i = 0; i < lp->nCol; ++i ) lp->colDualInverse[i] = HKKT->dPrimalX[iCone][i];
    } else {
        for ( int i = 0; i < lp->nCol; ++i ) lp->colDualInverse[i] = 1.0 / lp->colDual[i];
    }
    shim_prof_host_end(SHIM_CAT_SCHUR);
    /* the reference rescales colObj in place (LPConeScal): hand the current objective to the device image for the HSD terms */
    if ( typeKKT == KKT_TYPE_HOMOGENEOUS && hdsdpcu_lp_setobjective(kc->dcone[iCone], lp->colObj) != 0 ) return HDSDP_RETCODE_FAILED;
    if ( hdsdpcu_kkt_buildupextra_lp(kc->dkkt, kc->dcone[iCone], lp->colDualInverse, lp->dualResidual, typeKKT) != 0 )
        return HDSDP_RETCODE_FAILED;
    /* the HOMOGENEOUS terms of hdsdp_conic_lp.c:316-327 (dCSinv, dCSinvCSinv, A C s^-2) are part of the device kernel */
    return HDSDP_RETCODE_OK;
}

/* bound cone: interface/hdsdp_conic_bound.c:201-249 restated on host vectors, written through addhost */
static hdsdp_retcode build_bound_cone( hdsdp_kkt *HKKT, kkt_cuda *kc, hdsdp_cone *cone, int typeKKT ) {
    hdsdp_cone_bound_scalar *b = (hdsdp_cone_bound_scalar *) cone->coneData;
    if ( typeKKT == KKT_TYPE_PRIMAL ) return HDSDP_RETCODE_FAILED;
    int m = HKKT->nRow;
    double add[4] = {0, 0, 0, 0};
    int hsd = ( typeKKT == KKT_TYPE_HOMOGENEOUS );
    shim_prof_host_begin();
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
    shim_prof_host_end(SHIM_CAT_SCHUR);
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
    hdsdp_retcode rc = HDSDP_RETCODE_OK;
    shim_prof_begin(SHIM_CAT_SCHUR);
    if ( hdsdpcu_kkt_clean(kc->dkkt, typeKKT) != 0 ) rc = HDSDP_RETCODE_FAILED;
    for ( int iCone = 0; iCone < HKKT->nCones && rc == HDSDP_RETCODE_OK; ++iCone )
        rc = build_one(HKKT, kc, HKKT->cones[iCone], iCone, typeKKT);
    if ( rc == HDSDP_RETCODE_OK ) refresh_host_vectors(HKKT, kc, typeKKT);
    shim_prof_end(SHIM_CAT_SCHUR);
    return rc;
}

extern hdsdp_retcode HKKTBuildUpExtraCone( hdsdp_kkt *HKKT, hdsdp_cone *cone, int typeKKT ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return HDSDP_RETCODE_FAILED;
    shim_prof_begin(SHIM_CAT_SCHUR);
    hdsdp_retcode rc = build_one(HKKT, kc, cone, cone->iCone, typeKKT);
    if ( rc == HDSDP_RETCODE_OK ) refresh_host_vectors(HKKT, kc, typeKKT);
    shim_prof_end(SHIM_CAT_SCHUR);
    return rc;
}

extern hdsdp_retcode HKKTBuildUpFixed( hdsdp_kkt *HKKT, int typeKKT, int kktStrategy ) {
    /* M2..M5 are algebraically identical; the GPU build has ONE formula per class pair, so the reference's own
       M3-vs-M4 cross-check (HUtilKKTCheck, hdsdp_utils.c:536-707) compares the device result with itself through this
       hook -- the cross-formula check lives in tests/test_cpu_oracle.py and tests/test_gpu_schur.py instead */
    (void) kktStrategy;
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
    shim_prof_begin(SHIM_CAT_FACTOR);
    int rc = hdsdpcu_kkt_factorize(kc->dkkt);
    shim_prof_end(SHIM_CAT_FACTOR);
    return rc == 0 ? HDSDP_RETCODE_OK : HDSDP_RETCODE_FAILED;
}

extern hdsdp_retcode HKKTSolve( hdsdp_kkt *HKKT, double *dRhsVec, double *dLhsVec ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return HDSDP_RETCODE_FAILED;
    shim_prof_begin(SHIM_CAT_SOLVE);
    int rc = hdsdpcu_kkt_solve(kc->dkkt, dRhsVec, dLhsVec);
    shim_prof_end(SHIM_CAT_SOLVE);
    return rc == 0 ? HDSDP_RETCODE_OK : HDSDP_RETCODE_FAILED;
}

extern void HKKTRegularize( hdsdp_kkt *HKKT, double dKKTReg ) {
    kkt_cuda *kc = lookup(HKKT);
    if ( !kc ) return;
    shim_prof_begin(SHIM_CAT_SCHUR);
    hdsdpcu_kkt_regularize(kc->dkkt, dKKTReg);
    shim_prof_end(SHIM_CAT_SCHUR);
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
            /* SDP cone images belong to the cone hook (destroyed by coneDestroyData).  HDSDPClear (interface/hdsdp.c:943-950)
               frees the cones and their array BEFORE the Schur object, so HKKT->cones must not be touched here. */
            if ( kc->ownsLp[i] ) hdsdpcu_lp_destroy(&kc->dcone[i]);
        }
        hdsdpcu_kkt_destroy(&kc->dkkt);
        free(kc->dcone); free(kc->ownsLp);
        free(kc->vecA); free(kc->vecB); free(kc->vecC); free(kc->vecD); free(kc);
        for ( int i = 0; i < g_cap; ++i ) if ( g_keys[i] == HKKT ) { g_keys[i] = NULL; g_vals[i] = NULL; }
        shim_prof_report();
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
