/*
 * integration/hdsdp_conic_cuda.c -- the cone hook: the SDP cone vtable of the reference bound to the device.
 *
 * In the reference HConeSetData (interface/hdsdp_conic.c:54-200) fills the ~30 function pointers of hdsdp_cone
 * (interface/def_hdsdp_conic.h:56-107) per cone type.  A maintainer changes the two SDP cases there; for the patch-free
 * integration build the reference object is linked with
 *     objcopy --redefine-sym HConeSetData=HConeSetData_ref
 * and this file provides HConeSetData: it lets the reference fill the table and then re-points, for
 * HDSDP_CONETYPE_DENSE_SDP and HDSDP_CONETYPE_SPARSE_SDP, every member that touches the dual slack S:
 *
 *   coneUpdate                sdpDenseConeUpdateImpl                  hdsdp_conic_sdp.c:1616  -> hdsdpcu_cone_update
 *   coneRatioTest             sdpDenseConeRatioTestImpl               :1642                   -> hdsdpcu_cone_ratiotest
 *   coneInteriorCheck         sdpDenseConeInteriorCheck               :2172                   -> hdsdpcu_cone_interiorcheck
 *   coneInteriorCheckExpert   sdpDenseConeInteriorCheckExpert         :2192                   -> hdsdpcu_cone_interiorcheckexpert
 *   coneGetBarrier            sdpDenseConeGetBarrier                  :2252                   -> hdsdpcu_cone_getbarrier
 *   coneAxpyBufferAndCheck    sdpDenseConeAddStepToBufferAndCheck     :2333                   -> hdsdpcu_cone_addstepandcheck
 *   conePRecover              sdpDenseConeGetPrimal                   :2395                   -> hdsdpcu_cone_getprimal
 *   coneDRecover              sdpDenseConeGetDual                     :2497                   -> hdsdpcu_cone_getdual
 *   coneXDotS                 sdpDenseConeXDotS                       :2549                   -> hdsdpcu_cone_xdots
 *   coneBuildPrimalDirection  sdpDenseConeBuildPrimalXSXDirection     :2021                   -> hdsdpcu_cone_buildprimalxsx
 *   coneBuildSchur(/Fixed)    sdpDenseConeGetKKT(/ByFixedStrategy)    :1726 / :1888           -> hdsdpcu_cone_buildschur
 *   coneSetStart / coneReduceResi / coneSetPerturb / coneScal (:1546, :2226, :2238, :1604): reference + device image
 *   coneCreate / coneProcData / coneDestroyData (:1323, :1356, :2633): reference + hdsdpcu_cone_create / _destroy
 *
 * S, the checker copy, dS, their Cholesky factors, S^-1 and the Lanczos basis live in HBM only; per call the PCIe traffic is
 * y or dy (m doubles) down and one scalar (PSD flag, log det, step) up.  The dual matrices are ALWAYS dense on the device,
 * also where the reference would pick a sparse S with QDLDL (max-cut): the reference's own sparse dual structures stay
 * allocated but are never touched again.  The data-only members (coneTraceCX, coneATimesXpy, norms, feature detection,
 * KKT ordering) keep the reference's host implementations: they read the coefficient matrices, never S.
 *
 * The DIMACS check's eigenvalue (HDSDPCheckSolution, interface/hdsdp.c:852-861: fds_syev(n, X, d, Y, 1, ...), i.e. ONE
 * extreme eigenvalue of the primal block) is routed to hdsdpcu_sym_extreme_eig by linking hdsdp.o with
 *     objcopy --redefine-sym fds_syev=fds_syev_dimacs
 * (the Lanczos code keeps the reference's LAPACK fds_syev for its 30 x 30 tridiagonal problems).
 *
 * Original code written against the reference's headers.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "interface/hdsdp_utils.h"
#include "interface/def_hdsdp_user_data.h"
#include "interface/def_hdsdp_conic.h"
#include "interface/hdsdp_conic.h"
#include "interface/hdsdp_conic_sdp.h"
#include "linalg/dense_opts.h"
#include "hdsdpcu.h"
#include "hdsdpcu_shim.h"

extern hdsdp_retcode HConeSetData_ref( hdsdp_cone *HCone, user_data *usrData );

/* The reference functions we keep see their own struct at offset 0; the device handle rides behind it. */
typedef struct {
    union { hdsdp_cone_sdp_dense dense; hdsdp_cone_sdp_sparse sparse; } ref;
    void *dev;              /* hdsdpcu cone handle */
    int sparseType;         /* 1: the reference chose hdsdp_cone_sdp_sparse */
    int nCol;
    int hostBuilt;          /* the reference's coneProcData completed (its clean-up is only safe on a complete cone) */
} cone_cu;

#define DEV(c) (((cone_cu *) (c))->dev)
#define RC(x) ((x) == 0 ? HDSDP_RETCODE_OK : HDSDP_RETCODE_FAILED)

/* ---- life cycle ------------------------------------------------------------------------------------------------- */
static hdsdp_retcode cu_create_dense( void **pCone ) {
    if ( !pCone ) return HDSDP_RETCODE_FAILED;
    cone_cu *c = (cone_cu *) calloc(1, sizeof(cone_cu));
    if ( !c ) return HDSDP_RETCODE_MEMORY;
    *pCone = c;
    return HDSDP_RETCODE_OK;
}
static hdsdp_retcode cu_create_sparse( void **pCone ) {
    hdsdp_retcode rc = cu_create_dense(pCone);
    if ( rc == HDSDP_RETCODE_OK ) ((cone_cu *) *pCone)->sparseType = 1;
    return rc;
}

static hdsdp_retcode cu_procdata( void *cone, int nRow, int nCol, int *beg, int *idx, double *elem ) {
    cone_cu *c = (cone_cu *) cone;
    if ( hdsdpcu_device_count() <= 0 ) {
        printf("[hdsdpcu] no CUDA device: the SDP cone cannot be set up (there is no CPU fallback for the hot path)\n");
        return HDSDP_RETCODE_FAILED;
    }
    /* host image: coefficient objects for the data-only members (norms, tr(CX), A(X), features, KKT ordering) */
    hdsdp_retcode rc = c->sparseType ? sdpSparseConeProcDataImpl(&c->ref.sparse, nRow, nCol, beg, idx, elem)
                                     : sdpDenseConeProcDataImpl(&c->ref.dense, nRow, nCol, beg, idx, elem);
    if ( rc != HDSDP_RETCODE_OK ) return rc;
    c->hostBuilt = 1;
    c->nCol = nCol;
    /* device image: classification + upload once (hdsdp_conic_sdp.c:1356-1400, :1490-1518 restated in csrc/classify.cpp) */
    int drc = hdsdpcu_cone_create(&c->dev, nRow, nCol, beg, idx, elem);
    if ( drc != 0 ) {
        printf("[hdsdpcu] cone of dimension %d could not be created on the device (retcode %d); there is no CPU fallback\n", nCol, drc);
        return drc == 2 ? HDSDP_RETCODE_MEMORY : HDSDP_RETCODE_FAILED;
    }
    return HDSDP_RETCODE_OK;
}

static void cu_destroy( void **pCone ) {
    if ( !pCone || !*pCone ) return;
    cone_cu *c = (cone_cu *) *pCone;
    if ( c->dev ) hdsdpcu_cone_destroy(&c->dev);
    if ( c->hostBuilt ) {
        if ( c->sparseType ) sdpSparseConeClearImpl(&c->ref.sparse); else sdpDenseConeClearImpl(&c->ref.dense);
    }
    free(c);
    *pCone = NULL;
}

/* ---- scalars kept in both images ------------------------------------------------------------------------------------ */
static void cu_setstart( void *cone, double rResi ) {
    ((cone_cu *) cone)->ref.dense.dualResidual = rResi;      /* same offset in both reference structs */
    hdsdpcu_cone_setstart(DEV(cone), rResi);
}
static void cu_reduceresi( void *cone, double resi ) {
    ((cone_cu *) cone)->ref.dense.dualResidual = resi;
    hdsdpcu_cone_reduceresi(DEV(cone), resi);
}
static void cu_setperturb( void *cone, double dPerturb ) {
    ((cone_cu *) cone)->ref.dense.dualPerturb = dPerturb;
    hdsdpcu_cone_setperturb(DEV(cone), dPerturb);
}
static void cu_scal( void *cone, double dScal ) {
    cone_cu *c = (cone_cu *) cone;
    if ( c->sparseType ) sdpSparseConeScal(&c->ref.sparse, dScal); else sdpDenseConeScal(&c->ref.dense, dScal);
    if ( hdsdpcu_cone_scal(c->dev, dScal) != 0 ) printf("[hdsdpcu] coneScal failed on the device\n");
}

/* ---- S-side work ------------------------------------------------------------------------------------------------------ */
static void cu_update( void *cone, double barHsdTau, double *rowDual ) {
    shim_prof_begin(SHIM_CAT_SFORM);
    if ( hdsdpcu_cone_update(DEV(cone), barHsdTau, rowDual) != 0 ) printf("[hdsdpcu] coneUpdate failed on the device\n");
    shim_prof_end(SHIM_CAT_SFORM);
}

static hdsdp_retcode cu_ratiotest( void *cone, double barHsdTauStep, double *rowDualStep, double dAdaRatio, int whichBuffer,
                                   double *maxStep ) {
    shim_prof_begin(SHIM_CAT_RATIO);
    int rc = hdsdpcu_cone_ratiotest(DEV(cone), barHsdTauStep, rowDualStep, dAdaRatio, whichBuffer, maxStep);
    shim_prof_end(SHIM_CAT_RATIO);
    return RC(rc);
}

static hdsdp_retcode cu_interiorcheck( void *cone, double barHsdTau, double *rowDual, int *isInterior ) {
    shim_prof_begin(SHIM_CAT_SFORM);
    int rc = hdsdpcu_cone_interiorcheck(DEV(cone), barHsdTau, rowDual, isInterior);
    shim_prof_end(SHIM_CAT_SFORM);
    return RC(rc);
}

static hdsdp_retcode cu_interiorcheckexpert( void *cone, double dCCoef, double dACoefScal, double *dACoef, double dEyeCoef,
                                             int whichBuffer, int *isInterior ) {
    shim_prof_begin(SHIM_CAT_SFORM);
    int rc = hdsdpcu_cone_interiorcheckexpert(DEV(cone), dCCoef, dACoefScal, dACoef, dEyeCoef, whichBuffer, isInterior);
    shim_prof_end(SHIM_CAT_SFORM);
    return RC(rc);
}

static hdsdp_retcode cu_getbarrier( void *cone, double barHsdTau, double *rowDual, int whichBuffer, double *logdet ) {
    shim_prof_begin(SHIM_CAT_SFORM);
    int rc = hdsdpcu_cone_getbarrier(DEV(cone), barHsdTau, rowDual, whichBuffer, logdet);
    shim_prof_end(SHIM_CAT_SFORM);
    return RC(rc);
}

static hdsdp_retcode cu_axpybufferandcheck( void *cone, double dStep, int whichBuffer, int *isInterior ) {
    shim_prof_begin(SHIM_CAT_SFORM);
    int rc = hdsdpcu_cone_addstepandcheck(DEV(cone), dStep, whichBuffer, isInterior);
    shim_prof_end(SHIM_CAT_SFORM);
    return RC(rc);
}

/* ---- Schur complement: only reached if something calls the vtable directly (HKKTBuildUp of hdsdp_schur_cuda.c goes to
 * hdsdpcu_cone_buildschur itself) --------------------------------------------------------------------------------- */
static hdsdp_retcode cu_buildschur( void *cone, int iCone, void *kkt, int typeKKT ) {
    void *dkkt = hdsdpcu_shim_kkt_handle(kkt);
    if ( !dkkt ) return HDSDP_RETCODE_FAILED;
    return RC(hdsdpcu_cone_buildschur(DEV(cone), iCone, dkkt, typeKKT));
}
static hdsdp_retcode cu_buildschurfixed( void *cone, int iCone, void *kkt, int typeKKT, int kktStrategy ) {
    (void) kktStrategy;   /* M2..M5 are one algebraic quantity; the device picks a formula per class pair */
    return cu_buildschur(cone, iCone, kkt, typeKKT);
}

/* ---- primal side --------------------------------------------------------------------------------------------------- */
static void cu_buildprimaldirection( void *cone, void *kkt, double *dPrimalScalMatrix, double *dPrimalXSXBuffer, int iDualMat ) {
    (void) kkt;
    shim_prof_begin(SHIM_CAT_PRIMAL);
    if ( hdsdpcu_cone_buildprimalxsx(DEV(cone), dPrimalScalMatrix, dPrimalXSXBuffer, iDualMat) != 0 )
        printf("[hdsdpcu] coneBuildPrimalDirection failed on the device\n");
    shim_prof_end(SHIM_CAT_PRIMAL);
}

static void cu_precover( void *cone, double dBarrierMu, double *dRowDual, double *dRowDualStep, double *dConePrimal, double *dAuxiMat ) {
    (void) dAuxiMat;
    int feasible = 0;
    shim_prof_begin(SHIM_CAT_PRIMAL);
    int rc = hdsdpcu_cone_getprimal(DEV(cone), dBarrierMu, dRowDual, dRowDualStep, dConePrimal, &feasible);
    shim_prof_end(SHIM_CAT_PRIMAL);
    if ( rc != 0 ) printf("[hdsdpcu] conePRecover failed on the device\n");
    else if ( !feasible ) printf("Recovery step is infeasible\n");   /* hdsdp_conic_sdp.c:2400-2403: nothing is written */
}

static void cu_drecover( void *cone, double *dConeDual, double *dAuxi ) {
    (void) dAuxi;
    shim_prof_begin(SHIM_CAT_PRIMAL);
    if ( hdsdpcu_cone_getdual(DEV(cone), dConeDual) != 0 ) printf("[hdsdpcu] coneDRecover failed on the device\n");
    shim_prof_end(SHIM_CAT_PRIMAL);
}

static double cu_xdots( void *cone, double *dConePrimal ) {
    double v = HDSDP_INFINITY;
    shim_prof_begin(SHIM_CAT_PRIMAL);
    if ( hdsdpcu_cone_xdots(DEV(cone), dConePrimal, &v) != 0 ) v = HDSDP_INFINITY;
    shim_prof_end(SHIM_CAT_PRIMAL);
    return v;
}

void *hdsdpcu_shim_cone_handle( hdsdp_cone *HCone ) {
    if ( !HCone || !HCone->coneData || HCone->coneUpdate != cu_update ) return NULL;
    return DEV(HCone->coneData);
}

extern hdsdp_retcode HConeSetData( hdsdp_cone *HCone, user_data *usrData ) {
    hdsdp_retcode rc = HConeSetData_ref(HCone, usrData);
    if ( rc != HDSDP_RETCODE_OK ) return rc;
    if ( HCone->cone != HDSDP_CONETYPE_DENSE_SDP && HCone->cone != HDSDP_CONETYPE_SPARSE_SDP ) return rc;
    HCone->coneCreate = ( HCone->cone == HDSDP_CONETYPE_SPARSE_SDP ) ? cu_create_sparse : cu_create_dense;
    HCone->coneProcData = cu_procdata;
    HCone->coneDestroyData = cu_destroy;
    HCone->coneSetStart = cu_setstart;
    HCone->coneReduceResi = cu_reduceresi;
    HCone->coneSetPerturb = cu_setperturb;
    HCone->coneScal = cu_scal;
    HCone->coneUpdate = cu_update;
    HCone->coneRatioTest = cu_ratiotest;
    HCone->coneInteriorCheck = cu_interiorcheck;
    HCone->coneInteriorCheckExpert = cu_interiorcheckexpert;
    HCone->coneGetBarrier = cu_getbarrier;
    HCone->coneAxpyBufferAndCheck = cu_axpybufferandcheck;
    HCone->coneBuildSchur = cu_buildschur;
    HCone->coneBuildSchurFixed = cu_buildschurfixed;
    HCone->coneBuildPrimalDirection = cu_buildprimaldirection;
    HCone->conePRecover = cu_precover;
    HCone->coneDRecover = cu_drecover;
    HCone->coneXDotS = cu_xdots;
    return rc;
}

/* DIMACS check (interface/hdsdp.c:852-861): one extreme eigenvalue of the recovered primal block.  The reference asks
 * dsyevr for index range [n - m + 1, n] with m = 1 (linalg/dense_opts.c:56-78), i.e. the LARGEST eigenvalue; the device
 * Lanczos (csrc/lanczos.cu: sym_extreme_eig) returns the same number to 1e-12 ||X||.  Other m: the reference routine. */
extern hdsdp_retcode fds_syev_dimacs( int n, double *U, double *d, double *Y, int m, double *work, int *iwork, int lwork, int liwork ) {
    if ( m != 1 || n < 2 ) return fds_syev(n, U, d, Y, m, work, iwork, lwork, liwork);
    int steps = 0;
    shim_prof_begin(SHIM_CAT_PRIMAL);
    int rc = hdsdpcu_sym_extreme_eig(n, U, 1, &d[0], &steps);
    shim_prof_end(SHIM_CAT_PRIMAL);
    return RC(rc);
}
