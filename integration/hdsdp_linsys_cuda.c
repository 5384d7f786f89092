/*
 * integration/hdsdp_linsys_cuda.c -- the B1 hook: a CUDA back-end behind the reference's linear-system vtable.
 *
 * In the reference, HFpLinsysCreate (linalg/hdsdp_linsolver.c:1859) fills the 11 function pointers of
 * hdsdp_linsys_fp (linalg/def_hdsdp_linsolver.h:40-63) from a switch on linsys_type.  A maintainer adds one
 * case there; for the patch-free integration build the reference object is linked with
 *   objcopy --redefine-sym HFpLinsysCreate=HFpLinsysCreate_ref
 * and this file provides HFpLinsysCreate, routing HDSDP_LINSYS_DENSE_DIRECT (primal X factors of the PSDP refinement,
 * hdsdp_psdp.c:104; the dual factor / checker objects of a dense S, hdsdp_conic_sdp.c:109-110,208-209, are still created
 * through here but never used: the cone hook keeps S and its factors on the device) to libhdsdp_cuda.so -- at EVERY
 * dimension, there is no LAPACK fallback -- and everything else (sparse types the reference allocates for sparse S / sparse M,
 * which the GPU build never factorises) to the original constructor.
 */
#include <stdlib.h>
#include "linalg/hdsdp_linsolver.h"
#include "hdsdpcu.h"
#include "hdsdpcu_shim.h"

extern hdsdp_retcode HFpLinsysCreate_ref( hdsdp_linsys_fp **pHLin, int nCol, linsys_type Ltype );

static hdsdp_retcode cu_create( void **pchol, int nCol ) { return (hdsdp_retcode) hdsdpcu_linsys_create(pchol, nCol); }
static void cu_setparam( void *chol, void *param ) { hdsdpcu_linsys_setparam(chol, param); }
static hdsdp_retcode cu_symbolic( void *chol, int *b, int *i ) { return (hdsdp_retcode) hdsdpcu_linsys_symbolic(chol, b, i); }
static hdsdp_retcode cu_numeric( void *chol, int *b, int *i, double *e ) {
    shim_prof_begin(SHIM_CAT_LINSYS);
    int rc = hdsdpcu_linsys_numeric(chol, b, i, e);
    shim_prof_end(SHIM_CAT_LINSYS);
    return (hdsdp_retcode) rc;
}
static hdsdp_retcode cu_psdcheck( void *chol, int *b, int *i, double *e, int *p ) {
    shim_prof_begin(SHIM_CAT_LINSYS);
    int rc = hdsdpcu_linsys_psdcheck(chol, b, i, e, p);
    shim_prof_end(SHIM_CAT_LINSYS);
    return (hdsdp_retcode) rc;
}
static void cu_fsolve( void *chol, int n, double *r, double *s ) {
    shim_prof_begin(SHIM_CAT_LINSYS);
    hdsdpcu_linsys_fsolve(chol, n, r, s);
    shim_prof_end(SHIM_CAT_LINSYS);
}
static void cu_bsolve( void *chol, int n, double *r, double *s ) {
    shim_prof_begin(SHIM_CAT_LINSYS);
    hdsdpcu_linsys_bsolve(chol, n, r, s);
    shim_prof_end(SHIM_CAT_LINSYS);
}
static hdsdp_retcode cu_solve( void *chol, int n, double *r, double *s ) {
    shim_prof_begin(SHIM_CAT_LINSYS);
    int rc = hdsdpcu_linsys_solve(chol, n, r, s);
    shim_prof_end(SHIM_CAT_LINSYS);
    return (hdsdp_retcode) rc;
}
static hdsdp_retcode cu_getdiag( void *chol, double *d ) { return (hdsdp_retcode) hdsdpcu_linsys_getdiag(chol, d); }
static void cu_invert( void *chol, double *inv, double *aux ) {
    shim_prof_begin(SHIM_CAT_LINSYS);
    hdsdpcu_linsys_invert(chol, inv, aux);
    shim_prof_end(SHIM_CAT_LINSYS);
}
static void cu_destroy( void **pchol ) { hdsdpcu_linsys_destroy(pchol); }

int hdsdpcu_linsys_is_cuda( hdsdp_linsys_fp *lin ) { return lin && lin->cholNumeric == cu_numeric; }

extern hdsdp_retcode HFpLinsysCreate( hdsdp_linsys_fp **pHLin, int nCol, linsys_type Ltype ) {
    if ( Ltype != HDSDP_LINSYS_DENSE_DIRECT ) return HFpLinsysCreate_ref(pHLin, nCol, Ltype);
    if ( !pHLin ) return HDSDP_RETCODE_FAILED;
    hdsdp_linsys_fp *H = (hdsdp_linsys_fp *) calloc(1, sizeof(hdsdp_linsys_fp));
    if ( !H ) return HDSDP_RETCODE_MEMORY;
    H->LinType = Ltype; H->nCol = nCol;
    H->cholCreate = cu_create; H->cholSetParam = cu_setparam; H->cholSymbolic = cu_symbolic; H->cholNumeric = cu_numeric;
    H->cholPsdCheck = cu_psdcheck; H->cholFSolve = cu_fsolve; H->cholBSolve = cu_bsolve; H->cholSolve = cu_solve;
    H->cholGetDiag = cu_getdiag; H->cholInvert = cu_invert; H->cholDestroy = cu_destroy;
    hdsdp_retcode rc = H->cholCreate(&H->chol, nCol);
    if ( rc != HDSDP_RETCODE_OK ) { free(H); return rc; }
    *pHLin = H;
    return HDSDP_RETCODE_OK;
}
