/*
 * integration/hdsdpcu_shim.h -- glue shared by the three hook files of the drop-in build
 *   hdsdp_conic_cuda.c   (cone vtable, interface/hdsdp_conic.c:54-200)
 *   hdsdp_schur_cuda.c   (HKKT*, interface/hdsdp_schur.c)
 *   hdsdp_linsys_cuda.c  (HFpLinsysCreate, linalg/hdsdp_linsolver.c:1859)
 * Host-only C; nothing here is part of the libhdsdp_cuda.so ABI (include/hdsdpcu.h).
 */
#ifndef HDSDPCU_SHIM_H
#define HDSDPCU_SHIM_H

#include "interface/def_hdsdp_conic.h"

/* device image (hdsdpcu cone handle) of an SDP cone whose vtable was bound by hdsdp_conic_cuda.c; NULL otherwise */
void *hdsdpcu_shim_cone_handle( hdsdp_cone *cone );
/* device Schur object behind a hdsdp_kkt (hdsdp_schur_cuda.c); NULL if unknown */
void *hdsdpcu_shim_kkt_handle( void *hkkt );

/* ---- accounting: wall time of every hot-path call next to the device time inside it -------------------------------
 * One bracket = one call from the reference's IPM driver into a hook.  wall = CLOCK_MONOTONIC around the whole hook,
 * device = CUDA events on the library stream around the same region (hdsdpcu_timer_start/stop), hostArith = wall time the
 * hook itself spent in host loops (bound cone, LP slack inversion); reported when the Schur object is destroyed. */
enum {
    SHIM_CAT_SFORM = 0,   /* S / dS assembly + Cholesky(S): coneUpdate, interior checks, barrier, add-step-and-check */
    SHIM_CAT_RATIO,       /* coneRatioTest: dS assembly + device Lanczos */
    SHIM_CAT_SCHUR,       /* HKKTBuildUp*: S^-1 + Schur assembly + side vectors */
    SHIM_CAT_FACTOR,      /* HKKTFactorize: Cholesky(M) */
    SHIM_CAT_SOLVE,       /* HKKTSolve */
    SHIM_CAT_PRIMAL,      /* conePRecover, coneBuildPrimalDirection, coneXDotS, coneDRecover, DIMACS eigenvalue */
    SHIM_CAT_LINSYS,      /* B1 calls with host matrices (primal X factors of the PSDP refinement) */
    SHIM_NCAT
};
void shim_prof_begin( int cat );
void shim_prof_end( int cat );
void shim_prof_host_begin( void );   /* host arithmetic inside an open bracket */
void shim_prof_host_end( int cat );
void shim_prof_report( void );       /* prints once; later calls print only if something new was recorded */

#endif /* HDSDPCU_SHIM_H */
