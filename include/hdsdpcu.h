/*
 * include/hdsdpcu.h -- C ABI of libhdsdp_cuda.so, the B200 (sm_100a) implementation of HDSDP's
 * per-iteration Newton-system hot path (dual slack S, its Cholesky factor and inverse; the Schur
 * complement M_ij = <A_i, S^-1 A_j S^-1>; the dense FP64 Cholesky factorisation and solve of M).
 *
 * Plain C: opaque handles, pointers and sizes only.  Unless a name ends in _dev, every pointer is a
 * HOST pointer with exactly the meaning it has in the reference interface it replaces, so the ANSI-C
 * solver can bind these entry points directly (see INTEGRATION.md for the three hook sites).
 * Return values are the reference's hdsdp_retcode (interface/hdsdp.h:42-48): 0 OK, 1 FAILED, 2 MEMORY.
 * There is no CPU fallback: every compute entry point fails (1) when no CUDA device is usable.
 *
 * Dense matrices are column-major, full n x n storage with the LOWER triangle meaningful
 * (reference FULL_ENTRY, interface/hdsdp_utils.h:53).
 */
#ifndef HDSDPCU_H
#define HDSDPCU_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Runtime
 * ------------------------------------------------------------------------------------------- */
int hdsdpcu_init(int device);            /* select the device, create the library stream; 1 if no GPU */
int hdsdpcu_device_count(void);          /* 0 without a usable GPU (never touches the device otherwise) */
void *hdsdpcu_stream(void);              /* cudaStream_t all kernels are launched on (for CUDA-event timing) */
int hdsdpcu_sync(void);
const char *hdsdpcu_version(void);
/* count of kernel launches issued by this library since the last reset (bench.py "gpu_launches") */
long hdsdpcu_launch_count(int reset);
/* device-time bracket around calls of this library (two CUDA events on the library stream; stop synchronises and returns the
 * elapsed milliseconds).  Used by the host shims of integration/ to report how much of the hot path's wall time is device time. */
int hdsdpcu_timer_start(void);
int hdsdpcu_timer_stop(double *ms);
/* device-to-device copy on the library stream (bench plumbing for HBM-resident inputs) */
int hdsdpcu_copy_dev(void *d_dst, const void *d_src, long bytes);
/* debug: clock64() stamps of the phases of the last leaf factorisation (40 values, tools/leafclk.py); only filled when the
 * library is built with -DHDSDPCU_LEAFCLK */
int hdsdpcu_debug_leafclk(long long *out);
/* tuning knobs for measurements (also settable through the environment as HDSDPCU_GEMM_VARIANT / HDSDPCU_CHOL_BLOCK /
 * HDSDPCU_CHOL_LEAF before hdsdpcu_init):
 *   "gemm_variant" CTA tile / pipeline of the DMMA GEMM: 0 = 128x128x16, 4 stages; 1 = 128x128x32, 3 stages;
 *                  2 = 128x64x16, 4 stages, 2 CTAs/SM; 3 = 128x64x32, 2 stages, 2 CTAs/SM, 4 warps of 64x32;
 *                  4 = as 3 with 8 warps of 32x32 (16 warps/SM, default); 5 = 128x128x32 with 16 warps of 32x32
 *   "gemm_thin"    products with at most this many 128 x 64 tiles run on the thin-tile kernel (32 x 128 tiles; default 96, 0 = off)
 *   "chol_block"   NB of the blocked look-ahead Cholesky, used when the padded dimension is >= 4 NB
 *                  (-1 = chosen by size, the default; 0 = pure recursion)
 *   "chol_sched"   schedule of the blocked Cholesky: 1 = one step of look-ahead, whole panels on the side stream;
 *                  2 / 3 = strip chain on the side stream + two steps of look-ahead on the main stream (3: the look-ahead
 *                  columns as one GEMM); -1 (default) = 3 up to n = 10240, 1 beyond
 *   "dist_delay"   multi-GPU factorisation: 1 = panels are applied in pairs (K = 2 nb) to the block columns that are not next
 *                  in line, 0 (default; faster on 8 B200) = every panel at once (also HDSDPCU_DIST_DELAY)
 *   "chol_leaf"    128x128 leaf kernel: 1 = column sweep, 2 = DMMA panels (default)
 *   "ldl_pivot"    LDL^T fallback: 1 (default) = bounded Bunch-Kaufman pivoting (1 x 1 and 2 x 2 pivots) inside every 128 x 128
 *                  leaf; 0 = unpivoted L J L^T with static pivoting only (round-1 behaviour)
 *   "chol_partition" strip schedule: 1 = the chain runs in an 8-SM green-context partition, the bulk GEMMs in the other 140 SMs;
 *                  0 (default) = both share all SMs (stream priorities only).  Measured slower on B200
 *   "invert_fork"  S^-1: sub-blocks of the recursion for L^-T up to this size run on a pool of streams (default 2048, 0 = one stream)
 *   "chol_tail"    1 (default): late in a large factorisation the trailing matrix is handed over to the block size / schedule of
 *                  its own size (2048 -> 512 -> 256 -> strip chain); 0 = one block size throughout (also HDSDPCU_CHOL_TAIL)
 *   "trsv_version" triangular solves: 1 = streaming, 2 = register-prefetched tiles, one CTA per row block (default),
 *                  3 = one ticket per 128 x 128 tile with ordered accumulation (measured slower: DESIGN.md section 7)
 *   "chol_graph"   1 (default): factorisations up to n = 6144 are replayed from a captured CUDA graph from their third call on */
int hdsdpcu_set_option(const char *name, int value);

/* ---------------------------------------------------------------------------------------------
 * B1 -- dense linear-system back-end.  One-for-one replacement of the 11 function pointers of
 * hdsdp_linsys_fp (reference linalg/def_hdsdp_linsolver.h:40-63) as implemented for LAPACK by
 * lapackLinSolver* (reference linalg/hdsdp_linsolver.c:1044-1286).  Same argument meaning:
 *   numeric / psdcheck : elem = nCol x nCol column-major, lower triangle read, NOT modified
 *   fsolve / bsolve    : L \ rhs, L' \ rhs ; sol == NULL means in place        (:1146, :1172)
 *   solve              : (L L') \ rhs                                          (:1198)
 *   getdiag            : diag(L)                                               (:1227)
 *   invert             : full symmetric inverse into fullInv, aux unused       (:1238)
 * "not positive definite" is not an error: psdcheck sets *isPsd = 0 and returns 0; numeric returns 1
 * exactly where dpotrf's info != 0 makes the reference return FAILED (:1099-1103).
 * ------------------------------------------------------------------------------------------- */
int  hdsdpcu_linsys_create(void **pchol, int nCol);                                   /* cholCreate  */
void hdsdpcu_linsys_setparam(void *chol, void *param);                                /* cholSetParam (no-op) */
int  hdsdpcu_linsys_symbolic(void *chol, int *colBeg, int *colIdx);                   /* cholSymbolic (no-op) */
int  hdsdpcu_linsys_numeric(void *chol, int *colBeg, int *colIdx, double *elem);      /* cholNumeric */
int  hdsdpcu_linsys_psdcheck(void *chol, int *colBeg, int *colIdx, double *elem, int *isPsd); /* cholPsdCheck */
void hdsdpcu_linsys_fsolve(void *chol, int nRhs, double *rhs, double *sol);           /* cholFSolve  */
void hdsdpcu_linsys_bsolve(void *chol, int nRhs, double *rhs, double *sol);           /* cholBSolve  */
int  hdsdpcu_linsys_solve(void *chol, int nRhs, double *rhs, double *sol);            /* cholSolve   */
int  hdsdpcu_linsys_getdiag(void *chol, double *diag);                                /* cholGetDiag */
void hdsdpcu_linsys_invert(void *chol, double *fullInv, double *aux);                 /* cholInvert  */
void hdsdpcu_linsys_destroy(void **pchol);                                            /* cholDestroy */
/* Indefinite back-end (reference lapackIndefiniteLinSolver*, dsytrf/dsytrs, linalg/hdsdp_linsolver.c:1662-1825, selected by
 * HFpLinsysSwitchToIndefinite :1827 when dpotrf of the Schur matrix fails): with set_indefinite(1) numeric computes the
 * factorisation A = L J L^T, J = diag(+-1), whose 128 x 128 diagonal leaves are factored with Bunch-Kaufman pivoting
 * (1 x 1 and 2 x 2 pivots, as dsytrf; the pivot search is bounded to the leaf, so the block recursion and the multi-GPU
 * layout are those of the Cholesky path) and static pivoting as the backstop (an eigenvalue of a pivot block with
 * |lambda| <= 1e-13 max|A_ii| is replaced and counted); solve applies L^-1, J, L^-T.  set_option("ldl_pivot", 0) gives the
 * unpivoted round-1 variant. */
int  hdsdpcu_linsys_set_indefinite(void *chol, int on);
int  hdsdpcu_linsys_inertia(void *chol, int *nNegative, int *nPerturbed);
/* device-resident variants (no host round trip): d_elem has leading dimension ld >= nCol */
int  hdsdpcu_linsys_padded_dim(void *chol);                   /* leading dimension of the internal buffers */
int  hdsdpcu_linsys_numeric_dev(void *chol, const double *d_elem, long ld, int *info);
int  hdsdpcu_linsys_solve_dev(void *chol, int nRhs, double *d_x, long ldx);           /* in place, ldx >= padded dim, padding zero */
int  hdsdpcu_linsys_invert_dev(void *chol, double *d_inv);    /* padded_dim x padded_dim, full symmetric */
double *hdsdpcu_linsys_factor_dev(void *chol);                /* device pointer of L (ld = padded dim) */

/* ---------------------------------------------------------------------------------------------
 * B2 -- SDP cone.  Replaces the hot-path members of the cone vtable (reference
 * interface/def_hdsdp_conic.h:56-107) for hdsdp_cone_sdp_dense / hdsdp_cone_sdp_sparse:
 *   create          = coneProcData + conePresolveData   hdsdp_conic_sdp.c:1356-1400, :1490-1518
 *                     (user_data CSC [n(n+1)/2 x (m+1)], column 0 = objective, def_hdsdp_user_data.h:22-32)
 *   setstart        = coneSetStart      :1546        reduceresi = coneReduceResi :2226
 *   setperturb      = coneSetPerturb    :2238        scal       = coneScal       :1604
 *   update          = coneUpdate        :1616   S = -Rd I - A'y + tau C  into BUFFER_DUALVAR
 *   updatebuffer    = sdpDenseConeIUpdateBuffer :343 (generic linear combination, any buffer)
 *   interiorcheck   = coneInteriorCheck :2172   (update + Cholesky, *isInterior = PSD)
 *   interiorcheckexpert = coneInteriorCheckExpert :2192
 *   getbarrier      = coneGetBarrier    :2252   (optional update + factor; logdet = 2 sum log L_ii)
 *   addstepandcheck = coneAxpyBufferAndCheck :2333
 *   buildschur      = coneBuildSchur    :1726 / :1814  (accumulates into the hdsdpcu kkt object)
 * whichBuffer: 0 BUFFER_DUALVAR, 1 BUFFER_DUALCHECK, 2 BUFFER_DUALSTEP (interface/hdsdp_conic.h:24-26)
 * typeKKT: 0 INFEASIBLE, 1 CORRECTOR, 2 HOMOGENEOUS, 3 PRIMAL (interface/hdsdp_conic.h:16-19)
 * ------------------------------------------------------------------------------------------- */
int  hdsdpcu_cone_create(void **pcone, int nRow, int nCol, const int *coneMatBeg, const int *coneMatIdx,
                         const double *coneMatElem);
void hdsdpcu_cone_destroy(void **pcone);
int  hdsdpcu_cone_getdim(void *cone);
int  hdsdpcu_cone_gettypes(void *cone, int *types /* nRow + 1, last = objective; sdp_coeff_type values */);
void hdsdpcu_cone_setstart(void *cone, double rResi);
void hdsdpcu_cone_reduceresi(void *cone, double resiReduction);
void hdsdpcu_cone_setperturb(void *cone, double dDualPerturb);
int  hdsdpcu_cone_scal(void *cone, double dScal);
int  hdsdpcu_cone_update(void *cone, double barHsdTau, const double *rowDual);
int  hdsdpcu_cone_update_dev(void *cone, double barHsdTau, const double *d_rowDual);
int  hdsdpcu_cone_updatebuffer(void *cone, double dCCoef, double dACoefScal, const double *dACoef, double dEyeCoef,
                               int whichBuffer);
int  hdsdpcu_cone_interiorcheck(void *cone, double barHsdTau, const double *rowDual, int *isInterior);
int  hdsdpcu_cone_interiorcheckexpert(void *cone, double dCCoef, double dACoefScal, const double *dACoef,
                                      double dEyeCoef, int whichBuffer, int *isInterior);
int  hdsdpcu_cone_factorize(void *cone, int whichBuffer, int *isPsd);   /* HFpLinsysPsdCheck on the buffer */
int  hdsdpcu_cone_getbarrier(void *cone, double barHsdTau, const double *rowDual /* may be NULL */, int whichBuffer,
                             double *logdet);
int  hdsdpcu_cone_addstepandcheck(void *cone, double dStep, int whichBuffer, int *isInterior);
int  hdsdpcu_cone_buildschur(void *cone, int iCone, void *kkt, int typeKKT);
/* Dual ratio test on the device.  ratiotest = coneRatioTest (sdpDenseConeRatioTestImpl, hdsdp_conic_sdp.c:1642):
 * dS = dAdaRatio Rd I - A' dy + dTau C into BUFFER_DUALSTEP, then the reference's Lanczos (linalg/hdsdp_lanczos.c:161: same
 * srand(n) start vector, check frequency, residual tests and step formula) on w -> -L^-1 dS L^-T w with every n-vector in HBM;
 * *maxStep = largest alpha with S + alpha dS >= 0 (INFINITY if none).  whichBuffer selects the factor (DUALVAR / DUALCHECK),
 * which must be current (hdsdpcu_cone_factorize / _getbarrier).  lanczosmultiply = one operator application with host
 * vectors (sdpDenseConeILanczosMultiply :462) for a host-side HLanczosSolve; lanczossteps = steps of the last ratio test. */
int  hdsdpcu_cone_ratiotest(void *cone, double barHsdTauStep, const double *rowDualStep, double dAdaRatio, int whichBuffer,
                            double *maxStep);
int  hdsdpcu_cone_lanczosmultiply(void *cone, int whichBuffer, const double *x, double *y);
int  hdsdpcu_cone_lanczossteps(void *cone);
/* Primal recovery.  conePRecover = sdpDenseConeGetPrimal (hdsdp_conic_sdp.c:2395-2446): with S = C - A'y
 * (checked for positive definiteness in BUFFER_DUALCHECK) and dS = A'dy, dConePrimal (n x n, column-major, full symmetric)
 * = mu (S^-1 + S^-1 dS S^-1).  *isFeasible = 0 reproduces the reference's "Recovery step is infeasible" (nothing written). */
/* Extreme eigenvalue of a symmetric n x n host matrix by Lanczos on the device (largest != 0: lambda_max, else lambda_min),
 * to 1e-12 ||X||.  Replaces the dsyevr call of the DIMACS check (HDSDPCheckSolution, interface/hdsdp.c:852-861:
 * fds_syev(n, X, d, Y, 1, ...) returns the LARGEST eigenvalue of the primal block). */
int  hdsdpcu_sym_extreme_eig(int n, const double *X, int largest, double *eig, int *lanczosSteps);
/* coneBuildPrimalDirection = sdpDenseConeBuildPrimalXSXDirection (hdsdp_conic_sdp.c:2021, fds_trimultiply dense_opts.c:102), used by
 * the PSDP primal refinement: dPrimalXSXBuffer += X S X, S = BUFFER_DUALVAR (iDualMat != 0) or BUFFER_DUALSTEP; host n x n matrices. */
int  hdsdpcu_cone_buildprimalxsx(void *cone, const double *dPrimalScalMatrix, double *dPrimalXSXBuffer, int iDualMat);
int  hdsdpcu_cone_getprimal(void *cone, double dBarrierMu, const double *dRowDual, const double *dRowDualStep,
                            double *dConePrimal, int *isFeasible);
/* coneXDotS = sdpDenseConeXDotS (hdsdp_conic_sdp.c:2549, fds_dot_fds dense_opts.c:134): <S, X> over the lower triangles, X host n x n.
 * coneDRecover = sdpDenseConeGetDual (:2497): the dual matrix S (BUFFER_DUALVAR), full symmetric, to a host n x n matrix. */
int  hdsdpcu_cone_xdots(void *cone, const double *dConePrimal, double *xDotS);
int  hdsdpcu_cone_getdual(void *cone, double *dConeDual);
/* Hand the cone an S^-1 computed elsewhere (used by the integration shim, whose S factor is owned by the
 * reference's hdsdp_linsys_fp): from a host n x n matrix, or device-to-device from a hdsdpcu_linsys handle
 * (HFpLinsysInvert, linalg/hdsdp_linsolver.c:2120, without the host round trip).  Valid until the next update. */
int  hdsdpcu_cone_setsinv(void *cone, const double *fullInv);
int  hdsdpcu_cone_setsinv_linsys(void *cone, void *chol);
/* test / debug mirrors (D2H): n x n column-major */
int  hdsdpcu_cone_getbuffer(void *cone, int whichBuffer, double *out);
int  hdsdpcu_cone_getsinv(void *cone, double *out);
int  hdsdpcu_cone_getfactordiag(void *cone, int whichBuffer, double *diag);

/* ---------------------------------------------------------------------------------------------
 * B2 -- Schur complement object.  Replaces hdsdp_kkt (reference interface/def_hdsdp_schur.h:32-68)
 * for the dense-M case and the HKKT* calls of interface/hdsdp_schur.c:
 *   create (+addcone) = HKKTCreate + HKKTInit :170-254     buildup   = HKKTBuildUp :256
 *   buildupextra_*    = HKKTBuildUpExtraCone :270 for the bound cone (hdsdp_conic_bound.c:201-249)
 *                       and the LP cone (hdsdp_conic_lp.c:254-330)
 *   regularize = HKKTRegularize :348    export = HKKTExport :293    factorize = HKKTFactorize :328
 *   solve      = HKKTSolve :338 (direct Cholesky instead of PCG)    registerpsdp = HKKTRegisterPSDP :375
 * M (m x m, lower) lives in HBM from clean to the last solve; getmatrix is a test-only mirror.
 * ------------------------------------------------------------------------------------------- */
int  hdsdpcu_kkt_create(void **pkkt, int nRow);
int  hdsdpcu_kkt_addcone(void *kkt, void *cone);
void hdsdpcu_kkt_destroy(void **pkkt);
int  hdsdpcu_kkt_buildup(void *kkt, int typeKKT);
int  hdsdpcu_kkt_clean(void *kkt, int typeKKT);
int  hdsdpcu_kkt_buildupextra_bound(void *kkt, const double *diagAdd, const double *asinvAdd, const double *asinvRdAdd,
                                    int typeKKT);
/* generic host-computed increments (any pointer may be NULL): diag(M), dASinvVec, dASinvRdSinvVec, dASinvCSinvVec,
 * scalars {dCSinvCSinv, dCSinv, dCSinvRdSinv, dTraceSinv} */
int  hdsdpcu_kkt_addhost(void *kkt, const double *diagAdd, const double *asinvAdd, const double *asinvRdAdd,
                         const double *asinvCAdd, const double *scalarsAdd4);
int  hdsdpcu_lp_create(void **plp, int nRow, int nLpCol, const int *matBeg, const int *matIdx, const double *matElem);
int  hdsdpcu_lp_setobjective(void *lp, const double *colObj);   /* current (possibly rescaled) LP objective, used by the HOMOGENEOUS terms */
void hdsdpcu_lp_destroy(void **plp);
int  hdsdpcu_kkt_buildupextra_lp(void *kkt, void *lp, const double *colDualInverse, double dualResidual, int typeKKT);
int  hdsdpcu_kkt_regularize(void *kkt, double dKKTReg);
int  hdsdpcu_kkt_export(void *kkt, double *dASinvVec, double *dASinvRdSinvVec, double *dASinvCSinvVec,
                        double *dCSinvCSinv, double *dCSinv, double *dCSinvRdSinv, double *dTraceSinv);
int  hdsdpcu_kkt_factorize(void *kkt);   /* Cholesky; on a non-positive pivot switches (for good) to the LDL^T back-end */
int  hdsdpcu_kkt_ldl_status(void *kkt, int *isLdl, int *nNegative, int *nPerturbed);
/* solve: like HFpLinsysSolve (linalg/hdsdp_linsolver.c:2088-2103) a failed solve or a NaN in the first entry of the solution /
 * right-hand side switches M's back-end to LDL^T for good, refactors and solves again.  In LDL^T mode every solve is followed
 * by iterative refinement on b - M x (M stays intact in HBM) and FAILS if the residual stays above 1e-9 max|b|;
 * solve_status reports the last relative residual and the refinement steps taken.  symv: y = M x (the dsymv of the reference's
 * PCG loop, :1446-1588) on the device-resident lower triangle. */
int  hdsdpcu_kkt_solve(void *kkt, const double *dRhsVec, double *dLhsVec /* NULL: in place */);
int  hdsdpcu_kkt_solve_status(void *kkt, double *relResidual, int *refineSteps);
/* Solver policy for M.  0 (default): direct Cholesky, as the north star asks.  1: the reference's own default policy
 * (HDSDP_LINSYS_DENSE_ITERATIVE, conjGradSolve linalg/hdsdp_linsolver.c:1446-1588 with the limits of interface/hdsdp_schur.c:21-35):
 * factorize only records diag(M); solve runs Jacobi-preconditioned CG with every vector in HBM (one symv over the lower triangle
 * per step); the first solve that hits the iteration limit or stalls (iter > 20 and ||r|| > 0.01 ||b||) factors M by Cholesky and
 * every later solve is direct (sticky, as the reference's useJacobi = 0).  Single GPU only.  pcg_status: still on Jacobi?, CG
 * steps of the last solve, number of CG solves, number of fallbacks. */
int  hdsdpcu_kkt_set_solver(void *kkt, int mode);
int  hdsdpcu_kkt_pcg_status(void *kkt, int *useJacobi, int *lastIterations, int *nSolves, int *nFallbacks);
int  hdsdpcu_kkt_symv(void *kkt, const double *x, double *y);
int  hdsdpcu_kkt_solve_many(void *kkt, int nRhs, const double *dRhsVec, double *dLhsVec);
void hdsdpcu_kkt_registerpsdp(void *kkt, int nCones, double **dPrimalX);
int  hdsdpcu_kkt_getmatrix(void *kkt, double *M /* nRow x nRow column-major, lower meaningful */);
/* device-resident variants used by bench.py's HBM-resident timing */
int     hdsdpcu_kkt_padded_dim(void *kkt);
double *hdsdpcu_kkt_matrix_dev(void *kkt);
double *hdsdpcu_kkt_asinv_dev(void *kkt);
int     hdsdpcu_kkt_solve_dev(void *kkt, int nRhs, double *d_x /* stride padded dim, padding zero, in place */);
/* multi-GPU: this process builds only Schur columns j with (j / 128) % nRanks == rank */
int  hdsdpcu_kkt_setshard(void *kkt, int rank, int nRanks);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU Schur matrix (replaces the single dpotrf of HKKTFactorize, interface/hdsdp_schur.c:328,
 * linalg/hdsdp_linsolver.c:1096, when one process per GPU runs the solver).  M is distributed 1-D block-cyclic by
 * block columns of width blockSize: rank r assembles and factors the columns c with (c / blockSize) % nRanks == r;
 * finished factor panels are copied into the peers' buffers over NVLink peer memory (CUDA IPC), so after
 * hdsdpcu_kkt_factorize every rank holds the complete factor and hdsdpcu_kkt_solve needs no communication.
 *   1. every rank: hdsdpcu_kkt_dist_init, hdsdpcu_kkt_dist_export -> blob of hdsdpcu_dist_blob_bytes() bytes
 *   2. the host program all-gathers the blobs in rank order (MPI_Allgather, torch.distributed, a file ...)
 *   3. every rank: hdsdpcu_kkt_dist_connect(all blobs)
 * buildup / regularize / factorize / solve are then called in lock step by all ranks with the same arguments;
 * the side vectors (hdsdpcu_kkt_export) are complete on every rank, hdsdpcu_kkt_getmatrix returns the owned columns.
 * ------------------------------------------------------------------------------------------- */
int  hdsdpcu_dist_blob_bytes(void);
int  hdsdpcu_dist_owner(int col, int blockSize, int nRanks);   /* pure host arithmetic, no device needed */
int  hdsdpcu_kkt_dist_init(void *kkt, int rank, int nRanks, int blockSize);
int  hdsdpcu_kkt_dist_export(void *kkt, void *blob);
int  hdsdpcu_kkt_dist_connect(void *kkt, const void *blobs /* nRanks blobs, rank order */);
/* test hook: the same schedule with nRanks ranks inside this process on one GPU (events instead of peer flags);
 * indefinite != 0: the LDL^T mode (A = L J L^T), outSign0 = diag(J) as seen by rank 0 */
int  hdsdpcu_distchol_selftest(int n, int blockSize, int nRanks, const double *A, double *outL0, double *outLlast, int *info,
                               int indefinite, double *outSign0);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone kernels exposed for tests and roofline measurement
 * ------------------------------------------------------------------------------------------- */
/* C (M x N) = alpha A (M x K) B(N x K)^T + beta C on device pointers; M,N % 128 == 0, K % 16 == 0 */
int hdsdpcu_dgemm_nt_dev(int M, int N, int K, double alpha, const double *dA, long lda, const double *dB, long ldb,
                         double beta, double *dC, long ldc, int lowerOnly);

#ifdef __cplusplus
}
#endif
#endif /* HDSDPCU_H */
