/*
 * oracle/hdsdp_oracle.c -- plain-C CPU restatement of HDSDP's Newton-system hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for tests/, __graft_entry__.smoke() and the
 * "port" leg of bench.py's cpu_baseline.  Nothing in the product package (hdsdp_b200/) may import,
 * link or execute it; the product path has no CPU fallback.
 *
 * Parity is PINNED: tests/test_cpu_oracle.py checks every function here against the golden fixtures
 * in tests/golden/ (npz files), which tests/golden/make_golden.py produced by running the unmodified
 * reference (oracle/_ref, built by oracle/build_ref.sh) on its own examples (mcp100, theta1, truss1,
 * gpp100) and on small synthetic problems.
 *
 * It restates, without BLAS (naive loops), the algorithm of the reference (all paths /root/reference):
 *   classification        linalg/hdsdp_sdpdata.c:2321-2345, :2373-2449, :880-899;
 *                         linalg/sparse_opts.c:428-516; linalg/dense_opts.c:233-285
 *   cone type / rows      interface/hdsdp_user_data.c:73-98; interface/hdsdp_conic_sdp.c:1356-1486
 *   strategy + ordering   interface/hdsdp_conic_sdp.c:539-662 (cost model, descending nnz sort)
 *   S assembly            interface/hdsdp_conic_sdp.c:343-402; linalg/hdsdp_sdpdata.c:589-683
 *   Cholesky / inverse    linalg/hdsdp_linsolver.c:1082-1110, :1238-1260 (dpotrf / dpotri semantics)
 *   Schur columns M2..M5  interface/hdsdp_conic_sdp.c:687-985 (dense cone), :1058-1260 (sparse cone)
 *   per-type kernels      linalg/hdsdp_sdpdata.c:985-2165
 *   HSD / corrector       interface/hdsdp_conic_sdp.c:987-1056
 *   driver                interface/hdsdp_conic_sdp.c:1726-1886
 *   LP / bound cones      interface/hdsdp_conic_lp.c:254-330; interface/hdsdp_conic_bound.c:201-249
 *   KKT object            interface/hdsdp_schur.c:141-165 (clean), :348-373 (regularize)
 * Every function below names the lines it follows.  It is written from the algorithm, not copied.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { T_ZERO = 0, T_SPARSE = 1, T_DENSE = 2, T_SPR1 = 3, T_DSR1 = 4 };
enum { K_INFEASIBLE = 0, K_CORRECTOR = 1, K_HOMOGENEOUS = 2, K_PRIMAL = 3 };
enum { M1 = 0, M2 = 1, M3 = 2, M4 = 3, M5 = 4 };

typedef struct {
    int type, n;
    int nnz; int *row, *col; double *val;   /* SPARSE lower triplets */
    double *packed;                          /* DENSE packed lower, column-major */
    double sign; int nfac; int *idx; double *fv; /* SPR1: idx/fv ; */
    double *fac;                             /* SPR1 + DSR1: dense n-vector copy of the factor */
} ocoef;

typedef struct {
    int m, n;
    int dense_cone;       /* 1: dense-type SDP cone (permuted strategies), 0: sparse-type cone */
    ocoef *rows;          /* m coefficients (ZERO where absent) */
    ocoef obj;
    int nrowelem; int *rowidx; /* sparse cone: non-empty rows */
    int *perm, *strat;    /* dense cone */
    double rd, perturb;
    double *S, *L, *Sinv; /* n x n column-major */
    double *B1, *B2;      /* n x n work buffers (kktBuffer, kktBuffer2) */
    int factored;
} ocone;

#define FULL(A, n, i, j) ((A)[(size_t) (j) * (n) + (i)])
static long pack_start(long n, long j) { return j * n - j * (j - 1) / 2; }

/* ---------------------------------------------------------------------------------------------
 * classification
 * ------------------------------------------------------------------------------------------- */
/* sparse_opts.c:428-443 (robust to unsorted input: each slot decoded independently) */
static void unpack(int n, long p, int *r, int *c) {
    long lo = 0, hi = n - 1;
    while (lo < hi) { long mid = (lo + hi + 1) / 2; if (pack_start(n, mid) <= p) lo = mid; else hi = mid - 1; }
    *c = (int) lo; *r = (int) (p - pack_start(n, lo) + lo);
}

/* sparse_opts.c:453-516 */
static int sparse_r1(int n, int nnz, const int *Ai, const int *Aj, const double *Ax, double *sgn, double *a) {
    int i = Ai[0], j = Aj[0], k, anz = 0; double v = Ax[0], s, eps = 0.0;
    memset(a, 0, sizeof(double) * n);
    if (i != j) return 0;
    if (nnz == 1) { *sgn = Ax[0]; a[i] = 1.0; return 1; }
    s = v > 0 ? 1.0 : -1.0; v = sqrt(fabs(v));
    for (k = 0; k < nnz; ++k) { if (Aj[k] > i) break; a[Ai[k]] = Ax[k] / v; anz++; }
    if (nnz != anz * (anz + 1) / 2) return 0;
    if (k == n) return 0;
    for (k = 0; k < nnz; ++k) eps += fabs(Ax[k] - s * a[Ai[k]] * a[Aj[k]]);
    if (eps > 1e-10) return 0;
    *sgn = s; return 1;
}

/* dense_opts.c:233-285 */
static int dense_r1(int n, const double *A, double *sgn, double *a) {
    int i, c, jj; long k = 0, id = 0; double s, v, eps = 0.0;
    for (i = 0; i < n; ++i) { if (A[k] != 0) break; k += n - i; }
    if (i == n) return 0;
    s = A[k] > 0 ? 1.0 : -1.0; v = sqrt(fabs(A[k]));
    for (c = 0; c < n; ++c) a[c] = A[(long) ((2L * n - i - 1) * i / 2) + c] / v;
    for (c = 0; c < n; ++c) {
        for (jj = 0; jj < n - c; ++jj) eps += fabs(A[id + jj] - s * a[c] * a[c + jj]);
        id += n - c;
        if (eps > 1e-10) return 0;
    }
    *sgn = s; return 1;
}

/* hdsdp_sdpdata.c:2321-2345 (set data) + :2373-2449 (rank-one promotion) + :880-899 (normalise) */
static void classify(ocoef *o, int n, int nnz, const int *Ci, const double *Cx) {
    long npack = (long) n * (n + 1) / 2; int e, r1 = 0, nz = 0, r; double sgn = 0.0, nrm = 0.0;
    double *a = (double *) calloc(n, sizeof(double));
    memset(o, 0, sizeof(*o)); o->n = n;
    if (nnz == 0) { o->type = T_ZERO; free(a); return; }
    if ((double) nnz > 0.3 * (double) npack) {
        o->type = T_DENSE; o->packed = (double *) calloc(npack, sizeof(double));
        for (e = 0; e < nnz; ++e) o->packed[Ci[e]] = Cx[e];
        r1 = dense_r1(n, o->packed, &sgn, a);
    } else {
        o->type = T_SPARSE; o->nnz = nnz;
        o->row = (int *) malloc(sizeof(int) * nnz); o->col = (int *) malloc(sizeof(int) * nnz); o->val = (double *) malloc(sizeof(double) * nnz);
        for (e = 0; e < nnz; ++e) { unpack(n, Ci[e], &o->row[e], &o->col[e]); o->val[e] = Cx[e]; }
        r1 = sparse_r1(n, nnz, o->row, o->col, o->val, &sgn, a);
    }
    if (!r1) { free(a); return; }
    for (r = 0; r < n; ++r) if (fabs(a[r]) > 1e-10) nz++;
    free(o->row); free(o->col); free(o->val); free(o->packed); o->row = o->col = NULL; o->val = o->packed = NULL; o->nnz = 0;
    o->sign = sgn; o->fac = (double *) calloc(n, sizeof(double));
    if ((double) nz > 0.5 * (double) n) {
        o->type = T_DSR1;
        for (r = 0; r < n; ++r) { o->fac[r] = a[r]; nrm += a[r] * a[r]; }
        nrm = sqrt(nrm); o->sign *= nrm * nrm;
        for (r = 0; r < n; ++r) o->fac[r] /= nrm;
    } else {
        o->type = T_SPR1; o->nfac = nz; o->idx = (int *) malloc(sizeof(int) * nz); o->fv = (double *) malloc(sizeof(double) * nz);
        for (r = 0, e = 0; r < n; ++r) if (fabs(a[r]) > 1e-10) { o->idx[e] = r; o->fv[e] = a[r]; o->fac[r] = a[r]; nrm += a[r] * a[r]; e++; }
        nrm = sqrt(nrm); o->sign *= nrm * nrm;
        for (e = 0; e < nz; ++e) o->fv[e] /= nrm;
        for (r = 0; r < n; ++r) o->fac[r] /= nrm;
    }
    free(a);
}

static void coef_free(ocoef *o) { free(o->row); free(o->col); free(o->val); free(o->packed); free(o->idx); free(o->fv); free(o->fac); }
static int coef_rank(const ocoef *o) { return o->type == T_ZERO ? 0 : (o->type == T_SPR1 || o->type == T_DSR1) ? 1 : o->n; } /* hdsdp_sdpdata.c:2347-2358 */
static int coef_nnz(const ocoef *o) {  /* hdsdp_sdpdata.c:397-435 */
    switch (o->type) { case T_SPARSE: return o->nnz; case T_DENSE: case T_DSR1: return o->n * (o->n + 1) / 2;
                       case T_SPR1: return o->nfac * (o->nfac + 1) / 2; default: return 0; }
}

/* hdsdp_conic_sdp.c:539-600 */
static int choose_strategy(const int *ranks, const int *sps, const int *perm, int m, int n, int ip) {
    int best = M1, q; double bests = 1e30, n3 = (double) n * n * n, after = 0.0, k = 1.5, s2, s3, s4, s5;
    int rank = ranks[perm[ip]];
    for (q = ip; q < m; ++q) after += sps[q];
    s2 = rank * ((double) sps[ip] * n + 3 * k * after);
    s3 = (double) n * k * sps[ip] + n3 + k * after + n3 / m;
    s4 = (double) n * k * sps[ip] + k * (n + 1) * after + n3 / m;
    s5 = k * (2.0 * k * sps[ip] + 1) * after + n3 / m;
    if (s2 <= bests) { best = M2; bests = s2; }
    if (s3 < bests) { best = M3; bests = s3; }
    if (s4 < bests) { best = M4; bests = s4; }
    if (s5 < bests) { best = M5; bests = s5; }
    return best;
}

/* Descending sort of `perm` by key; the reference uses its own quicksort (hdsdp_utils.c HUtilDescendSortIntByInt)
 * which also permutes the key array; ties may land in a different order than here, which changes which of two
 * equal-nnz rows is visited first but not any M entry.  Keys are sorted together with perm as the reference does. */
static void sort_desc(int *perm, int *key, int lo, int hi) {
    int i = lo, j = hi, mid = (lo + hi) / 2, p = key[mid], t;
    if (lo >= hi) return;
    while (i <= j) {
        while (key[i] > p) i++;
        while (key[j] < p) j--;
        if (i <= j) { t = key[i]; key[i] = key[j]; key[j] = t; t = perm[i]; perm[i] = perm[j]; perm[j] = t; i++; j--; }
    }
    sort_desc(perm, key, lo, j); sort_desc(perm, key, i, hi);
}

void *orc_cone_create(int m, int n, const int *beg, const int *idx, const double *elem) {
    ocone *c = (ocone *) calloc(1, sizeof(ocone)); int i, nz = 0;
    c->m = m; c->n = n;
    c->rows = (ocoef *) calloc(m, sizeof(ocoef));
    classify(&c->obj, n, beg[1] - beg[0], idx + beg[0], elem + beg[0]);
    for (i = 0; i < m; ++i) { classify(&c->rows[i], n, beg[i + 2] - beg[i + 1], idx + beg[i + 1], elem + beg[i + 1]); if (beg[i + 2] > beg[i + 1]) nz++; }
    c->dense_cone = ((double) nz > 0.3 * (double) m); /* hdsdp_user_data.c:81-85 */
    c->rowidx = (int *) malloc(sizeof(int) * (m > 0 ? m : 1));
    for (i = 0; i < m; ++i) if (c->rows[i].type != T_ZERO) c->rowidx[c->nrowelem++] = i;
    c->perm = (int *) malloc(sizeof(int) * (m > 0 ? m : 1)); c->strat = (int *) malloc(sizeof(int) * (m > 0 ? m : 1));
    {   /* hdsdp_conic_sdp.c:602-662 */
        int *ranks = (int *) malloc(sizeof(int) * (m > 0 ? m : 1)), *sps = (int *) malloc(sizeof(int) * (m > 0 ? m : 1));
        for (i = 0; i < m; ++i) { c->perm[i] = i; ranks[i] = coef_rank(&c->rows[i]); sps[i] = coef_nnz(&c->rows[i]); }
        if (m > 0) sort_desc(c->perm, sps, 0, m - 1);
        for (i = 0; i < m; ++i) c->strat[i] = choose_strategy(ranks, sps, c->perm, m, n, i);
        free(ranks); free(sps);
    }
    c->S = (double *) calloc((size_t) n * n, sizeof(double)); c->L = (double *) calloc((size_t) n * n, sizeof(double));
    c->Sinv = (double *) calloc((size_t) n * n, sizeof(double)); c->B1 = (double *) calloc((size_t) n * n, sizeof(double));
    c->B2 = (double *) calloc((size_t) n * n + 2 * n, sizeof(double));
    return c;
}

void orc_cone_destroy(void *h) {
    ocone *c = (ocone *) h; int i;
    if (!c) return;
    for (i = 0; i < c->m; ++i) coef_free(&c->rows[i]);
    coef_free(&c->obj); free(c->rows); free(c->rowidx); free(c->perm); free(c->strat);
    free(c->S); free(c->L); free(c->Sinv); free(c->B1); free(c->B2); free(c);
}

int orc_cone_is_dense_type(void *h) { return ((ocone *) h)->dense_cone; }
void orc_cone_types(void *h, int *types) { ocone *c = (ocone *) h; int i; for (i = 0; i < c->m; ++i) types[i] = c->rows[i].type; types[c->m] = c->obj.type; }
void orc_cone_strategies(void *h, int *perm, int *strat) { ocone *c = (ocone *) h; memcpy(perm, c->perm, sizeof(int) * c->m); memcpy(strat, c->strat, sizeof(int) * c->m); }
double orc_cone_r1sign(void *h, int i) { ocone *c = (ocone *) h; ocoef *o = i < c->m ? &c->rows[i] : &c->obj; return (o->type == T_SPR1 || o->type == T_DSR1) ? o->sign : 0.0; }
void orc_cone_set_resi(void *h, double rd) { ((ocone *) h)->rd = rd; }
void orc_cone_set_perturb(void *h, double p) { ((ocone *) h)->perturb = p; }
void orc_cone_scal_obj(void *h, double s) {  /* hdsdp_conic_sdp.c:1604-1614 -> sdpDataMatScal */
    ocone *c = (ocone *) h; ocoef *o = &c->obj; long e, np = (long) c->n * (c->n + 1) / 2;
    if (o->type == T_SPARSE) for (e = 0; e < o->nnz; ++e) o->val[e] *= s;
    else if (o->type == T_DENSE) for (e = 0; e < np; ++e) o->packed[e] *= s;
    else if (o->type == T_SPR1 || o->type == T_DSR1) o->sign *= s;
}

/* ---------------------------------------------------------------------------------------------
 * S assembly: hdsdp_sdpdata.c:589-683 (add2buffer per type), hdsdp_conic_sdp.c:343-402
 * ------------------------------------------------------------------------------------------- */
static void add_to_buffer(const ocoef *o, double a, double *T) {
    int n = o->n, e, i, j; long p;
    if (a == 0.0) return; /* hdsdp_sdpdata.c:2479 */
    switch (o->type) {
        case T_SPARSE: for (e = 0; e < o->nnz; ++e) FULL(T, n, o->row[e], o->col[e]) += a * o->val[e]; break;
        case T_DENSE: for (j = 0, p = 0; j < n; ++j) for (i = j; i < n; ++i, ++p) FULL(T, n, i, j) += a * o->packed[p]; break;
        case T_SPR1: for (i = 0; i < o->nfac; ++i) for (j = 0; j <= i; ++j) FULL(T, n, o->idx[i], o->idx[j]) += a * o->sign * o->fv[i] * o->fv[j]; break;
        case T_DSR1: for (j = 0; j < n; ++j) for (i = j; i < n; ++i) FULL(T, n, i, j) += a * o->sign * o->fac[i] * o->fac[j]; break;
        default: break;
    }
}

/* T <- eye*I + aScal * sum a_i A_i + cC * C ; perturbation added unless is_step (hdsdp_conic_sdp.c:382-384) */
void orc_cone_update_buffer(void *h, double cC, double aScal, const double *a, double eye, int is_step, double *T) {
    ocone *c = (ocone *) h; int n = c->n, i;
    memset(T, 0, sizeof(double) * (size_t) n * n);
    for (i = 0; i < c->m; ++i) add_to_buffer(&c->rows[i], aScal * a[i], T);
    add_to_buffer(&c->obj, cC, T);
    if (!is_step) eye += c->perturb;
    if (eye != 0.0) for (i = 0; i < n; ++i) FULL(T, n, i, i) += eye;
}

/* ---------------------------------------------------------------------------------------------
 * dense Cholesky (lower, dpotrf semantics), inverse (dpotri + mirror), triangular solves
 * ------------------------------------------------------------------------------------------- */
int orc_potrf(int n, const double *A, double *L) {
    int i, j, k;
    for (j = 0; j < n; ++j) for (i = 0; i < n; ++i) FULL(L, n, i, j) = (i >= j) ? FULL(A, n, i, j) : 0.0;
    for (j = 0; j < n; ++j) {
        double d = FULL(L, n, j, j);
        for (k = 0; k < j; ++k) d -= FULL(L, n, j, k) * FULL(L, n, j, k);
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d); FULL(L, n, j, j) = d;
        for (i = j + 1; i < n; ++i) {
            double s = FULL(L, n, i, j);
            for (k = 0; k < j; ++k) s -= FULL(L, n, i, k) * FULL(L, n, j, k);
            FULL(L, n, i, j) = s / d;
        }
    }
    return 0;
}
void orc_fsolve(int n, const double *L, double *x) { int i, k; for (i = 0; i < n; ++i) { double s = x[i]; for (k = 0; k < i; ++k) s -= FULL(L, n, i, k) * x[k]; x[i] = s / FULL(L, n, i, i); } }
void orc_bsolve(int n, const double *L, double *x) { int i, k; for (i = n - 1; i >= 0; --i) { double s = x[i]; for (k = i + 1; k < n; ++k) s -= FULL(L, n, k, i) * x[k]; x[i] = s / FULL(L, n, i, i); } }
void orc_invert(int n, const double *L, double *inv) {
    int i, j; double *e = (double *) malloc(sizeof(double) * n);
    for (j = 0; j < n; ++j) {
        memset(e, 0, sizeof(double) * n); e[j] = 1.0;
        orc_fsolve(n, L, e); orc_bsolve(n, L, e);
        for (i = 0; i < n; ++i) FULL(inv, n, i, j) = e[i];
    }
    for (j = 0; j < n; ++j) for (i = j + 1; i < n; ++i) { double v = 0.5 * (FULL(inv, n, i, j) + FULL(inv, n, j, i)); FULL(inv, n, i, j) = v; FULL(inv, n, j, i) = v; }
    free(e);
}

/* S = -Rd I - A'y + tau C (hdsdp_conic_sdp.c:1630), factor; returns 1 if PSD; *logdet = 2 sum log L_ii (:2252-2290) */
int orc_cone_set_point(void *h, const double *y, double tau, double *logdet) {
    ocone *c = (ocone *) h; int n = c->n, i, info; double ld = 0.0;
    orc_cone_update_buffer(h, tau, -1.0, y, -c->rd, 0, c->S);
    info = orc_potrf(n, c->S, c->L);
    c->factored = (info == 0);
    if (info) return 0;
    for (i = 0; i < n; ++i) ld += log(FULL(c->L, n, i, i));
    if (logdet) *logdet = 2.0 * ld;
    return 1;
}
void orc_cone_get_S(void *h, double *out) { ocone *c = (ocone *) h; memcpy(out, c->S, sizeof(double) * (size_t) c->n * c->n); }
void orc_cone_get_L(void *h, double *out) { ocone *c = (ocone *) h; memcpy(out, c->L, sizeof(double) * (size_t) c->n * c->n); }
void orc_cone_get_Sinv(void *h, double *out) { ocone *c = (ocone *) h; memcpy(out, c->Sinv, sizeof(double) * (size_t) c->n * c->n); }

/* ---------------------------------------------------------------------------------------------
 * per-type kernels
 * ------------------------------------------------------------------------------------------- */
static void symv(int n, const double *A, const double *x, double *y) { int i, j; for (i = 0; i < n; ++i) { double s = 0; for (j = 0; j < n; ++j) s += FULL(A, n, i, j) * x[j]; y[i] = s; } }
static double ddot(int n, const double *x, const double *y) { int i; double s = 0; for (i = 0; i < n; ++i) s += x[i] * y[i]; return s; }
static double packed_entry(const ocoef *o, int i, int j) { int r = i >= j ? i : j, cc = i >= j ? j : i; return o->packed[pack_start(o->n, cc) + (r - cc)]; }

/* v = S^-1 a : hdsdp_sdpdata.c:1003-1044 */
static void k2_solve(const ocone *c, const ocoef *o, double *v) {
    int n = c->n, e, i;
    if (o->type == T_SPR1) {
        if (o->nfac >= 0.3 * n) { memcpy(v, o->fac, sizeof(double) * n); orc_fsolve(n, c->L, v); orc_bsolve(n, c->L, v); }
        else { memset(v, 0, sizeof(double) * n); for (e = 0; e < o->nfac; ++e) for (i = 0; i < n; ++i) v[i] += o->fv[e] * FULL(c->Sinv, n, i, o->idx[e]); }
    } else symv(n, c->Sinv, o->fac, v);
}
/* sign * a' v : hdsdp_sdpdata.c:1066-1088 */
static double k2_trace(const ocoef *o, const double *v) {
    int e; double s = 0;
    if (o->type == T_SPR1) { for (e = 0; e < o->nfac; ++e) s += v[o->idx[e]] * o->fv[e]; return o->sign * s; }
    return o->sign * ddot(o->n, o->fac, v);
}
/* v' A v : hdsdp_sdpdata.c:1090-1118, sparse_opts.c:565, dense_opts.c:287, r1_opts.c:43-72 */
static double k2_quad(const ocoef *o, const double *v) {
    int n = o->n, e, i, j; double s = 0, t;
    switch (o->type) {
        case T_SPARSE: for (e = 0; e < o->nnz; ++e) { t = o->val[e] * v[o->row[e]] * v[o->col[e]]; s += (o->row[e] == o->col[e]) ? 0.5 * t : t; } return 2.0 * s;
        case T_DENSE: for (i = 0; i < n; ++i) { t = 0; for (j = 0; j < n; ++j) t += packed_entry(o, i, j) * v[j]; s += t * v[i]; } return s;
        case T_SPR1: for (e = 0; e < o->nfac; ++e) s += o->fv[e] * v[o->idx[e]]; return o->sign * s * s;
        case T_DSR1: s = ddot(n, o->fac, v); return o->sign * s * s;
        default: return 0.0;
    }
}
/* B = S^-1 A S^-1 (lower), returns tr(A S^-1) : hdsdp_sdpdata.c:1127-1270 */
static double k3_sinvasinv(const ocone *c, const ocoef *o, double *aux, double *B) {
    int n = c->n, e, i, j, k; double tr = 0.0; const double *Si = c->Sinv;
    if (o->type == T_ZERO) return 0.0;
    if (o->type == T_SPARSE || o->type == T_DENSE) {
        memset(aux, 0, sizeof(double) * (size_t) n * n);
        if (o->type == T_SPARSE) { /* aux = S^-1 A by column combinations */
            for (e = 0; e < o->nnz; ++e) {
                int r = o->row[e], cc = o->col[e]; double a = o->val[e];
                for (i = 0; i < n; ++i) FULL(aux, n, i, cc) += a * FULL(Si, n, i, r);
                if (r != cc) for (i = 0; i < n; ++i) FULL(aux, n, i, r) += a * FULL(Si, n, i, cc);
            }
            for (i = 0; i < n; ++i) { tr += FULL(aux, n, i, i); for (j = 0; j <= i; ++j) { double s = 0; for (k = 0; k < n; ++k) s += FULL(aux, n, i, k) * FULL(Si, n, k, j); FULL(B, n, i, j) = s; } }
        } else { /* aux = A S^-1 */
            for (j = 0; j < n; ++j) for (i = 0; i < n; ++i) { double s = 0; for (k = 0; k < n; ++k) s += packed_entry(o, i, k) * FULL(Si, n, k, j); FULL(aux, n, i, j) = s; }
            for (j = 0; j < n; ++j) { tr += FULL(aux, n, j, j); for (i = 0; i <= j; ++i) { double s = 0; for (k = 0; k < n; ++k) s += FULL(aux, n, k, j) * FULL(Si, n, k, i); FULL(B, n, j, i) = s; } }
        }
        return tr;
    }
    memset(B, 0, sizeof(double) * (size_t) n * n);
    k2_solve(c, o, aux); tr = k2_trace(o, aux);
    for (j = 0; j < n; ++j) for (i = j; i < n; ++i) FULL(B, n, i, j) += o->sign * aux[i] * aux[j];
    return tr;
}
/* <A, B> with B full buffer whose LOWER triangle is meaningful : hdsdp_sdpdata.c:1280-1359 */
static double k3_dot(const ocoef *o, const double *B, double *aux) {
    int n = o->n, e, i, j; double s = 0; long p;
    switch (o->type) {
        case T_SPARSE: for (e = 0; e < o->nnz; ++e) { double t = o->val[e] * FULL(B, n, o->row[e], o->col[e]); s += (o->row[e] == o->col[e]) ? 0.5 * t : t; } return 2.0 * s;
        case T_DENSE: for (j = 0, p = 0; j < n; ++j) for (i = j; i < n; ++i, ++p) s += (i == j ? 0.5 : 1.0) * o->packed[p] * FULL(B, n, i, j); return 2.0 * s;
        case T_SPR1: for (j = 0; j < o->nfac; ++j) { s += 0.5 * o->fv[j] * o->fv[j] * FULL(B, n, o->idx[j], o->idx[j]); for (i = j + 1; i < o->nfac; ++i) s += o->fv[i] * o->fv[j] * FULL(B, n, o->idx[i], o->idx[j]); } return 2.0 * o->sign * s;
        case T_DSR1: /* fds_symv('L') reads the lower triangle only */
            for (i = 0; i < n; ++i) { double t = 0; for (j = 0; j < n; ++j) t += (i >= j ? FULL(B, n, i, j) : FULL(B, n, j, i)) * o->fac[j]; aux[i] = t; }
            return o->sign * ddot(n, o->fac, aux);
        default: return 0.0;
    }
}
/* B = A S^-1 (full, unsymmetric); returns tr(S^-1 A S^-1) when rd != 0 : hdsdp_sdpdata.c:1370-1557 */
static double k4_asinv(const ocone *c, const ocoef *o, double rd, double *aux, double *B) {
    int n = c->n, e, i, j, k; const double *Si = c->Sinv; double t = 0.0;
    memset(B, 0, sizeof(double) * (size_t) n * n);
    if (o->type == T_SPARSE) {
        for (e = 0; e < o->nnz; ++e) {
            int r = o->row[e], cc = o->col[e]; double a = o->val[e];
            for (k = 0; k < n; ++k) FULL(B, n, cc, k) += a * FULL(Si, n, k, r);
            if (r != cc) for (k = 0; k < n; ++k) FULL(B, n, r, k) += a * FULL(Si, n, k, cc);
        }
        if (rd == 0.0) return 0.0;
        if (o->nnz > 0.1 * n) { for (j = 0; j < n; ++j) for (i = 0; i < n; ++i) t += FULL(Si, n, i, j) * FULL(B, n, i, j); }
        else for (e = 0; e < o->nnz; ++e) {
            int r = o->row[e], cc = o->col[e]; double d = 0; for (k = 0; k < n; ++k) d += FULL(Si, n, k, cc) * FULL(Si, n, k, r);
            t += o->val[e] * d; if (r != cc) t += o->val[e] * d;
        }
        return t;
    }
    if (o->type == T_DENSE) {
        for (j = 0; j < n; ++j) for (i = 0; i < n; ++i) { double s = 0; for (k = 0; k < n; ++k) s += packed_entry(o, i, k) * FULL(Si, n, k, j); FULL(B, n, i, j) = s; }
        if (rd == 0.0) return 0.0;
        for (j = 0; j < n; ++j) for (i = 0; i < n; ++i) t += FULL(Si, n, i, j) * FULL(B, n, i, j);
        return t;
    }
    if (o->type == T_SPR1 && !(o->nfac >= 0.5 * sqrt((double) n))) {
        for (i = 0; i < o->nfac; ++i) for (j = 0; j < o->nfac; ++j) {
            double a = o->sign * o->fv[i] * o->fv[j];
            for (k = 0; k < n; ++k) FULL(B, n, o->idx[i], k) += a * FULL(Si, n, k, o->idx[j]);
            if (i <= j && rd != 0.0) {
                double d = 0; for (k = 0; k < n; ++k) d += FULL(Si, n, k, o->idx[i]) * FULL(Si, n, k, o->idx[j]);
                t += (i == j ? 0.5 : 1.0) * a * d;
            }
        }
        return 2.0 * t;
    }
    /* rank-one via v = S^-1 a: B = sign * a v' */
    k2_solve(c, o, aux);
    for (j = 0; j < n; ++j) for (i = 0; i < n; ++i) FULL(B, n, i, j) += o->sign * o->fac[i] * aux[j];
    if (rd == 0.0) return 0.0;
    return o->sign * ddot(n, aux, aux);
}
/* tr(A S^-1 ASinv) : hdsdp_sdpdata.c:1559-1690 */
static double k4_dot(const ocone *c, const ocoef *o, const double *AS, double *aux) {
    int n = c->n, e, i, j, k; const double *Si = c->Sinv; double s = 0;
    switch (o->type) {
        case T_SPARSE:
            for (e = 0; e < o->nnz; ++e) {
                int r = o->row[e], cc = o->col[e]; double d = 0;
                for (k = 0; k < n; ++k) d += FULL(Si, n, k, r) * FULL(AS, n, k, cc);
                s += o->val[e] * d;
                if (r != cc) { d = 0; for (k = 0; k < n; ++k) d += FULL(Si, n, k, cc) * FULL(AS, n, k, r); s += o->val[e] * d; }
            }
            return s;
        case T_DENSE:
            for (j = 0; j < n; ++j) for (i = j; i < n; ++i) {
                double a = packed_entry(o, i, j), d = 0;
                if (i != j && !(fabs(a) >= 1e-15)) continue;
                for (k = 0; k < n; ++k) d += FULL(Si, n, k, i) * FULL(AS, n, k, j);
                s += (i == j ? 0.5 : 1.0) * a * d;
            }
            return 2.0 * s;
        case T_SPR1:
            if (o->nfac >= sqrt((double) n)) {
                double *z = aux + n; k2_solve(c, o, aux);
                memset(z, 0, sizeof(double) * n);
                for (e = 0; e < o->nfac; ++e) for (k = 0; k < n; ++k) z[k] += o->fv[e] * FULL(AS, n, k, o->idx[e]);
                return o->sign * ddot(n, aux, z);
            }
            for (i = 0; i < o->nfac; ++i) {
                double d;
                for (j = 0; j < i; ++j) { d = 0; for (k = 0; k < n; ++k) d += FULL(Si, n, k, o->idx[i]) * FULL(AS, n, k, o->idx[j]); s += o->fv[i] * o->fv[j] * d; }
                d = 0; for (k = 0; k < n; ++k) d += FULL(Si, n, k, o->idx[i]) * FULL(AS, n, k, o->idx[i]);
                s += 0.5 * o->fv[i] * o->fv[i] * d;
            }
            return 2.0 * o->sign * s;
        case T_DSR1: {
            double *z = aux + n; k2_solve(c, o, aux);
            for (i = 0; i < n; ++i) { double t = 0; for (j = 0; j < n; ++j) t += FULL(AS, n, i, j) * o->fac[j]; z[i] = t; }
            return o->sign * ddot(n, aux, z);
        }
        default: return 0.0;
    }
}
/* tr(A S^-1 B S^-1), pair kernels : hdsdp_sdpdata.c:1711-2058.  X is the matrix playing S^-1. */
static double k5_pair(const ocone *c, const ocoef *A, const ocoef *B, const double *X, double *aux) {
    int n = c->n, e, f, i, j; double s = 0;
    if (A->type == T_ZERO || B->type == T_ZERO) return 0.0;
    if (A->type == T_SPARSE && B->type == T_SPARSE) {
        for (e = 0; e < A->nnz; ++e) {
            int r = A->row[e], cc = A->col[e]; double buf = 0;
            for (f = 0; f < B->nnz; ++f) {
                int r2 = B->row[f], c2 = B->col[f];
                buf += B->val[f] * FULL(X, n, r2, r) * FULL(X, n, c2, cc);
                if (r2 != c2) buf += B->val[f] * FULL(X, n, c2, r) * FULL(X, n, r2, cc);
            }
            s += (r == cc ? 0.5 : 1.0) * A->val[e] * buf;
        }
        return 2.0 * s;
    }
    if (A->type == T_SPARSE && B->type == T_DENSE) {
        for (e = 0; e < A->nnz; ++e) {
            int r = A->row[e], cc = A->col[e]; double buf = 0;
            for (j = 0; j < n; ++j) {
                buf += packed_entry(B, j, j) * FULL(X, n, j, r) * FULL(X, n, j, cc);
                for (i = j + 1; i < n; ++i) { double b = packed_entry(B, i, j); buf += b * FULL(X, n, i, r) * FULL(X, n, j, cc) + b * FULL(X, n, j, r) * FULL(X, n, i, cc); }
            }
            s += (r == cc ? 0.5 : 1.0) * A->val[e] * buf;
        }
        return 2.0 * s;
    }
    if (A->type == T_SPARSE && B->type == T_SPR1) {
        for (e = 0; e < A->nnz; ++e) {
            int r = A->row[e], cc = A->col[e]; double buf = 0;
            for (i = 0; i < B->nfac; ++i) {
                for (j = 0; j < i; ++j) { double b = B->fv[i] * B->fv[j]; buf += b * FULL(X, n, B->idx[i], r) * FULL(X, n, B->idx[j], cc) + b * FULL(X, n, B->idx[j], r) * FULL(X, n, B->idx[i], cc); }
                buf += B->fv[i] * B->fv[i] * FULL(X, n, B->idx[i], r) * FULL(X, n, B->idx[i], cc);
            }
            s += (r == cc ? 0.5 : 1.0) * A->val[e] * buf;
        }
        return 2.0 * B->sign * s;
    }
    if (A->type == T_SPR1 && B->type == T_SPARSE) return k5_pair(c, B, A, X, aux);
    if (A->type == T_SPARSE && B->type == T_DSR1) { symv(n, X, B->fac, aux); return B->sign * k2_quad(A, aux); }
    if (A->type == T_DSR1 && B->type == T_SPARSE) { symv(n, X, A->fac, aux); return A->sign * k2_quad(B, aux); }
    if (A->type == T_SPR1 && B->type == T_DENSE) {
        memset(aux, 0, sizeof(double) * n);
        for (e = 0; e < A->nfac; ++e) for (i = 0; i < n; ++i) aux[i] += A->fv[e] * FULL(X, n, i, A->idx[e]);
        return A->sign * k2_quad(B, aux);
    }
    if (A->type == T_SPR1 && B->type == T_SPR1) {
        for (i = 0; i < A->nfac; ++i) for (j = 0; j < B->nfac; ++j) s += A->fv[i] * B->fv[j] * FULL(X, n, A->idx[i], B->idx[j]);
        return s * s * A->sign * B->sign;
    }
    if ((A->type == T_SPR1 && B->type == T_DSR1) || (A->type == T_DSR1 && B->type == T_SPR1)) {
        const ocoef *sp = A->type == T_SPR1 ? A : B, *ds = A->type == T_SPR1 ? B : A;
        memset(aux, 0, sizeof(double) * n);
        for (e = 0; e < sp->nfac; ++e) for (i = 0; i < n; ++i) aux[i] += sp->fv[e] * FULL(X, n, i, sp->idx[e]);
        s = ddot(n, aux, ds->fac);
        return sp->sign * ds->sign * s * s;
    }
    if (A->type == T_DSR1 && B->type == T_DENSE) { symv(n, X, A->fac, aux); return A->sign * k2_quad(B, aux); }
    if (A->type == T_DSR1 && B->type == T_DSR1) { symv(n, X, A->fac, aux); s = ddot(n, B->fac, aux); return A->sign * B->sign * s * s; }
    /* dense x anything: the reference asserts (hdsdp_sdpdata.c:1997-2001); never selected for dense rows */
    fprintf(stderr, "oracle: M5 invoked on a dense coefficient (unsupported in the reference)\n");
    return NAN;
}
/* tr(S^-1 A S^-1) : hdsdp_sdpdata.c:2068-2165 */
static double k5_sinvadotsinv(const ocone *c, const ocoef *o, double *aux) {
    int n = c->n, d, e, i, j; const double *X = c->Sinv; double s = 0;
    switch (o->type) {
        case T_SPARSE: for (d = 0; d < n; ++d) for (e = 0; e < o->nnz; ++e) s += (o->row[e] == o->col[e] ? 0.5 : 1.0) * o->val[e] * FULL(X, n, o->row[e], d) * FULL(X, n, o->col[e], d); return 2.0 * s;
        case T_DENSE:
            for (i = 0; i < n; ++i) { for (j = 0; j <= i; ++j) { double dd = 0; for (d = 0; d < n; ++d) dd += FULL(X, n, d, i) * FULL(X, n, d, j); s += (i == j ? 0.5 : 1.0) * packed_entry(o, i, j) * dd; } }
            return 2.0 * s;
        case T_SPR1:
            for (d = 0; d < n; ++d) for (i = 0; i < o->nfac; ++i) {
                for (j = 0; j < i; ++j) s += o->fv[i] * o->fv[j] * FULL(X, n, o->idx[i], d) * FULL(X, n, o->idx[j], d);
                s += 0.5 * o->fv[i] * o->fv[i] * FULL(X, n, o->idx[i], d) * FULL(X, n, o->idx[i], d);
            }
            return 2.0 * s * o->sign;
        case T_DSR1: symv(n, X, o->fac, aux); return o->sign * ddot(n, aux, aux);
        default: return 0.0;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Schur column builders.  `rows[pos]` enumerates the visited constraints in visiting order;
 * for the dense cone pos -> perm[pos], for the sparse cone pos -> rowidx[pos].
 * ------------------------------------------------------------------------------------------- */
typedef struct { int m; double *M, *asinv, *asinvrd, *asinvc; double *scal; } okkt; /* scal: CSinvCSinv, CSinv, CSinvRdSinv, TraceSinv */

static void m_add(okkt *k, int a, int b, double v) { int r = a >= b ? a : b, c = a >= b ? b : a; FULL(k->M, k->m, r, c) += v; }

static void column(ocone *c, okkt *k, const int *order, int cnt, int pos, int strat, int typeKKT) {
    int n = c->n, q, i; int ci = order[pos]; const ocoef *A = &c->rows[ci]; int hsd = (typeKKT == K_HOMOGENEOUS);
    double *aux = c->B2, *B = c->B1, sg;
    switch (strat) {
        case M2: { /* hdsdp_conic_sdp.c:687-778 */
            double *v = B; k2_solve(c, A, v); sg = A->sign;
            k->asinv[ci] += k2_trace(A, v);
            if (c->rd) { double nv = sqrt(ddot(n, v, v)); k->asinvrd[ci] += sg * c->rd * nv * nv; }
            if (hsd) k->asinvc[ci] += sg * k2_quad(&c->obj, v);
            for (q = pos; q < cnt; ++q) m_add(k, order[q], ci, sg * k2_quad(&c->rows[order[q]], v));
            break; }
        case M3: { /* :780-851 */
            k->asinv[ci] += k3_sinvasinv(c, A, aux, B);
            if (c->rd) { double t = 0; for (i = 0; i < n; ++i) t += FULL(B, n, i, i); k->asinvrd[ci] += t * c->rd; }
            if (hsd) k->asinvc[ci] += k3_dot(&c->obj, B, aux);
            for (q = pos; q < cnt; ++q) m_add(k, order[q], ci, k3_dot(&c->rows[order[q]], B, aux));
            break; }
        case M4: { /* :853-921 */
            double t = 0;
            k->asinvrd[ci] += c->rd * k4_asinv(c, A, c->rd, aux, B);
            for (i = 0; i < n; ++i) t += FULL(B, n, i, i);
            k->asinv[ci] += t;
            if (hsd) k->asinvc[ci] += k4_dot(c, &c->obj, B, aux);
            for (q = pos; q < cnt; ++q) m_add(k, order[q], ci, k4_dot(c, &c->rows[order[q]], B, aux));
            break; }
        case M5: { /* :923-985 */
            k->asinv[ci] += k3_dot(A, c->Sinv, aux);
            if (c->rd) k->asinvrd[ci] += k5_sinvadotsinv(c, A, aux) * c->rd;
            if (hsd) k->asinvc[ci] += k5_pair(c, A, &c->obj, c->Sinv, aux);
            for (q = pos; q < cnt; ++q) m_add(k, order[q], ci, k5_pair(c, A, &c->rows[order[q]], c->Sinv, aux));
            break; }
        default: break;
    }
}

/* hdsdp_conic_sdp.c:987-1033 (including the SPR1||SPR1 test that sends sparse C down the M3 branch) */
static void hsd_components(ocone *c, okkt *k) {
    int n = c->n, i; const ocoef *C = &c->obj;
    if (C->type == T_ZERO) return;
    if (C->type == T_SPR1) {
        k->scal[0] += k5_pair(c, C, C, c->Sinv, c->B1);
        k->scal[1] += k3_dot(C, c->Sinv, c->B1);
        if (c->rd) k->scal[2] += c->rd * k5_sinvadotsinv(c, C, c->B1);
    } else {
        k->scal[1] += k3_sinvasinv(c, C, c->B2, c->B1);
        k->scal[0] += k3_dot(C, c->B1, c->B2);
        if (c->rd) { double t = 0; for (i = 0; i < n; ++i) t += FULL(c->B1, n, i, i); k->scal[2] += t * c->rd; }
    }
}

/* Driver: hdsdp_conic_sdp.c:1726-1812 (dense cone) / :1814-1886 (sparse cone).
 * fixed_strategy < 0: the reference's automatic choice; primalX: n x n "S^-1" for K_PRIMAL (may be NULL otherwise).
 * M (m x m column-major, lower), vectors and scal[4] are ACCUMULATED into (callers clean them, hdsdp_schur.c:141-165). */
int orc_cone_build_schur(void *h, int typeKKT, int fixed_strategy, const double *primalX, int m, double *M, double *asinv,
                         double *asinvrd, double *asinvc, double *scal) {
    ocone *c = (ocone *) h; int n = c->n, i, pos; okkt k; k.m = m; k.M = M; k.asinv = asinv; k.asinvrd = asinvrd; k.asinvc = asinvc; k.scal = scal;
    if (typeKKT == K_PRIMAL) { if (!primalX) return 1; memcpy(c->Sinv, primalX, sizeof(double) * (size_t) n * n); }
    else { if (!c->factored) return 1; orc_invert(n, c->L, c->Sinv); }
    if (typeKKT == K_CORRECTOR) { /* :1035-1056 */
        for (i = 0; i < c->m; ++i) asinv[i] += k3_dot(&c->rows[i], c->Sinv, c->B1);
        if (c->rd) for (i = 0; i < c->m; ++i) asinvrd[i] += c->rd * k5_sinvadotsinv(c, &c->rows[i], c->B1);
        return 0;
    }
    if (c->rd) for (i = 0; i < n; ++i) scal[3] += FULL(c->Sinv, n, i, i);
    if (c->dense_cone) {
        for (pos = 0; pos < c->m; ++pos) {
            int s = fixed_strategy >= 0 ? fixed_strategy : c->strat[pos];
            if (c->rows[c->perm[pos]].type == T_ZERO) continue;
            if (s == M2 && typeKKT == K_PRIMAL) s = M5;
            column(c, &k, c->perm, c->m, pos, s, typeKKT);
        }
    } else {
        for (pos = 0; pos < c->nrowelem; ++pos) {
            int t = c->rows[c->rowidx[pos]].type, s;
            s = fixed_strategy >= 0 ? fixed_strategy : (t == T_SPARSE || t == T_SPR1) ? M5 : t == T_DENSE ? M3 : (typeKKT == K_PRIMAL ? M5 : M2);
            column(c, &k, c->rowidx, c->nrowelem, pos, s, typeKKT);
        }
    }
    if (typeKKT == K_HOMOGENEOUS) hsd_components(c, &k);
    return 0;
}

/* LP cone: hdsdp_conic_lp.c:254-330.  beg/idx/elem: user CSC [ncol x (m+1)]; s: LP slack (colDual). */
void orc_lp_schur(int m, int ncol, const int *beg, const int *idx, const double *elem, const double *s, double rd, int typeKKT,
                  double *M, double *asinv, double *asinvrd, double *asinvc, double *scal) {
    int r, q, e, f; double *w = (double *) calloc(ncol, sizeof(double));
    for (r = 0; r < m; ++r) for (e = beg[r + 1]; e < beg[r + 2]; ++e) asinv[r] += elem[e] / s[idx[e]];
    if (rd) {
        for (e = 0; e < ncol; ++e) scal[3] += 1.0 / s[e];
        for (r = 0; r < m; ++r) for (e = beg[r + 1]; e < beg[r + 2]; ++e) asinvrd[r] += elem[e] * rd / (s[idx[e]] * s[idx[e]]);
    }
    if (typeKKT == K_CORRECTOR) { free(w); return; }
    for (r = 0; r < m; ++r) {
        memset(w, 0, sizeof(double) * ncol);
        for (e = beg[r + 1]; e < beg[r + 2]; ++e) w[idx[e]] = elem[e] / (s[idx[e]] * s[idx[e]]);
        for (q = 0; q <= r; ++q) { double v = 0; for (f = beg[q + 1]; f < beg[q + 2]; ++f) v += elem[f] * w[idx[f]]; FULL(M, m, r, q) += v; }
    }
    if (typeKKT == K_HOMOGENEOUS) {
        for (e = beg[0]; e < beg[1]; ++e) { double cs = elem[e] / s[idx[e]]; scal[1] += cs; scal[0] += cs * cs; w[idx[e]] = 0; }
        memset(w, 0, sizeof(double) * ncol);
        for (e = beg[0]; e < beg[1]; ++e) w[idx[e]] = elem[e] / (s[idx[e]] * s[idx[e]]);
        for (r = 0; r < m; ++r) for (e = beg[r + 1]; e < beg[r + 2]; ++e) asinvc[r] += elem[e] * w[idx[e]];
    }
    free(w);
}

/* bound cone: hdsdp_conic_bound.c:201-249 (sl = y - l type slacks supplied by the caller) */
void orc_bound_schur(int m, const double *sl, const double *su, int typeKKT, double *M, double *asinv) {
    int i;
    for (i = 0; i < m; ++i) asinv[i] += 1.0 / su[i] - 1.0 / sl[i];
    if (typeKKT == K_CORRECTOR) return;
    for (i = 0; i < m; ++i) FULL(M, m, i, i) += 1.0 / (sl[i] * sl[i]) + 1.0 / (su[i] * su[i]);
}

/* HKKTRegularize: hdsdp_schur.c:348-373 */
void orc_regularize(int m, double *M, double reg) {
    int i; double mind = 1e30;
    for (i = 0; i < m; ++i) if (FULL(M, m, i, i) < mind) mind = FULL(M, m, i, i);
    reg *= mind; if (reg > 1e-05) reg = 1e-05; if (reg < 1e-14) reg = 0.0;
    for (i = 0; i < m; ++i) FULL(M, m, i, i) += reg;
}

/* direct solve of M x = rhs through Cholesky (what the reference's PCG converges to, hdsdp_linsolver.c:1446-1588) */
int orc_kkt_solve(int m, const double *M, const double *rhs, double *x) {
    double *L = (double *) malloc(sizeof(double) * (size_t) m * m); int info = orc_potrf(m, M, L);
    if (!info) { memcpy(x, rhs, sizeof(double) * m); orc_fsolve(m, L, x); orc_bsolve(m, L, x); }
    free(L); return info;
}

/* one "iteration" of the hot path on the CPU, timed by bench.py's port baseline: returns 0 on success */
int orc_iteration(void *h, const double *y, double tau, int m, double *M, double *asinv, double *asinvrd, double *asinvc,
                  double *scal, double reg, const double *rhs, double *x) {
    double ld;
    memset(M, 0, sizeof(double) * (size_t) m * m); memset(asinv, 0, sizeof(double) * m); memset(asinvrd, 0, sizeof(double) * m);
    memset(scal, 0, sizeof(double) * 4);
    if (!orc_cone_set_point(h, y, tau, &ld)) return 1;
    if (orc_cone_build_schur(h, K_INFEASIBLE, -1, NULL, m, M, asinv, asinvrd, asinvc, scal)) return 2;
    if (reg > 0) orc_regularize(m, M, reg);
    return orc_kkt_solve(m, M, rhs, x) ? 3 : 0;
}
