/*
 * hess_oracle.c -- CPU oracle for the Hessenberg hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library (oracle/liboracle.so). The product library (libstarneig.so) never
 * links, loads or calls it.
 *
 * Two CPU solvers and the reference test driver's generators and checks:
 *
 *  oracle_hessenberg_port()   "port": a dense (untiled) restatement of the reference's blocked
 *      algorithm: panel/column loop of src/hessenberg/core.c:399-587, deferred updates of
 *      core.c:301-349, two-phase W updates of core.c:95-266, and -- statement by statement -- the
 *      codelet arithmetic of src/hessenberg/cpu.c:50-560 (same CBLAS/LAPACK calls, same operand
 *      offsets). The only difference from the reference is that one "tile" spans the whole window,
 *      i.e. the summation order inside a GEMV/GEMM is OpenBLAS's rather than per-tile.
 *  oracle_hessenberg_lapack() the reference test driver's own `--solver lapack`
 *      (test/hessenberg/solvers.c:227-271): dgehrd_ + dormhr_("Right","No transpose") + zeroing.
 *
 * Pinning (see oracle/README.md, DESIGN.md "Oracle"): the reference ships NO golden vectors for this
 * path; its tests are invariants. The port is pinned against (i) the reference's OWN
 * sources compiled from /root/reference with a StarPU stand-in (oracle/_ref, built by
 * oracle/Makefile; tests/test_oracle.py compares entrywise) with outputs committed as fixtures under
 * tests/golden/, (ii) LAPACK dgehrd/dormhr, and (iii) the reference driver's invariants
 * (exact-zero Hessenberg form, residual and orthogonality thresholds).
 *
 * BLAS/LAPACK: OpenBLAS bundled with scipy (symbols carry a scipy_ prefix), see ref_shim/cblas.h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "ref_shim/cblas.h"

#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))

extern void dlarfg_(int const *, double *, double *, int const *, double *);
extern void dgehrd_(int const *, int const *, int const *, double *, int const *, double *, double *, int const *, int *);
extern void dormhr_(char const *, char const *, int const *, int const *, int const *, int const *, double const *,
    int const *, double const *, double *, int const *, double *, int const *, int *);
extern void dhseqr_(char const *, char const *, int const *, int const *, int const *, double *, int const *,
    double *, double *, double *, int const *, double *, int const *, int *);

/* ------------------------------------------------------------------------------------------------
 * generators -- test/common/common.c:48-59 (LCG), test/common/init.c:95-120 (fullpos, full)
 * ---------------------------------------------------------------------------------------------- */

#define PRAND_MAX 0x7fffffff              /* test/common/common.h:102 */
static unsigned long prand_seed = 2019;   /* test/common/common.c:48 */

void oracle_prand_init(unsigned int seed) { prand_seed = seed; }

int oracle_prand(void)
{
    return (prand_seed = ((prand_seed * 1103515245) + 12345) & 0x7fffffff);
}

/* crawl_random_fullpos_dr, test/common/init.c:108-120: column by column, entries in [0,1] */
void oracle_fill_fullpos(int n, double *A, int ldA)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            A[(size_t)i * ldA + j] = 1.0 * oracle_prand() / PRAND_MAX;
}

/* crawl_random_full_dr, test/common/init.c:95-106: entries in [-1,1] */
void oracle_fill_full(int n, double *A, int ldA)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            A[(size_t)i * ldA + j] = 2.0 * (1.0 * oracle_prand() / PRAND_MAX) - 1.0;
}

void oracle_fill_identity(int n, double *Q, int ldQ)
{
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) Q[(size_t)i * ldQ + j] = 0.0;
        Q[(size_t)i * ldQ + i] = 1.0;
    }
}

/* test/misc/partial_hessenberg.c:142-157: upper triangular with a full diagonal block [begin,end) */
void oracle_fill_partial(int n, int begin, int end, double *A, int ldA)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            A[(size_t)i * ldA + j] = j <= i ? 2.0 * (1.0 * oracle_prand() / PRAND_MAX) - 1.0 : 0.0;
    for (int i = begin; i < end - 1; i++)
        for (int j = i + 1; j < end; j++)
            A[(size_t)i * ldA + j] = 2.0 * (1.0 * oracle_prand() / PRAND_MAX) - 1.0;
}

void oracle_set_threads(int threads) { openblas_set_num_threads(threads); }
int oracle_get_threads(void) { return openblas_get_num_threads(); }

/* src/hessenberg/interface.c:74-78 */
int oracle_default_panel_width(int n)
{
    int w = (int)ceil((0.001875596476 * n + 273.5908216) / 8.0) * 8;
    return MAX(64, w);
}

/* ------------------------------------------------------------------------------------------------
 * "port": dense restatement of src/hessenberg/cpu.c driven by the loops of src/hessenberg/core.c
 * ---------------------------------------------------------------------------------------------- */

/* cpu.c:50-161 -- update panel column i (0-based inside the panel), form the reflector */
static void prepare_column(int i, int m, int nb, double *Y, int ldY, double *V, int ldV,
    double *T, int ldT, double *P, int ldP)
{
    double *p = P + (size_t)i * ldP;
    if (0 < i) {
        /* cpu.c:98-99   p <- p - Y * V(i-1,:)^T */
        cblas_dgemv(CblasColMajor, CblasNoTrans, m, i, -1.0, Y, ldY, V + i - 1, ldV, 1.0, p, 1);
        /* cpu.c:106     last column of T is the work space */
        double *w = T + (size_t)(nb - 1) * ldT;
        /* cpu.c:109-111 w <- V1^T b1 */
        cblas_dcopy(i, p, 1, w, 1);
        cblas_dtrmv(CblasColMajor, CblasLower, CblasTrans, CblasUnit, i, V, ldV, w, 1);
        /* cpu.c:114-115 w <- w + V2^T b2 */
        cblas_dgemv(CblasColMajor, CblasTrans, m - i, i, 1.0, V + i, ldV, p + i, 1, 1.0, w, 1);
        /* cpu.c:118-120 w <- T^T w */
        cblas_dtrmv(CblasColMajor, CblasUpper, CblasTrans, CblasNonUnit, i, T, ldT, w, 1);
        /* cpu.c:123-124 b2 <- b2 - V2 w */
        cblas_dgemv(CblasColMajor, CblasNoTrans, m - i, i, -1.0, V + i, ldV, w, 1, 1.0, p + i, 1);
        /* cpu.c:127-130 b1 <- b1 - V1 w */
        cblas_dtrmv(CblasColMajor, CblasLower, CblasNoTrans, CblasUnit, i, V, ldV, w, 1);
        cblas_daxpy(i, -1.0, w, 1, p, 1);
    }
    /* cpu.c:137-141 reflector */
    int height = m - i, one = 1;
    double tau, *v = V + (size_t)i * ldV + i;
    memcpy(v, p + i, height * sizeof(double));
    dlarfg_(&height, p + i, v + 1, &one, &tau);
    v[0] = 1.0;
    /* cpu.c:153-154 exact zeros below the sub-diagonal */
    for (int j = i + 1; j < m; j++) p[j] = 0.0;
    /* cpu.c:160 */
    T[(size_t)i * ldT + i] = tau;
}

/* cpu.c:226-285 -- Y(:,i) and T(:,i); y holds A*v on entry */
static void finish_column(int i, int m, double const *y, double *V, int ldV, double *T, int ldT, double *Y, int ldY)
{
    double tau = T[(size_t)i * ldT + i];
    double *v = V + (size_t)i * ldV + i;
    memcpy(Y + (size_t)i * ldY, y, m * sizeof(double));                       /* cpu.c:260 */
    cblas_dgemv(CblasColMajor, CblasTrans, m - i, i, 1.0, V + i, ldV, v, 1, 0.0, T + (size_t)i * ldT, 1);   /* :263 */
    cblas_dgemv(CblasColMajor, CblasNoTrans, m, i, -1.0, Y, ldY, T + (size_t)i * ldT, 1, 1.0, Y + (size_t)i * ldY, 1); /* :267 */
    cblas_dscal(m, tau, Y + (size_t)i * ldY, 1);                               /* :270 */
    cblas_dscal(i, -tau, T + (size_t)i * ldT, 1);                              /* :277 */
    cblas_dtrmv(CblasColMajor, CblasUpper, CblasNoTrans, CblasNonUnit, i, T, ldT, T + (size_t)i * ldT, 1);  /* :280 */
    T[(size_t)i * ldT + i] = tau;                                              /* :284 */
}

/* core.c:198-266 with cpu.c:443-560: X(rows, cols) <- X - (X V T) V^T, V rows <-> cols */
static void right_update(int rows, int cols, int nb, double const *V, int ldV, double const *T, int ldT,
    double *X, int ldX, double *W, double *Pw)
{
    if (rows < 1 || cols < 1 || nb < 1) return;
    int ldW = rows;
    cblas_dgemm(CblasColMajor, CblasNoTrans, CblasNoTrans, rows, nb, cols, 1.0, X, ldX, V, ldV, 0.0, Pw, ldW);   /* cpu.c:492 */
    cblas_dtrmm(CblasColMajor, CblasRight, CblasUpper, CblasNoTrans, CblasNonUnit, rows, nb, 1.0, T, ldT, Pw, ldW); /* :497 */
    memcpy(W, Pw, (size_t)rows * nb * sizeof(double));                         /* W (zero) += P, cpu.c:502-503 */
    cblas_dgemm(CblasColMajor, CblasNoTrans, CblasTrans, rows, cols, nb, -1.0, W, ldW, V, ldV, 1.0, X, ldX);     /* :552 */
}

/* core.c:95-163 with cpu.c:324-441: X(rows, cols) <- X - V (X^T V T)^T, V rows <-> rows */
static void left_update(int rows, int cols, int nb, double const *V, int ldV, double const *T, int ldT,
    double *X, int ldX, double *W, double *Pw)
{
    if (rows < 1 || cols < 1 || nb < 1) return;
    int ldW = cols;
    cblas_dgemm(CblasColMajor, CblasTrans, CblasNoTrans, cols, nb, rows, 1.0, X, ldX, V, ldV, 0.0, Pw, ldW);     /* cpu.c:373 */
    cblas_dtrmm(CblasColMajor, CblasRight, CblasUpper, CblasNoTrans, CblasNonUnit, cols, nb, 1.0, T, ldT, Pw, ldW); /* :378 */
    memcpy(W, Pw, (size_t)cols * nb * sizeof(double));                         /* cpu.c:383-384 */
    cblas_dgemm(CblasColMajor, CblasNoTrans, CblasTrans, rows, cols, nb, -1.0, V, ldV, W, ldW, 1.0, X, ldX);     /* :433 */
}

struct deferred { int i, nb, m; double *P, *V, *T; };

/* starneig_hessenberg_insert_tasks, core.c:351-599 (dense). Returns 0, or 3 for an invalid panel width. */
int oracle_hessenberg_port(int n, int begin, int end, int panel_width,
    double *A, int ldA, double *Q, int ldQ)
{
    if (panel_width < 0) panel_width = oracle_default_panel_width(n);
    if (panel_width < 8) return 3;      /* STARNEIG_INVALID_CONFIGURATION, interface.c:80-83 */

    int npanels = 0;
    for (int i = begin; i < end - 1; i += panel_width) npanels++;
    struct deferred *updates = calloc(npanels > 0 ? npanels : 1, sizeof(*updates));
    double *W = malloc((size_t)n * panel_width * sizeof(double));
    double *Pw = malloc((size_t)n * panel_width * sizeof(double));
    double *y = malloc((size_t)n * sizeof(double));
    int count = 0;

    for (int i = begin; i < end - 1; i += panel_width) {                    /* core.c:399 */
        const int nb = MIN(panel_width, end - i - 1);                        /* :400 */
        const int m = end - i - 1;
        const int ld = m;
        double *P = malloc((size_t)m * nb * sizeof(double));                 /* :428-436 */
        double *V = calloc((size_t)m * nb, sizeof(double));                  /* :455 set to zero */
        double *Y = malloc((size_t)m * nb * sizeof(double));
        double *T = calloc((size_t)nb * nb, sizeof(double));
        for (int j = 0; j < nb; j++)                                          /* :451 copy panel */
            memcpy(P + (size_t)j * ld, A + (size_t)(i + j) * ldA + i + 1, m * sizeof(double));

        for (int j = 0; j < nb; j++) {                                        /* :461 */
            prepare_column(j, m, nb, Y, ld, V, ld, T, nb, P, ld);             /* :479 */
            /* :486-506 with cpu.c:217-219: y = A(i+1:end, i+j+1:end) * v, v = V(j:, j) */
            cblas_dgemv(CblasColMajor, CblasNoTrans, m, end - (i + j + 1), 1.0,
                A + (size_t)(i + j + 1) * ldA + i + 1, ldA, V + (size_t)j * ld + j, 1, 0.0, y, 1);
            finish_column(j, m, y, V, ld, T, nb, Y, ld);                      /* :512 */
        }

        /* :523-540 with cpu.c:315-316: A(i+1:end, i+nb:end) -= Y * V(nb-1:, :)^T */
        if (end - (i + nb) > 0)
            cblas_dgemm(CblasColMajor, CblasNoTrans, CblasTrans, m, end - (i + nb), nb, -1.0,
                Y, ld, V + nb - 1, ld, 1.0, A + (size_t)(i + nb) * ldA + i + 1, ldA);

        /* :546-547 trailing left update */
        left_update(m, end - (i + nb), nb, V, ld, T, nb, A + (size_t)(i + nb) * ldA + i + 1, ldA, W, Pw);

        free(Y);                                                              /* :549 */
        updates[count].i = i; updates[count].nb = nb; updates[count].m = m;  /* :555-571 */
        updates[count].P = P; updates[count].V = V; updates[count].T = T;
        count++;
    }

    /* insert_remaining, core.c:301-349 */
    for (int k = 0; k < count; k++) {
        int i = updates[k].i, nb = updates[k].nb, m = updates[k].m;
        double *P = updates[k].P, *V = updates[k].V, *T = updates[k].T;
        for (int j = 0; j < nb; j++)                                          /* :317 P -> A */
            memcpy(A + (size_t)(i + j) * ldA + i + 1, P + (size_t)j * m, m * sizeof(double));
        /* :320-327 rows above the panel, columns i+1..end */
        right_update(i + 1, m, nb, V, m, T, nb, A + (size_t)(i + 1) * ldA, ldA, W, Pw);
        /* :329-336 columns right of the reduced block (partial reductions) */
        left_update(m, n - end, nb, V, m, T, nb, A + (size_t)end * ldA + i + 1, ldA, W, Pw);
        /* :338-340 Q */
        right_update(n, m, nb, V, m, T, nb, Q + (size_t)(i + 1) * ldQ, ldQ, W, Pw);
        free(P); free(V); free(T);
    }

    free(updates); free(W); free(Pw); free(y);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * "lapack": test/hessenberg/solvers.c:227-271
 * ---------------------------------------------------------------------------------------------- */
int oracle_hessenberg_lapack(int n, double *A, int ldA, double *Q, int ldQ)
{
    int ilo = 1, ihi = n, info = 0, lwork = -1;
    double dlwork, *work = NULL, *tau = malloc((size_t)n * sizeof(double));
    dgehrd_(&n, &ilo, &ihi, A, &ldA, tau, &dlwork, &lwork, &info);
    if (info != 0) goto cleanup;
    lwork = (int)dlwork; work = malloc((size_t)lwork * sizeof(double));
    dgehrd_(&n, &ilo, &ihi, A, &ldA, tau, work, &lwork, &info);
    if (info != 0) goto cleanup;
    free(work); work = NULL; lwork = -1;
    dormhr_("Right", "No transpose", &n, &n, &ilo, &ihi, A, &ldA, tau, Q, &ldQ, &dlwork, &lwork, &info);
    if (info != 0) goto cleanup;
    lwork = (int)dlwork; work = malloc((size_t)lwork * sizeof(double));
    dormhr_("Right", "No transpose", &n, &n, &ilo, &ihi, A, &ldA, tau, Q, &ldQ, work, &lwork, &info);
    if (info != 0) goto cleanup;
    for (int i = 0; i < n; i++)
        for (int j = i + 2; j < n; j++)
            A[(size_t)i * ldA + j] = 0.0;
cleanup:
    free(work); free(tau);
    return info;
}

/* ------------------------------------------------------------------------------------------------
 * checks -- test/common/hooks.c:434-456, test/misc/partial_hessenberg.c:178-214, test/common/checks.c:180-208
 * ---------------------------------------------------------------------------------------------- */

/* number of entries violating the (partial) Hessenberg form; must be 0. Columns begin..end-2 must be
 * zero below the sub-diagonal; for a partial reduction of an upper-triangular-plus-block input all other
 * columns must be zero below the diagonal (check_outside != 0). */
long oracle_check_hessenberg_form(int n, int begin, int end, int check_outside, double const *A, int ldA)
{
    long failed = 0;
    for (int i = 0; i < n - 1; i++) {
        int inside = begin <= i && i < end - 1;
        if (!inside && !check_outside) continue;
        int k = inside ? 2 : 1;
        for (int j = i + k; j < n; j++)
            if (A[(size_t)i * ldA + j] != 0.0) failed++;
    }
    return failed;
}

/* 2^52 * ||Q H Q^T - A0||_F / ||A0||_F   (compute_qazt_c_norm, checks.c:180-194) */
double oracle_residual_u(int n, double const *Q, int ldQ, double const *H, int ldH, double const *A0, int ldA0)
{
    double *tmp = malloc((size_t)n * n * sizeof(double));
    double *B = malloc((size_t)n * n * sizeof(double));
    for (int j = 0; j < n; j++) memcpy(B + (size_t)j * n, A0 + (size_t)j * ldA0, n * sizeof(double));
    cblas_dgemm(CblasColMajor, CblasNoTrans, CblasNoTrans, n, n, n, 1.0, Q, ldQ, H, ldH, 0.0, tmp, n);
    cblas_dgemm(CblasColMajor, CblasNoTrans, CblasTrans, n, n, n, 1.0, tmp, n, Q, ldQ, -1.0, B, n);
    double nb = 0.0, nc = 0.0;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            double b = B[(size_t)j * n + i], c = A0[(size_t)j * ldA0 + i];
            nb += b * b; nc += c * c;
        }
    free(tmp); free(B);
    return 4503599627370496.0 * sqrt(nb) / sqrt(nc);
}

/* 2^52 * ||Q Q^T - I||_F / sqrt(n)   (compute_qqt_norm, checks.c:196-208) */
double oracle_orthogonality_u(int n, double const *Q, int ldQ)
{
    double *B = calloc((size_t)n * n, sizeof(double));
    for (int i = 0; i < n; i++) B[(size_t)i * n + i] = 1.0;
    cblas_dgemm(CblasColMajor, CblasNoTrans, CblasTrans, n, n, n, 1.0, Q, ldQ, Q, ldQ, -1.0, B, n);
    double nb = 0.0;
    for (size_t k = 0; k < (size_t)n * n; k++) nb += B[k] * B[k];
    free(B);
    return 4503599627370496.0 * sqrt(nb) / sqrt((double)n);
}

/* max_ij |X - Y| and max_ij |Y| over the leading n x n blocks */
void oracle_max_abs_diff(int n, double const *X, int ldX, double const *Y, int ldY, double *maxdiff, double *maxref)
{
    double d = 0.0, r = 0.0;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            double x = X[(size_t)j * ldX + i], y = Y[(size_t)j * ldY + i];
            double e = fabs(x - y);
            if (!(e <= d)) d = e;          /* NaN propagates */
            if (fabs(y) > r) r = fabs(y);
        }
    *maxdiff = d; *maxref = r;
}

double oracle_frobenius(int n, double const *X, int ldX)
{
    double s = 0.0;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) s += X[(size_t)j * ldX + i] * X[(size_t)j * ldX + i];
    return sqrt(s);
}

/* eigenvalues of an upper Hessenberg matrix with LAPACK dhseqr_("E","N") -- the stand-in for the
 * downstream starneig_SEP_SM_Schur stage (src/schur cannot be built here). H is copied. */
int oracle_hessenberg_eigenvalues(int n, double const *H, int ldH, double *wr, double *wi)
{
    int ilo = 1, ihi = n, info = 0, lwork = -1, ldz = 1;
    double *Hc = malloc((size_t)n * n * sizeof(double)), dlwork, z;
    for (int j = 0; j < n; j++) memcpy(Hc + (size_t)j * n, H + (size_t)j * ldH, n * sizeof(double));
    dhseqr_("E", "N", &n, &ilo, &ihi, Hc, &n, wr, wi, &z, &ldz, &dlwork, &lwork, &info);
    lwork = (int)dlwork;
    double *work = malloc((size_t)MAX(1, lwork) * sizeof(double));
    dhseqr_("E", "N", &n, &ilo, &ihi, Hc, &n, wr, wi, &z, &ldz, work, &lwork, &info);
    free(work); free(Hc);
    return info;
}

double oracle_wall_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
