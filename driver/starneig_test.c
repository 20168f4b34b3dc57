/* starneig_test.c -- a test driver for the Hessenberg experiment that speaks the reference driver's language.
 *
 * SURVEY.md section 8(f) rank 2. It is a plain C client of the drop-in boundary: it includes <starneig/starneig.h>,
 * links libstarneig.so and calls starneig_node_init / starneig_SEP_SM_Hessenberg[_expert] / starneig_node_finalize
 * exactly as the reference solver plugins do (test/hessenberg/solvers.c:506-545 "prepare", :548-614 "run" for
 * --solver starneig; :713-765 for --solver starneig-simple). What it keeps from the reference driver:
 *   --experiment hessenberg [--experiment partial-hessenberg]   test/hessenberg/experiment.c, test/misc/partial_hessenberg.c
 *   --n N --seed S --init default|fullpos|full|read-raw --input FMT          test/common/init.c:95-120, common.c:48-59
 *   --solver starneig|starneig-simple|lapack --cores C --gpus G --tile-size T --panel-width W
 *       (lapack = dgehrd + dormhr on the CPU, the reference driver's comparison solver, test/hessenberg/solvers.c:
 *        227-271; it never touches libstarneig.so and exists so that both solvers can be run on the same input)
 *   --begin B --end E                                            (partial reductions through the expert interface)
 *   --hooks hessenberg residual print store-raw --store-raw-output FMT       test/common/hooks.c
 *   --residual-fail-threshold X --residual-warn-threshold X (units of u = 2^-52; defaults 10000 / 500, hooks.c:52-57)
 *   --repeat R --warmup W
 * and the raw matrix format "STARNEIG RAW REAL DOUBLE M %d N %d\n" + column-major doubles (test/common/io.c:236-360),
 * so that inputs and outputs can be exchanged with a real StarNEig build: `store-raw` here, `--init read-raw` there.
 * Output lines follow the reference's wording (docs/_7_test_driver.md:230-250) so that scripts which parse them
 * keep working. The checks use a CPU BLAS (dgemm), like the reference driver; nothing under oracle/ is used.
 *
 * Exit status: 0 all hooks passed (warnings allowed), 1 a hook failed, 2 usage / solver error.
 */
#include <starneig/starneig.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef DRIVER_HAVE_CUDART
#include <cuda_runtime_api.h>
#endif

/* CBLAS prototype; the symbol may carry a vendor prefix (the OpenBLAS bundled with scipy: scipy_cblas_dgemm) */
#ifndef DRIVER_CBLAS_DGEMM
#define DRIVER_CBLAS_DGEMM cblas_dgemm
#endif
void DRIVER_CBLAS_DGEMM(int order, int transa, int transb, int m, int n, int k, double alpha, const double *A, int lda,
                        const double *B, int ldb, double beta, double *C, int ldc);
enum { COL_MAJOR = 102, NO_TRANS = 111, TRANS = 112 };
/* LAPACK, for the comparison solver `--solver lapack` (the reference driver has the same one) */
#ifndef DRIVER_DGEHRD
#define DRIVER_DGEHRD dgehrd_
#define DRIVER_DORMHR dormhr_
#endif
void DRIVER_DGEHRD(const int *n, const int *ilo, const int *ihi, double *A, const int *lda, double *tau, double *work,
                   const int *lwork, int *info);
void DRIVER_DORMHR(const char *side, const char *trans, const int *m, const int *n, const int *ilo, const int *ihi,
                   const double *A, const int *lda, const double *tau, double *C, const int *ldc, double *work,
                   const int *lwork, int *info);

/* ------------------------------------------------------------------------------------------------------------ */
/* arguments                                                                                                      */
/* ------------------------------------------------------------------------------------------------------------ */
static const char *arg_str(int argc, char **argv, const char *name, const char *dflt)
{
    for (int i = 1; i + 1 < argc; i++)
        if (strcmp(argv[i], name) == 0) return argv[i + 1];
    return dflt;
}
static int arg_is_default(const char *s) { return s == NULL || strcmp(s, "default") == 0; }
static int arg_int(int argc, char **argv, const char *name, int dflt)
{
    const char *s = arg_str(argc, argv, name, NULL);
    return arg_is_default(s) ? dflt : atoi(s);
}
static double arg_double(int argc, char **argv, const char *name, double dflt)
{
    const char *s = arg_str(argc, argv, name, NULL);
    return arg_is_default(s) ? dflt : atof(s);
}
/* is `hook` listed after --hooks (up to the next --option)? "name:mode" suffixes are accepted and ignored */
static int hook_enabled(int argc, char **argv, const char *hook, int dflt)
{
    for (int i = 1; i < argc; i++) {
        if (strcmp(argv[i], "--hooks") != 0) continue;
        for (int j = i + 1; j < argc && strncmp(argv[j], "--", 2) != 0; j++) {
            size_t len = strcspn(argv[j], ":");
            if (len == strlen(hook) && strncmp(argv[j], hook, len) == 0) return 1;
        }
        return 0;
    }
    return dflt;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* matrices                                                                                                       */
/* ------------------------------------------------------------------------------------------------------------ */
static int pinned_alloc = 0;
static double *alloc_matrix(int n, int *ld)
{
    *ld = (n + 7) / 8 * 8;                      /* leading dimension rounded up to 64 bytes */
    size_t bytes = (size_t)*ld * n * sizeof(double);
    void *p = NULL;
#ifdef DRIVER_HAVE_CUDART
    /* the reference driver page-locks its matrices when StarNEig is built with CUDA (test/common/common.c:96-112) */
    if (pinned_alloc && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) return (double *)p;
    p = NULL;
#endif
    if (posix_memalign(&p, 64, bytes) != 0) { fprintf(stderr, "Out of memory.\n"); exit(2); }
    return (double *)p;
}
static void free_matrix(double *p)
{
#ifdef DRIVER_HAVE_CUDART
    if (pinned_alloc && p && cudaFreeHost(p) == cudaSuccess) return;
#endif
    free(p);
}

/* the driver's linear congruential generator */
static unsigned long lcg_state = 2019;
static int lcg_next(void) { return (int)(lcg_state = (lcg_state * 1103515245UL + 12345UL) & 0x7fffffffUL); }
static double lcg_unit(void) { return (double)lcg_next() / 2147483647.0; }

static void fill_random(int n, double *A, int ld, int positive)
{
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) A[(size_t)c * ld + r] = positive ? lcg_unit() : 2.0 * lcg_unit() - 1.0;
}
/* partial-hessenberg experiment: upper triangular outside the diagonal block [begin, end), full inside it */
static void fill_partial(int n, int begin, int end, double *A, int ld)
{
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) A[(size_t)c * ld + r] = r <= c ? 2.0 * lcg_unit() - 1.0 : 0.0;
    for (int c = begin; c < end - 1; c++)
        for (int r = c + 1; r < end; r++) A[(size_t)c * ld + r] = 2.0 * lcg_unit() - 1.0;
}
static void set_identity(int n, double *Q, int ld)
{
    for (int c = 0; c < n; c++) {
        memset(Q + (size_t)c * ld, 0, (size_t)n * sizeof(double));
        Q[(size_t)c * ld + c] = 1.0;
    }
}
static void copy_matrix(int n, const double *S, int lds, double *D, int ldd)
{
    for (int c = 0; c < n; c++) memcpy(D + (size_t)c * ldd, S + (size_t)c * lds, (size_t)n * sizeof(double));
}

/* raw format: one text header line, then the columns back to back */
static int write_raw(const char *name, int n, const double *A, int ld)
{
    FILE *f = fopen(name, "wb");
    if (!f) { fprintf(stderr, "Invalid filename.\n"); return -1; }
    printf("WRITING TO %s...\n", name);
    fprintf(f, "STARNEIG RAW REAL DOUBLE M %d N %d\n", n, n);
    for (int c = 0; c < n; c++)
        if (fwrite(A + (size_t)c * ld, sizeof(double), (size_t)n, f) != (size_t)n) { fclose(f); return -1; }
    fclose(f);
    return 0;
}
static int read_raw_header(const char *name, int *m, int *n)
{
    FILE *f = fopen(name, "rb");
    if (!f) { fprintf(stderr, "Invalid filename.\n"); return -1; }
    int ok = fscanf(f, "STARNEIG RAW REAL DOUBLE M %d N %d", m, n) == 2 && *m >= 1 && *n >= 1;
    fclose(f);
    if (!ok) fprintf(stderr, "Invalid file.\n");
    return ok ? 0 : -1;
}
static int read_raw(const char *name, int n, double *A, int ld)
{
    int fm, fn;
    if (read_raw_header(name, &fm, &fn) != 0 || fm != n || fn != n) return -1;
    FILE *f = fopen(name, "rb");
    int ch;
    while ((ch = fgetc(f)) != '\n' && ch != EOF) { }
    printf("READING A %d X %d MATRIX ...\n", n, n);
    for (int c = 0; c < n; c++)
        if (fread(A + (size_t)c * ld, sizeof(double), (size_t)n, f) != (size_t)n) { fclose(f); return -1; }
    fclose(f);
    return 0;
}
/* "hessenberg_%s.dat" + "A" -> "hessenberg_A.dat" */
static void format_name(char *out, size_t cap, const char *fmt, const char *tag)
{
    const char *p = strstr(fmt, "%s");
    if (!p) { snprintf(out, cap, "%s", fmt); return; }
    snprintf(out, cap, "%.*s%s%s", (int)(p - fmt), fmt, tag, p + 2);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* checks (test/common/checks.c:180-208, hooks.c:434-456): all in units of u = 2^-52                              */
/* ------------------------------------------------------------------------------------------------------------ */
static double frobenius(int n, const double *A, int ld)
{
    double scale = 0.0, ssq = 1.0;              /* scaled sum of squares: no overflow for badly scaled matrices */
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) {
            double a = fabs(A[(size_t)c * ld + r]);
            if (a == 0.0) continue;
            if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    return scale * sqrt(ssq);
}
/* |Q H Q^T - A| / |A| */
static double residual_u(int n, const double *Q, int ldQ, const double *H, int ldH, const double *A, int ldA)
{
    int ld;
    double *T = alloc_matrix(n, &ld), *R = alloc_matrix(n, &ld);
    copy_matrix(n, A, ldA, R, ld);
    DRIVER_CBLAS_DGEMM(COL_MAJOR, NO_TRANS, NO_TRANS, n, n, n, 1.0, Q, ldQ, H, ldH, 0.0, T, ld);
    DRIVER_CBLAS_DGEMM(COL_MAJOR, NO_TRANS, TRANS, n, n, n, 1.0, T, ld, Q, ldQ, -1.0, R, ld);
    double res = ldexp(1.0, 52) * frobenius(n, R, ld) / frobenius(n, A, ldA);
    free_matrix(T); free_matrix(R);
    return res;
}
/* |Q Q^T - I| / |I| */
static double orthogonality_u(int n, const double *Q, int ldQ)
{
    int ld;
    double *R = alloc_matrix(n, &ld);
    set_identity(n, R, ld);
    DRIVER_CBLAS_DGEMM(COL_MAJOR, NO_TRANS, TRANS, n, n, n, 1.0, Q, ldQ, Q, ldQ, -1.0, R, ld);
    double res = ldexp(1.0, 52) * frobenius(n, R, ld) / sqrt((double)n);
    free_matrix(R);
    return res;
}
/* entries that must be exactly zero and are not */
static long form_violations(int n, const double *H, int ld, int begin, int end, int partial)
{
    long bad = 0;
    for (int c = 0; c < n; c++) {
        /* partial: one sub-diagonal inside columns [begin, end-1), none elsewhere (partial_hessenberg.c:183-186);
         * otherwise the hessenberg hook's rule: nothing below the first sub-diagonal (hooks.c:442-444) */
        int first = partial ? ((c >= begin && c < end - 1) ? c + 2 : c + 1) : c + 2;
        for (int r = first; r < n; r++)
            if (H[(size_t)c * ld + r] != 0.0) bad++;
    }
    return bad;
}

static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }
static void print_stats(const char *label, const char *unit, int count, double *v)
{
    qsort(v, (size_t)count, sizeof(double), cmp_double);
    double mean = 0.0, var = 0.0;
    for (int i = 0; i < count; i++) mean += v[i] / count;
    for (int i = 0; i < count; i++) var += (v[i] - mean) * (v[i] - mean) / count;
    double cv = (var > 0.0 && mean != 0.0) ? sqrt(var) / mean : 0.0;
    printf("%s = [avg %.0f%s, cv %.2f, min %.0f%s, max %.0f%s]\n", label, mean, unit, cv, v[0], unit, v[count - 1], unit);
}

/* --solver lapack: A <- H, Q <- Q U with dgehrd / dormhr; the reflectors below the sub-diagonal are cleared */
static int lapack_solver(int n, int begin, int end, double *A, int ldA, double *Q, int ldQ)
{
    int ilo = begin + 1, ihi = end, info = 0, lwork = -1;
    double query = 0.0;
    double *tau = (double *)calloc((size_t)n, sizeof(double));
    DRIVER_DGEHRD(&n, &ilo, &ihi, A, &ldA, tau, &query, &lwork, &info);
    lwork = (int)query;
    double *work = (double *)malloc((size_t)(lwork > 1 ? lwork : 1) * sizeof(double));
    if (info == 0) DRIVER_DGEHRD(&n, &ilo, &ihi, A, &ldA, tau, work, &lwork, &info);
    free(work);
    lwork = -1;
    if (info == 0) DRIVER_DORMHR("R", "N", &n, &n, &ilo, &ihi, A, &ldA, tau, Q, &ldQ, &query, &lwork, &info);
    lwork = (int)query;
    work = (double *)malloc((size_t)(lwork > 1 ? lwork : 1) * sizeof(double));
    if (info == 0) DRIVER_DORMHR("R", "N", &n, &n, &ilo, &ihi, A, &ldA, tau, Q, &ldQ, work, &lwork, &info);
    for (int c = begin; c < end; c++)
        for (int r = c + 2; r < end; r++) A[(size_t)c * ldA + r] = 0.0;
    free(work); free(tau);
    return info;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

static void usage(void)
{
    printf("Usage: starneig-test --experiment hessenberg|partial-hessenberg [--n N] [--seed S]\n"
           "  [--init default|fullpos|full|read-raw] [--input FMT] [--begin B] [--end E]\n"
           "  [--solver starneig|starneig-simple|lapack] [--cores C] [--gpus G] [--tile-size T] [--panel-width W]\n"
           "  [--hooks [hessenberg] [residual] [print] [store-raw]] [--store-raw-output FMT]\n"
           "  [--residual-fail-threshold U] [--residual-warn-threshold U] [--repeat R] [--warmup W] [--pinning on|off]\n");
}

int main(int argc, char **argv)
{
    const char *experiment = arg_str(argc, argv, "--experiment", NULL);
    if (experiment == NULL || (strcmp(experiment, "hessenberg") != 0 && strcmp(experiment, "partial-hessenberg") != 0)) {
        usage();
        return 2;
    }
    const int partial = strcmp(experiment, "partial-hessenberg") == 0;
    const char *init = arg_str(argc, argv, "--init", "default");
    const char *input = arg_str(argc, argv, "--input", NULL);
    const char *solver = arg_str(argc, argv, "--solver", "starneig");
    const char *store_fmt = arg_str(argc, argv, "--store-raw-output", "hessenberg_%s.dat");
    unsigned seed = (unsigned)arg_int(argc, argv, "--seed", (int)time(NULL));
    int n = arg_int(argc, argv, "--n", 1000);
    const int repeat = arg_int(argc, argv, "--repeat", 1), warmup = arg_int(argc, argv, "--warmup", 0);
    const int cores = arg_int(argc, argv, "--cores", STARNEIG_USE_ALL), gpus = arg_int(argc, argv, "--gpus", STARNEIG_USE_ALL);
    const double fail_thr = arg_double(argc, argv, "--residual-fail-threshold", 10000.0);
    const double warn_thr = arg_double(argc, argv, "--residual-warn-threshold", 500.0);
    const int h_form = hook_enabled(argc, argv, "hessenberg", 1), h_res = hook_enabled(argc, argv, "residual", 1);
    const int h_print = hook_enabled(argc, argv, "print", 0), h_store = hook_enabled(argc, argv, "store-raw", 0);
    pinned_alloc = strcmp(arg_str(argc, argv, "--pinning", "on"), "off") != 0 && strcmp(solver, "lapack") != 0;
    const int use_lapack = strcmp(solver, "lapack") == 0;
    if (!use_lapack && strcmp(solver, "starneig") != 0 && strcmp(solver, "starneig-simple") != 0) {
        fprintf(stderr, "Invalid solver.\n");
        return 2;
    }
    if (repeat < 1 || warmup < 0) { fprintf(stderr, "Invalid repeat / warmup count.\n"); return 2; }

    char name[1024];
    if (strcmp(init, "read-raw") == 0) {
        int fm, fn;
        if (input == NULL) { fprintf(stderr, "Input file name is missing.\n"); return 2; }
        format_name(name, sizeof(name), input, "A");
        if (read_raw_header(name, &fm, &fn) != 0 || fm != fn) return 2;
        n = fm;
    }
    if (n < 1) { fprintf(stderr, "Invalid matrix dimension.\n"); return 2; }
    int begin = arg_int(argc, argv, "--begin", partial ? n / 4 : 0);
    int end = arg_int(argc, argv, "--end", partial ? 3 * n / 4 : n);
    if (begin < 0 || end < begin || n < end) { fprintf(stderr, "Invalid begin / end.\n"); return 2; }
    const int expert = strcmp(solver, "starneig") == 0;
    if (!expert && !use_lapack && (begin != 0 || end != n)) { fprintf(stderr, "Solver does not support partial reductions.\n"); return 2; }

    printf("TEST: --seed %u --experiment %s --init %s --n %d --begin %d --end %d --solver %s --cores %s --gpus %s "
           "--tile-size %s --panel-width %s --hooks%s%s%s%s --residual-fail-threshold %.0f --residual-warn-threshold %.0f "
           "--repeat %d --warmup %d\n", seed, experiment, init, n, begin, end, solver,
           arg_str(argc, argv, "--cores", "default"), arg_str(argc, argv, "--gpus", "default"),
           arg_str(argc, argv, "--tile-size", "default"), arg_str(argc, argv, "--panel-width", "default"),
           h_form ? " hessenberg:normal" : "", h_res ? " residual:normal" : "", h_print ? " print:normal" : "",
           h_store ? " store-raw:normal" : "", fail_thr, warn_thr, repeat, warmup);

    /* ---- INIT: the pencil (A, Q = I) and a pristine copy for the residual check */
    printf("INIT...\n");
    int ld;
    double *A0 = alloc_matrix(n, &ld), *A = alloc_matrix(n, &ld), *Q = alloc_matrix(n, &ld);
    lcg_state = seed;
    if (strcmp(init, "read-raw") == 0) {
        format_name(name, sizeof(name), input, "A");
        if (read_raw(name, n, A0, ld) != 0) { fprintf(stderr, "Invalid file.\n"); return 2; }
    } else if (partial) {
        fill_partial(n, begin, end, A0, ld);
    } else if (strcmp(init, "full") == 0) {
        fill_random(n, A0, ld, 0);
    } else if (strcmp(init, "default") == 0 || strcmp(init, "fullpos") == 0) {
        fill_random(n, A0, ld, 1);
    } else {
        fprintf(stderr, "Invalid initializer.\n");
        return 2;
    }

    double *times = (double *)calloc((size_t)repeat, sizeof(double));
    double *res_a = (double *)calloc((size_t)repeat, sizeof(double)), *res_q = (double *)calloc((size_t)repeat, sizeof(double));
    int form_fails = 0, res_fails = 0, res_warns = 0, solver_error = 0;

    for (int iter = -warmup; iter < repeat && !solver_error; iter++) {
        printf("PREPARE...\n");
        copy_matrix(n, A0, ld, A, ld);
        set_identity(n, Q, ld);
        if (!use_lapack) starneig_node_init(cores, gpus, STARNEIG_HINT_SM | STARNEIG_AWAKE_WORKERS);

        printf("PROCESS...\n");
        fflush(stdout);
        double t0 = now_ms();
        starneig_error_t ret;
        if (use_lapack) {
            ret = lapack_solver(n, begin, end, A, ld, Q, ld);
        } else if (expert) {
            struct starneig_hessenberg_conf conf;
            starneig_hessenberg_init_conf(&conf);
            conf.tile_size = arg_int(argc, argv, "--tile-size", STARNEIG_HESSENBERG_DEFAULT_TILE_SIZE);
            conf.panel_width = arg_int(argc, argv, "--panel-width", STARNEIG_HESSENBERG_DEFAULT_PANEL_WIDTH);
            ret = starneig_SEP_SM_Hessenberg_expert(&conf, n, begin, end, A, ld, Q, ld);
        } else {
            ret = starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld);
        }
        double dt = now_ms() - t0;
        printf(iter < 0 ? "WARMUP TIME = %.0f MS\n" : "EXPERIMENT TIME = %.0f MS\n", dt);

        printf("FINALIZE...\n");
        if (!use_lapack) starneig_node_finalize();
        if (ret != STARNEIG_SUCCESS) {
            fprintf(stderr, "The solver returned %d.\n", ret);
            solver_error = 1;
            break;
        }
        if (iter < 0) continue;
        times[iter] = dt;

        if (h_form && form_violations(n, A, ld, begin, end, partial) > 0) form_fails++;
        if (h_res) {
            printf("|Q ~A Q^T - A| / |A|"); fflush(stdout);
            res_a[iter] = residual_u(n, Q, ld, A, ld, A0, ld);
            printf(" = %.0f u\n", res_a[iter]);
            printf("|Q Q^T - I| / |I|"); fflush(stdout);
            res_q[iter] = orthogonality_u(n, Q, ld);
            printf(" = %.0f u\n", res_q[iter]);
            int warn = (warn_thr < res_a[iter]) + (warn_thr < res_q[iter]);
            int fail = (fail_thr < res_a[iter] || isnan(res_a[iter])) + (fail_thr < res_q[iter] || isnan(res_q[iter]));
            if (fail) res_fails++; else if (warn) res_warns++;
        }
        if (h_print && n <= 20) {
            for (int r = 0; r < n; r++) {
                for (int c = 0; c < n; c++) printf(" %10.3e", A[(size_t)c * ld + r]);
                printf("\n");
            }
        }
        if (h_store) {
            format_name(name, sizeof(name), store_fmt, "A");  write_raw(name, n, A, ld);
            format_name(name, sizeof(name), store_fmt, "Q");  write_raw(name, n, Q, ld);
            format_name(name, sizeof(name), store_fmt, "CA"); write_raw(name, n, A0, ld);
        }
    }

    int status = solver_error ? 2 : 0;
    if (!solver_error) {
        printf("================================================================\n");
        double *sorted = (double *)malloc((size_t)repeat * sizeof(double));
        memcpy(sorted, times, (size_t)repeat * sizeof(double));
        qsort(sorted, (size_t)repeat, sizeof(double), cmp_double);
        double mean = 0.0, var = 0.0;
        for (int i = 0; i < repeat; i++) mean += sorted[i] / repeat;
        for (int i = 0; i < repeat; i++) var += (sorted[i] - mean) * (sorted[i] - mean) / repeat;
        double median = repeat % 2 ? sorted[repeat / 2] : 0.5 * (sorted[repeat / 2 - 1] + sorted[repeat / 2]);
        printf("TIME = %.0f MS [avg %.0f MS, cv %.2f, min %.0f MS, max %.0f MS]\n", median, mean,
               var > 0.0 ? sqrt(var) / mean : 0.0, sorted[0], sorted[repeat - 1]);
        printf("GFLOPS = %.1f (10 n^3 / 3 over the median time)\n", 10.0 / 3.0 * (double)n * n * n / (median * 1e-3) / 1e9);
        free(sorted);
        if (h_form) {
            if (form_fails == 0) printf("NO FAILED HESSENBERG FORM TESTS\n");
            else printf("%d HESSENBERG FORM TESTS FAILED\n", form_fails);
        }
        if (h_res) {
            print_stats("|Q ~A Q^T - A| / |A|", " u", repeat, res_a);
            print_stats("|Q Q^T - I| / |I|", " u", repeat, res_q);
            if (res_warns) printf("RESIDUAL CHECK (WARNINGS): %d runs effected\n", res_warns);
            if (res_fails) printf("RESIDUAL CHECK (FAILS): %d runs effected\n", res_fails);
        }
        if (form_fails || res_fails) status = 1;
    }
    free(times); free(res_a); free(res_q);
    free_matrix(A0); free_matrix(A); free_matrix(Q);
    return status;
}
