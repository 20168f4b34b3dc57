/* CBLAS prototypes (no cblas.h on this image); symbols come from the OpenBLAS bundled with
 * scipy, whose exports carry a scipy_ prefix. Test infrastructure only. */
#ifndef ORACLE_CBLAS_SHIM_H
#define ORACLE_CBLAS_SHIM_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
#ifdef ORACLE_BLAS_PREFIX_SCIPY
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_dtrmv scipy_cblas_dtrmv
#define cblas_dtrmm scipy_cblas_dtrmm
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_dscal scipy_cblas_dscal
#define cblas_dcopy scipy_cblas_dcopy
#define cblas_ddot  scipy_cblas_ddot
#define cblas_dnrm2 scipy_cblas_dnrm2
#define dlarfg_ scipy_dlarfg_
#define dgehrd_ scipy_dgehrd_
#define dormhr_ scipy_dormhr_
#define dhseqr_ scipy_dhseqr_
#define dlange_ scipy_dlange_
#define openblas_set_num_threads scipy_openblas_set_num_threads
#define openblas_get_num_threads scipy_openblas_get_num_threads
#endif
void cblas_dgemv(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int m, int n, double alpha, const double *A, int lda,
    const double *x, int incx, double beta, double *y, int incy);
void cblas_dgemm(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, enum CBLAS_TRANSPOSE, int m, int n, int k, double alpha,
    const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
void cblas_dtrmv(enum CBLAS_ORDER, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, enum CBLAS_DIAG, int n,
    const double *A, int lda, double *x, int incx);
void cblas_dtrmm(enum CBLAS_ORDER, enum CBLAS_SIDE, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, enum CBLAS_DIAG,
    int m, int n, double alpha, const double *A, int lda, double *B, int ldb);
void cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
void cblas_dscal(int n, double alpha, double *x, int incx);
void cblas_dcopy(int n, const double *x, int incx, double *y, int incy);
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
double cblas_dnrm2(int n, const double *x, int incx);
void openblas_set_num_threads(int);
int openblas_get_num_threads(void);
#endif
