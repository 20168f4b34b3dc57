/* <starneig/sep_sm.h> -- standard eigenvalue problem, shared memory, Hessenberg stage only.
 * Drop-in for reference src/include/starneig/sep_sm.h:60-92,344-384. */
#ifndef STARNEIG_SEP_SM_H
#define STARNEIG_SEP_SM_H

#include <starneig/configuration.h>
#include <starneig/error.h>
#include <starneig/expert.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sep_sm.h:89-92 / src/hessenberg/interface.c:170-185.
 * A (n x n, column-major, ldA >= n) is overwritten by the upper Hessenberg H with exact zeros below
 * the sub-diagonal; Q (n x n, orthogonal on entry, usually I) is overwritten by Q*U where
 * A_in = U H U^T. Host pointers; blocking; in place.
 * Returns 0, -1 (n<1), -2 (A NULL), -3 (ldA<n), -4 (Q NULL), -5 (ldQ<n), STARNEIG_NOT_INITIALIZED. */
starneig_error_t starneig_SEP_SM_Hessenberg(
    int n, double A[], int ldA, double Q[], int ldQ);

/* sep_sm.h:380-384 / src/hessenberg/interface.c:138-167. Reduces columns begin .. end-2 only.
 * Returns 0, -2 (n<1), -3 (begin<0), -4 (n<end), -5 (A NULL), -6 (ldA<n), -7 (Q NULL), -8 (ldQ<n),
 * STARNEIG_NOT_INITIALIZED, STARNEIG_INVALID_CONFIGURATION (tile_size or panel_width in [0,8)). */
starneig_error_t starneig_SEP_SM_Hessenberg_expert(
    struct starneig_hessenberg_conf *conf, int n, int begin, int end,
    double A[], int ldA, double Q[], int ldQ);

#ifdef __cplusplus
}
#endif

#endif
