/* <starneig/error.h> -- error codes. Same values as reference src/include/starneig/error.h:66-127.
 * Convention (reference src/hessenberg/interface.c:144-150,175-179): 0 = success, a negative value -i
 * means the i-th argument was invalid, a positive value is one of the library codes below. */
#ifndef STARNEIG_ERROR_H
#define STARNEIG_ERROR_H

typedef int starneig_error_t;

#define STARNEIG_SUCCESS                0
#define STARNEIG_GENERIC_ERROR          1
#define STARNEIG_NOT_INITIALIZED        2
#define STARNEIG_INVALID_CONFIGURATION  3
#define STARNEIG_INVALID_ARGUMENTS      4
#define STARNEIG_INVALID_DISTR_MATRIX   5
#define STARNEIG_DID_NOT_CONVERGE       6
#define STARNEIG_PARTIAL_REORDERING     7
#define STARNEIG_CLOSE_EIGENVALUES      8

#endif
