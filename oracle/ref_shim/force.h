/* force-included when compiling reference sources: route Fortran BLAS symbols to the scipy OpenBLAS */
#define dgemm_ scipy_dgemm_
#define dgehrd_ scipy_dgehrd_
#define dormhr_ scipy_dormhr_
#define dlarfg_ scipy_dlarfg_
#define dlamch_ scipy_dlamch_
#define dlag2_ scipy_dlag2_
#define dlanv2_ scipy_dlanv2_
