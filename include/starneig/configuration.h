/* <starneig/configuration.h> -- build configuration of the B200-native Hessenberg library.
 * Replaces the CMake-generated header (reference: src/include/starneig/configuration.h.in). */
#ifndef STARNEIG_CONFIGURATION_H
#define STARNEIG_CONFIGURATION_H

#define STARNEIG_VERSION_MAJOR 0
#define STARNEIG_VERSION_MINOR 2
#define STARNEIG_VERSION_PATCH 0

/* The hot path is CUDA-only (sm_100a); there is no CPU fallback. */
#define STARNEIG_ENABLE_CUDA

/* marks this build for callers that want to detect it */
#define STARNEIG_B200_NATIVE 1

#endif
