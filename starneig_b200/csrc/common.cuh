// common.cuh -- shared helpers for the sm_100a Hessenberg kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

namespace sb200 {

// Unrecoverable faults follow the reference's convention (src/common/common.c:143-152):
// print "[starneig][fatal error] ..." and exit(EXIT_FAILURE).
[[noreturn]] inline void fatal(const char *what, const char *file, int line)
{
    fprintf(stderr, "[starneig][fatal error] %s (%s:%d)\n", what, file, line);
    fflush(stderr);
    exit(EXIT_FAILURE);
}

#define SB_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t err__ = (expr);                                          \
        if (err__ != cudaSuccess)                                            \
            ::sb200::fatal(cudaGetErrorString(err__), __FILE__, __LINE__);   \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// deterministic butterfly sum: every lane ends with the same value
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// "last block done" election (CUDA threadFenceReduction pattern). Every thread of the block must have
// issued its global partial-result stores before the call. Returns true in exactly one block of the
// grid: the one that arrives last; that block may then read all partials with __ldcg().
__device__ __forceinline__ bool last_block_done(unsigned *counter, unsigned total_blocks)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == total_blocks - 1);
        if (is_last) *counter = 0;      // re-arm for the next launch (stream ordered)
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

} // namespace sb200
