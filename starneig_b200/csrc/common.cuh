// common.cuh -- shared helpers for the sm_100a Hessenberg kernels.
#pragma once
#include <cuda_runtime.h>
#ifndef SB_CUSIM
#include <cuda.h>               // CUtensorMap (types only: the encoder is resolved through cudaGetDriverEntryPoint)
#endif
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#ifdef SB_CUSIM
#include "cusim_device.h"       // tests/cusim/include: the emulator's versions of the hardware primitives below
#endif

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

// ---------------------------------------------------------------------------------------------
// Hardware primitives. Everything that is inline PTX or launch syntax lives here, in one place. With
// -DSB_CUSIM (tests/cusim: the kernel-logic emulator of the test suite, never part of the product) the same
// names are provided by cusim_device.h on top of host threads/fibers, so that the kernels and the engine can
// be compiled unchanged by g++ and stepped through on a machine without a GPU.
// ---------------------------------------------------------------------------------------------
#ifndef SB_CUSIM
#define SB_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// cooperative launch (all CTAs co-resident) of a kernel that takes one argument
#define SB_LAUNCH_COOP(kernel, grid, block, smem, stream, arg)                                           \
    do {                                                                                                 \
        void *args__[] = {(void *)&(arg)};                                                               \
        SB_CUDA(cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(block), args__, (smem), (stream))); \
    } while (0)

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 64-bit variants: an arrival counter (low word) that carries grid-wide vote counts in its high word, so that the
// thread that polls the barrier learns the outcome of the vote with the same load
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 16-byte volatile accesses (one transaction each): the carriers of the "LL" exchange entries
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4 *p)
{
    uint4 e;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "l"(p) : "memory");
    return e;
}
__device__ __forceinline__ void st_volatile_v4(uint4 *p, uint4 e)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(e.x), "r"(e.y), "r"(e.z), "r"(e.w) : "memory");
}
// named barrier `id` (1..15) over `nthreads` threads of the CTA
__device__ __forceinline__ void group_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// bulk variant (TMA engine): `bytes` (multiple of 16) from a 16-byte aligned address with ONE instruction
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// 8-byte asynchronous global -> shared copy; !valid: the 8 destination bytes are zero-filled (src-size 0)
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src, bool valid)
{
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int src_size = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_size));
}
// 16-byte variant (both addresses 16-byte aligned): copies the first src_bytes (0, 8 or 16) bytes, zero-fills the rest
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes)
{
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
// FP64 tensor-core instruction (SASS DMMA.8x8x4): C(8x8) += A(8x4) B(4x8); lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2*(l%4) + {0,1}]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// ---- TMA + mbarrier (dgemm_tma.cuh). A tensor map is a 128-byte descriptor made on the host and passed to the kernel as a
// __grid_constant__ parameter; one thread issues a bulk tensor copy global -> shared memory that signals an mbarrier with
// the number of bytes it delivered.
typedef CUtensorMap SbTensorMap;
#define SB_GRID_CONSTANT __grid_constant__
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// first address >= p that is `align`-byte aligned in the shared window (TMA swizzle patterns are functions of the address)
__device__ __forceinline__ double *sb_align_shared(double *p, unsigned align)
{
    const unsigned a = smem_u32(p);
    return (double *)((char *)p + ((align - (a & (align - 1))) & (align - 1)));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the other threads' and the TMA engine's view of shared memory
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// blocks until the phase of the barrier with the given parity has completed
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "SB_MBAR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra SB_MBAR_DONE;\n"
                 "bra SB_MBAR_WAIT;\n"
                 "SB_MBAR_DONE:\n"
                 "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before its later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// box of the tensor map at coordinates (c0 = contiguous dimension, c1) -> shared memory; completes `bar` by the box's bytes
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const SbTensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
#endif

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
