// cusim_device.h -- device-side half of the kernel-logic emulator (see cuda_runtime.h in this directory).
// TEST INFRASTRUCTURE, not part of the product. Included by starneig_b200/csrc/common.cuh when SB_CUSIM is defined;
// it provides, on top of fibers (one per CUDA thread, cusim.cpp), exactly the names common.cuh otherwise implements
// with inline PTX, plus the CUDA built-ins the kernels use:
//   threadIdx / blockIdx / blockDim / gridDim, __syncthreads, named barriers, warp shuffles and votes (every
//   participating lane must arrive, as on the hardware), the m8n8k4 FP64 MMA with the hardware's fragment layout,
//   acquire/release and volatile accesses (each polling access yields to the other fibers, other blocks and other
//   ranks), cp.async as an immediate copy with zero-fill, atomics, clocks.
// Dynamic shared memory and cudaMalloc memory start out as NaN patterns, so a kernel that relies on zero-initialised
// memory fails here even though the arithmetic is otherwise identical to the GPU's (same fma/sqrt/hypot in IEEE double).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <tuple>
#include <utility>

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __shared__ static thread_local

namespace cusim {
void cta_barrier();
void named_barrier(int id, int nthreads);
void poll_yield();
// all-to-all of up to 16 bytes per lane inside the warp; blocks until all live lanes of the warp arrived
void warp_exchange(const void *mine, size_t bytes, void *all /* 32 x 16 bytes */);
unsigned long long now_ns();

template <class K, class... Args>
static inline void launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t, Args... args)
{
    run_grid(grid, block, smem, false, [=]() { kernel(args...); });
}
template <class K, class Arg>
static inline void launch_coop(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t, const Arg &arg)
{
    const Arg copy = arg;
    run_grid(grid, block, smem, true, [=]() { kernel(copy); });
}
} // namespace cusim

#define threadIdx (::cusim::g_thread->t_idx)
#define blockIdx (::cusim::g_thread->b_idx)
#define blockDim (::cusim::g_thread->b_dim)
#define gridDim (::cusim::g_thread->g_dim)

#define SB_DYNAMIC_SMEM(type, name) type *const name = (type *)::cusim::g_thread->smem
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) ::cusim::launch(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#define SB_LAUNCH_COOP(kernel, grid, block, smem, stream, arg) ::cusim::launch_coop(kernel, dim3(grid), dim3(block), (smem), (stream), (arg))

static inline void __syncthreads() { ::cusim::cta_barrier(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { ::cusim::poll_yield(); }
static inline long long clock64() { return (long long)::cusim::now_ns(); }      // "cycles" = ns: time-outs stay in seconds
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }

// cache-hinted loads: a real (re)load from memory, never a value the compiler kept in a register
template <class T> static inline T __ldcg(const T *p)
{
    T v;
    asm volatile("" ::: "memory");
    memcpy(&v, (const void *)p, sizeof(T));
    asm volatile("" ::: "memory");
    return v;
}
template <class T> static inline T __ldcs(const T *p) { return __ldcg(p); }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }

template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
    static_assert(sizeof(T) <= 16, "shuffle payload");
    alignas(16) unsigned char all[32][16];
    ::cusim::warp_exchange(&v, sizeof(T), all);
    T r;
    memcpy(&r, all[(::cusim::g_thread->lane ^ (unsigned)lane_mask) & 31], sizeof(T));
    return r;
}
static inline void __syncwarp()
{
    alignas(16) unsigned char all[32][16];
    const int mine = 0;
    ::cusim::warp_exchange(&mine, sizeof(int), all);
}
static inline int __any_sync(unsigned, int pred)
{
    alignas(16) unsigned char all[32][16];
    const int mine = pred ? 1 : 0;
    ::cusim::warp_exchange(&mine, sizeof(int), all);
    int any = 0;
    const unsigned nl = std::min(32u, ::cusim::g_thread->b_dim.x * ::cusim::g_thread->b_dim.y * ::cusim::g_thread->b_dim.z
                                          - 32u * ::cusim::g_thread->warp);
    for (unsigned l = 0; l < nl; l++) { int v; memcpy(&v, all[l], sizeof(int)); any |= v; }
    return any;
}

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// CUDA's integer min/max overloads (the kernels mix int and long long)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
using std::fma; using std::sqrt; using std::fabs; using std::hypot; using std::copysign; using std::fmin; using std::fmax;

namespace cusim { extern unsigned long long g_keep_loads, g_prefetches; }     // CUSIM_STATS=1 prints them at exit

namespace sb200 {
// ---- the primitives of common.cuh ----------------------------------------------------------------------------
static inline unsigned ld_acquire_gpu(const unsigned *p) { ::cusim::poll_yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline unsigned ld_acquire_sys(const unsigned *p) { ::cusim::poll_yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void red_release_gpu_add(unsigned *p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) { ::cusim::poll_yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void red_release_gpu_add_u64(unsigned long long *p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
static inline void st_release_sys(unsigned *p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
// 16-byte entries: two 8-byte words, each accessed atomically (what the LL protocol assumes of the interconnect)
static inline uint4 ld_volatile_v4(const uint4 *p)
{
    ::cusim::poll_yield();
    const unsigned long long lo = __atomic_load_n((const unsigned long long *)p, __ATOMIC_RELAXED);
    const unsigned long long hi = __atomic_load_n((const unsigned long long *)p + 1, __ATOMIC_RELAXED);
    return make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
}
static inline void st_volatile_v4(uint4 *p, uint4 e)
{
    __atomic_store_n((unsigned long long *)p, (unsigned long long)e.x | ((unsigned long long)e.y << 32), __ATOMIC_RELAXED);
    __atomic_store_n((unsigned long long *)p + 1, (unsigned long long)e.z | ((unsigned long long)e.w << 32), __ATOMIC_RELAXED);
}
static inline void group_barrier(int id, int nthreads) { ::cusim::named_barrier(id, nthreads); }
static inline unsigned long long globaltimer_ns() { return ::cusim::now_ns(); }
static inline unsigned long long l2_policy_evict_last() { return 0ull; }
static inline double2 ld_l2_keep(const double2 *p, unsigned long long) { __atomic_fetch_add(&::cusim::g_keep_loads, 1ull, __ATOMIC_RELAXED); return __ldcg(p); }
// a hint on the device; here (CUSIM_CHECK_PREFETCH=1) the address must at least lie inside a device allocation
static inline void prefetch_l2(const void *p)
{
    static const bool check = getenv("CUSIM_CHECK_PREFETCH") != nullptr;
    __atomic_fetch_add(&::cusim::g_prefetches, 1ull, __ATOMIC_RELAXED);
    if (check && !::cusim::inside_device_allocation(p, 8)) { fprintf(stderr, "cusim: prefetch outside every device allocation\n"); abort(); }
}
static inline void prefetch_l2_bulk(const void *p, unsigned bytes)
{
    static const bool check = getenv("CUSIM_CHECK_PREFETCH") != nullptr;
    __atomic_fetch_add(&::cusim::g_prefetches, 1ull, __ATOMIC_RELAXED);
    if ((((uintptr_t)p) & 15) || (bytes & 15) || bytes == 0) { fprintf(stderr, "cusim: misaligned bulk prefetch\n"); abort(); }
    if (check && !::cusim::inside_device_allocation(p, bytes)) { fprintf(stderr, "cusim: bulk prefetch outside every device allocation\n"); abort(); }
}
static inline void cp_async8(void *smem_dst, const void *gmem_src, bool valid)
{
    if (valid) memcpy(smem_dst, gmem_src, 8);
    else memset(smem_dst, 0, 8);
}
static inline void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes)
{
    if ((((uintptr_t)smem_dst) | ((uintptr_t)gmem_src)) & 15) { fprintf(stderr, "cusim: misaligned 16-byte cp.async\n"); abort(); }
    memset(smem_dst, 0, 16);
    if (src_bytes > 0) memcpy(smem_dst, gmem_src, (size_t)src_bytes);
}
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
// ---- TMA + mbarrier (dgemm_tma.cuh). The tensor map is a plain struct, the bulk tensor copy is performed at once by the
// issuing thread (zero fill outside the tensor, 128-byte swizzle relative to the box start -- on the device the box starts
// on a 1024-byte boundary, where the address-based pattern of the hardware is the same), and an mbarrier is
// {phase : 1, pending arrivals : 15, arrival count : 16, pending bytes : 32} in its 8 bytes. The fibers of a CTA run on
// one host thread, so no atomics are needed.
struct alignas(64) SbTensorMap { const double *base; unsigned long long dim0, dim1, ld; unsigned box0, box1; int swizzle; char pad[76]; };
#define SB_GRID_CONSTANT
static inline void sb_make_tensor_map(SbTensorMap *map, const double *base, unsigned long long dim0, unsigned long long dim1,
                                      unsigned long long ld, unsigned box0, unsigned box1)
{
    if (((uintptr_t)base & 15) != 0 || (ld & 1) != 0 || box0 * sizeof(double) != 128 || box1 > 256 || dim0 == 0 || dim1 == 0) {
        fprintf(stderr, "cusim: invalid tensor map (base %p ld %llu box %u x %u)\n", (const void *)base, ld, box0, box1);
        abort();
    }
    map->base = base; map->dim0 = dim0; map->dim1 = dim1; map->ld = ld; map->box0 = box0; map->box1 = box1; map->swizzle = 1;
}
static inline double *sb_align_shared(double *p, unsigned align) { return (double *)(((uintptr_t)p + align - 1) / align * align); }
namespace mbar_bits {
static inline unsigned phase(unsigned long long b) { return (unsigned)(b >> 63); }
static inline unsigned pending(unsigned long long b) { return (unsigned)((b >> 48) & 0x7fff); }
static inline unsigned count(unsigned long long b) { return (unsigned)((b >> 32) & 0xffff); }
static inline int tx(unsigned long long b) { return (int)(unsigned)(b & 0xffffffffull); }
static inline unsigned long long pack(unsigned ph, unsigned pend, unsigned cnt, int t)
{
    return ((unsigned long long)ph << 63) | ((unsigned long long)pend << 48) | ((unsigned long long)cnt << 32) | (unsigned)t;
}
static inline void settle(unsigned long long *bar)      // phase completes when no arrival and no byte is pending
{
    const unsigned long long b = *bar;
    if (pending(b) == 0 && tx(b) == 0) *bar = pack(phase(b) ^ 1u, count(b), count(b), 0);
}
}
static inline void mbar_init(unsigned long long *bar, unsigned count) { *bar = mbar_bits::pack(0, count, count, 0); }
static inline void mbar_fence_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    const unsigned long long b = *bar;
    if (mbar_bits::pending(b) == 0) { fprintf(stderr, "cusim: mbarrier arrival beyond its count\n"); abort(); }
    *bar = mbar_bits::pack(mbar_bits::phase(b), mbar_bits::pending(b) - 1, mbar_bits::count(b), mbar_bits::tx(b) + (int)bytes);
    mbar_bits::settle(bar);
}
static inline void mbar_arrive(unsigned long long *bar) { mbar_arrive_expect_tx(bar, 0u); }
static inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    while (mbar_bits::phase(*(volatile unsigned long long *)bar) == parity) ::cusim::poll_yield();
}
static inline void tma_load_2d(void *smem_dst, const SbTensorMap *map, int c0, int c1, unsigned long long *bar)
{
    double *dst = (double *)smem_dst;
    // the hardware addresses global memory in 16-byte units: an odd coordinate in the contiguous dimension traps
    if (((uintptr_t)smem_dst & 127) != 0) { fprintf(stderr, "cusim: TMA destination is not 128-byte aligned\n"); abort(); }
    if (c0 & 1) { fprintf(stderr, "cusim: TMA box starts at an odd element of the contiguous dimension (c0 = %d)\n", c0); abort(); }
    for (unsigned r = 0; r < map->box1; r++)
        for (unsigned e = 0; e < map->box0; e++) {
            const long long x0 = (long long)c0 + e, x1 = (long long)c1 + r;
            const bool inside = x0 >= 0 && x1 >= 0 && (unsigned long long)x0 < map->dim0 && (unsigned long long)x1 < map->dim1;
            const size_t off = map->swizzle ? r * 16 + ((((e >> 1) ^ (r & 7)) << 1) | (e & 1)) : (size_t)r * map->box0 + e;
            dst[off] = inside ? map->base[(size_t)x1 * map->ld + x0] : 0.0;
        }
    const unsigned long long b = *bar;
    *bar = mbar_bits::pack(mbar_bits::phase(b), mbar_bits::pending(b), mbar_bits::count(b), mbar_bits::tx(b) - (int)(map->box0 * map->box1 * sizeof(double)));
    mbar_bits::settle(bar);
    ::cusim::poll_yield();
}

// mma.sync.m8n8k4.f64: lane l holds A[l/4][l%4], B[l%4][l/4] and C[l/4][2*(l%4) + {0,1}]; the products of one
// output element are accumulated in ascending k with fused multiply-adds
static inline void dmma884(double &c0, double &c1, double a, double b)
{
    alignas(16) double all[32][2];
    const double mine[2] = {a, b};
    ::cusim::warp_exchange(mine, 16, all);
    const unsigned lane = ::cusim::g_thread->lane, g = lane >> 2, t = lane & 3;
    for (int k = 0; k < 4; k++) {
        const double av = all[g * 4 + k][0];
        c0 = fma(av, all[(2 * t) * 4 + k][1], c0);
        c1 = fma(av, all[(2 * t + 1) * 4 + k][1], c1);
    }
}
} // namespace sb200
