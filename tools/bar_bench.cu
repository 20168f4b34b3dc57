// Micro-benchmark of device-side grid barriers on a co-resident grid (one CTA of 512 threads per SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/bar_bench tools/bar_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rlx(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// 0: one atomic counter, thread 0 of every CTA polls it
__device__ void bar0(unsigned *f, unsigned &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += gridDim.x;
        __threadfence();
        atomicAdd(f, 1u);
        while ((int)(ld_acq(f) - gen) < 0) { }
    }
    __syncthreads();
}
// 1: slots + master gather + go word
__device__ void bar1(unsigned *f, unsigned &gen)
{
    __syncthreads();
    gen++;
    if (blockIdx.x == 0) {
        if (threadIdx.x > 0 && threadIdx.x < gridDim.x) while ((int)(ld_acq(f + threadIdx.x) - gen) < 0) { }
        __syncthreads();
        if (threadIdx.x == 0) st_rel(f + 512, gen);
    } else {
        if (threadIdx.x == 0) { st_rel(f + blockIdx.x, gen); while ((int)(ld_acq(f + 512) - gen) < 0) { } }
        __syncthreads();
    }
}
// 2: two-level: groups of GS CTAs; group leader gathers its members, publishes a group flag; everybody polls the group flags
template <int GS>
__device__ void bar2(unsigned *f, unsigned &gen)
{
    __syncthreads();
    gen++;
    const int b = blockIdx.x, G = gridDim.x, grp = b / GS, ngrp = (G + GS - 1) / GS;
    if (b % GS == 0) {
        const int members = min(GS, G - grp * GS);
        if (threadIdx.x > 0 && threadIdx.x < members) while ((int)(ld_acq(f + b + threadIdx.x) - gen) < 0) { }
        __syncthreads();
        if (threadIdx.x == 0) st_rel(f + 512 + 32 * grp, gen);
    } else {
        if (threadIdx.x == 0) st_rel(f + b, gen);
    }
    if (threadIdx.x < ngrp) while ((int)(ld_acq(f + 512 + 32 * threadIdx.x) - gen) < 0) { }
    __syncthreads();
}
// 3: atomic counters per group (GS CTAs each) + top counter
template <int GS>
__device__ void bar3(unsigned *f, unsigned &gen)
{
    __syncthreads();
    gen++;
    const int b = blockIdx.x, G = gridDim.x, grp = b / GS, ngrp = (G + GS - 1) / GS;
    if (threadIdx.x == 0) {
        const int members = min(GS, G - grp * GS);
        __threadfence();
        unsigned t = atomicAdd(f + 32 * (1 + grp), 1u);
        if (t == gen * members - 1) { __threadfence(); atomicAdd(f, 1u); }
        while ((int)(ld_acq(f) - gen * ngrp) < 0) { }
    }
    __syncthreads();
}
// 4: like 1 but relaxed polling + fences
__device__ void bar4(unsigned *f, unsigned &gen)
{
    __syncthreads();
    gen++;
    if (blockIdx.x == 0) {
        if (threadIdx.x > 0 && threadIdx.x < gridDim.x) { while ((int)(ld_rlx(f + threadIdx.x) - gen) < 0) { } }
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); st_rlx(f + 512, gen); }
    } else {
        if (threadIdx.x == 0) { __threadfence(); st_rlx(f + blockIdx.x, gen); while ((int)(ld_rlx(f + 512) - gen) < 0) { } __threadfence(); }
        __syncthreads();
    }
}

// 7: atomic counter, red.release instead of fence + atom
__device__ void bar7(unsigned *f, unsigned &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
        while ((int)(ld_acq(f) - gen) < 0) { }
    }
    __syncthreads();
}
// 8: atomic counter, relaxed polls + fence
__device__ void bar8(unsigned *f, unsigned &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += gridDim.x;
        __threadfence();
        atomicAdd(f, 1u);
        while ((int)(ld_rlx(f) - gen) < 0) { }
        __threadfence();
    }
    __syncthreads();
}
// 9: atomic counter, the last arriver (sees it from the return value) releases a go word
__device__ void bar9(unsigned *f, unsigned &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += gridDim.x;
        __threadfence();
        const unsigned t = atomicAdd(f, 1u);
        if (t == gen - 1) st_rel(f + 64, gen);
        else while ((int)(ld_acq(f + 64) - gen) < 0) { }
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(unsigned *f, int iters, double *data, int work)
{
    unsigned gen = 0;
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        // a little global traffic between barriers, as in the panel kernel
        for (int q = 0; q < work; q++) { data[(size_t)blockIdx.x * 8192 + q * 512 + threadIdx.x] = acc + it; }
        if (MODE == 0) bar0(f, gen);
        if (MODE == 1) bar1(f, gen);
        if (MODE == 2) bar2<12>(f, gen);
        if (MODE == 3) bar3<12>(f, gen);
        if (MODE == 4) bar4(f, gen);
        if (MODE == 5) bar2<8>(f, gen);
        if (MODE == 6) bar2<16>(f, gen);
        if (MODE == 7) bar7(f, gen);
        if (MODE == 8) bar8(f, gen);
        if (MODE == 9) bar9(f, gen);
        acc += __ldcg(data + (size_t)((blockIdx.x + 1) % gridDim.x) * 8192 + threadIdx.x);
    }
    if (acc == 1.2345) data[0] = acc;
}

template <int MODE> void run(const char *name, unsigned *f, double *data, int G, int work)
{
    const int iters = 20000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    void *args[] = {&f, (void *)&iters, &data, &work};
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaMemset(f, 0, 4096 * 4));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void *)k<MODE>, dim3(G), dim3(512), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-44s work=%d : %7.3f us per iteration\n", name, work, 1e3 * ms / iters);
}

int main()
{
    int G; CK(cudaDeviceGetAttribute(&G, cudaDevAttrMultiProcessorCount, 0));
    unsigned *f; double *data;
    CK(cudaMalloc(&f, 4096 * 4)); CK(cudaMalloc(&data, (size_t)G * 8192 * 8 + 8192 * 8));
    CK(cudaMemset(data, 0, (size_t)G * 8192 * 8));
    for (int work = 0; work <= 4; work += 4) {
        run<0>("0 atomic counter", f, data, G, work);
        run<1>("1 slots, master gather, go word (acq/rel)", f, data, G, work);
        run<4>("4 same, relaxed polls + fences", f, data, G, work);
        run<2>("2 two-level slots, groups of 12", f, data, G, work);
        run<5>("5 two-level slots, groups of 8", f, data, G, work);
        run<6>("6 two-level slots, groups of 16", f, data, G, work);
        run<3>("3 two-level atomics, groups of 12", f, data, G, work);
        run<7>("7 atomic counter, red.release", f, data, G, work);
        run<8>("8 atomic counter, relaxed polls + fence", f, data, G, work);
        run<9>("9 atomic counter, last arriver sets go word", f, data, G, work);
    }
    return 0;
}
