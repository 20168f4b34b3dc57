// Development experiment (not part of the product): what happens to an HBM-streaming kernel on G SMs when another
// kernel runs on the remaining SMs? The stream kernel mimics the GEMV phase of k_panel_fused (640 threads, one
// SM-exclusive CTA per SM, 16 warps x 8 x 16-byte loads in flight); the side kernel is one of
//   dmma   : register-only DMMA loop (FP64 tensor pipe, no memory traffic)
//   l2     : reads a small (L2-resident) buffer over and over
//   dram   : streams a large buffer (extra HBM traffic)
//   dfma   : register-only DFMA loop
// Both are SM-exclusive (register-file sized). Prints the stream kernel's bandwidth alone and with each side load,
// and the SM clock sampled through NVML-free means (clock64 vs globaltimer).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/overlap_probe tools/overlap_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(640, 1) k_stream(const double2 *__restrict__ A, size_t n2, double *out, unsigned long long *clk)
{
    // 512 threads stream, 128 idle (like the look-ahead warps); grid-stride over 16-byte elements
    const int tid = threadIdx.x;
    unsigned long long c0 = clock64(), t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    double2 acc = make_double2(0, 0);
    if (tid < 512) {
        const size_t stride = (size_t)gridDim.x * 512;
        size_t i = (size_t)blockIdx.x * 512 + tid;
        for (; i + 7 * stride < n2; i += 8 * stride) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldcs(A + i + u * stride);
#pragma unroll
            for (int u = 0; u < 8; u++) { acc.x += v[u].x; acc.y += v[u].y; }
        }
    }
    if (acc.x == 1.2345) out[0] = acc.y;
    if (tid == 0 && blockIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        clk[0] = clock64() - c0; clk[1] = t1 - t0;
    }
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// SM-exclusive side kernels: 256 threads, 200 registers forced through a big accumulator array
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_side(const double *__restrict__ buf, size_t n, double *out, volatile int *stop, long long iters)
{
    double acc[80];
#pragma unroll
    for (int i = 0; i < 80; i++) acc[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (long long it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 80; i += 2) dmma884(acc[i], acc[i + 1], a, b);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 80; i++) acc[i] = fma(acc[i], a, b);
        } else {
            // MODE 1: L2-resident buffer (n small); MODE 2: large buffer
            size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x + (size_t)it * gridDim.x * 256 * 16) % (n - 16 * (size_t)gridDim.x * 256);
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] += __ldcg(buf + base + (size_t)i * gridDim.x * 256);
        }
        if ((it & 63) == 0 && *stop) break;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 80; i++) s += acc[i];
    if (s == 1.2345) out[1] = s;
}

int main(int argc, char **argv)
{
    const int G = argc > 1 ? atoi(argv[1]) : 116;
    const int K = argc > 2 ? atoi(argv[2]) : 148 - G;
    const size_t bytes = (size_t)8 << 30;       // 8 GB streamed per launch
    double2 *A; double *out, *small_buf, *big; unsigned long long *clk; int *stop;
    CK(cudaMalloc(&A, bytes)); CK(cudaMemset(A, 0, bytes));
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&clk, 64));
    const size_t nsmall = (size_t)4 << 20, nbig = (size_t)256 << 20;      // 32 MB, 2 GB
    CK(cudaMalloc(&small_buf, nsmall * 8)); CK(cudaMemset(small_buf, 0, nsmall * 8));
    CK(cudaMalloc(&big, nbig * 8)); CK(cudaMemset(big, 0, nbig * 8));
    CK(cudaMallocHost(&stop, 4)); *stop = 0;
    int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s_main, s_side;
    CK(cudaStreamCreateWithPriority(&s_main, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&s_side, cudaStreamNonBlocking, lo));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // dynamic shared memory makes every CTA SM-exclusive (160 + 120 KB and 2 x 120 KB exceed the 227 KB of an SM)
    const int STREAM_SMEM = 160 * 1024, SIDE_SMEM = 120 * 1024;
    CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
    CK(cudaFuncSetAttribute(k_side<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SIDE_SMEM));
    CK(cudaFuncSetAttribute(k_side<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SIDE_SMEM));
    CK(cudaFuncSetAttribute(k_side<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SIDE_SMEM));
    CK(cudaFuncSetAttribute(k_side<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SIDE_SMEM));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_stream)); printf("k_stream regs %d\n", fa.numRegs);
    CK(cudaFuncGetAttributes(&fa, k_side<0>)); printf("k_side<dmma> regs %d\n", fa.numRegs);
    CK(cudaFuncGetAttributes(&fa, k_side<1>)); printf("k_side<l2> regs %d\n", fa.numRegs);

    auto run = [&](const char *name, int mode, int g, int k) {
        *stop = 0;
        if (mode >= 0 && k > 0) {
            const long long iters = 1ll << 40;
            if (mode == 0) k_side<0><<<k, 256, SIDE_SMEM, s_side>>>(small_buf, nsmall, out, stop, iters);
            if (mode == 1) k_side<1><<<k, 256, SIDE_SMEM, s_side>>>(small_buf, nsmall, out, stop, iters);
            if (mode == 2) k_side<2><<<k, 256, SIDE_SMEM, s_side>>>(big, nbig, out, stop, iters);
            if (mode == 3) k_side<3><<<k, 256, SIDE_SMEM, s_side>>>(small_buf, nsmall, out, stop, iters);
        }
        float best = 1e30f; unsigned long long c[2] = {0, 0};
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0, s_main));
            k_stream<<<g, 640, STREAM_SMEM, s_main>>>(A, bytes / 16, out, clk);
            CK(cudaEventRecord(e1, s_main));
            CK(cudaStreamSynchronize(s_main));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) { best = ms; CK(cudaMemcpy(c, clk, 16, cudaMemcpyDeviceToHost)); }
        }
        *stop = 1;
        CK(cudaDeviceSynchronize());
        printf("%-28s G=%3d K=%3d : %8.3f ms  %7.1f GB/s  SM clock during stream %6.0f MHz\n", name, g, k, best, bytes / best / 1e6,
               c[1] ? 1e3 * (double)c[0] / (double)c[1] : 0.0);
        fflush(stdout);
    };
    run("stream alone", -1, 148, 0);
    run("stream alone", -1, G, 0);
    run("stream + dmma side", 0, G, K);
    run("stream + dfma side", 3, G, K);
    run("stream + l2-read side", 1, G, K);
    run("stream + dram-read side", 2, G, K);
    run("stream + dmma side (K/2)", 0, G, K / 2);
    run("stream + dmma side (8 SMs)", 0, G, 8);
    return 0;
}
