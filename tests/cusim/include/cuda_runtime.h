// cuda_runtime.h -- stand-in for the CUDA runtime header, used ONLY by the kernel-logic emulator of the test suite
// (tests/cusim). TEST INFRASTRUCTURE, not part of the product: the shipped library (starneig_b200/lib/libstarneig.so)
// is compiled by nvcc against the real CUDA runtime and never sees this file.
//
// With -DSB_CUSIM the product sources (starneig_b200/csrc/*.cu, *.cuh, node.cpp) are compiled unchanged by g++:
// "device memory" is host memory, a stream executes every operation immediately (a valid schedule: the engine only
// waits on events that were recorded earlier in program order), and a kernel launch runs the kernel body on
// fibers -- one per CUDA thread -- with the block/warp/grid semantics the kernels rely on (cusim_device.h, cusim.cpp).
#pragma once
#include <cstdlib>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>

// ---- vector types -------------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- runtime API (the subset the engine uses) ------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorPeerAccessAlreadyEnabled = 704 };
struct cusimStream { int priority; };
struct cusimEvent { double t_ms; };
typedef cusimStream *cudaStream_t;
typedef cusimEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrCooperativeLaunch = 95 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
struct cudaFuncAttributes { int numRegs; size_t sharedSizeBytes; int maxThreadsPerBlock; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1 };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cusimMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cusimMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
// free / total device memory: 64 GiB, or CUSIM_FREE_MB (tests of the low-memory fall-backs)
static inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b)
{
    const char *e = getenv("CUSIM_FREE_MB");
    *total_b = (size_t)64 << 30;
    *free_b = e ? (size_t)atoll(e) << 20 : *total_b;
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t st);
cudaError_t cudaMemset2DAsync(void *p, size_t pitch, int v, size_t width, size_t height, cudaStream_t st);
cudaError_t cudaMemcpy(void *d, const void *s, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t bytes, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *attr, const void *p);
cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned flags);
static inline cudaError_t cudaMallocHost(void **p, size_t bytes) { *p = malloc(bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, F) { memset(a, 0, sizeof(*a)); return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }

namespace cusim {
// true if [p, p + bytes) lies inside a live cusimMalloc block (CUSIM_CHECK_PREFETCH: prefetch addresses are validated)
bool inside_device_allocation(const void *p, size_t bytes);
// what a CUDA thread knows about itself (threadIdx, blockIdx, ... are macros over this, cusim_device.h)
struct ThreadCtx {
    uint3 t_idx, b_idx;
    dim3 b_dim, g_dim;
    char *smem;             // dynamic shared memory of the block
    unsigned lane, warp;
    ThreadCtx() : t_idx{0, 0, 0}, b_idx{0, 0, 0}, smem(nullptr), lane(0), warp(0) {}
};
extern thread_local ThreadCtx *g_thread;
// runs `body` once per CUDA thread of the grid; cooperative: all blocks are resident at the same time
void run_grid(dim3 grid, dim3 block, size_t smem_bytes, bool cooperative, const std::function<void()> &body);
}
