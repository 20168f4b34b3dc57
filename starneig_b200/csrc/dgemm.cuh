// dgemm.cuh -- FP64 tensor-core (DMMA) GEMM for the rank-nb Hessenberg updates.
//
//   C(MxN) = alpha * op(A) * op(B) + beta * C        all operands column-major FP64
//
// Replaces the cblas_dgemm / cublasDgemm calls of the reference codelets
// (src/hessenberg/cpu.c:315-316,373-375,433-435,492-494,552-554; src/hessenberg/cuda.cu:183,242,303)
// together with their tile gather/scatter (starneig_join_window): the kernels address the dense
// ld-matrix directly, with arbitrary row/column offsets and sizes.
//
// Hardware mapping (B200, sm_100a): FP64 has no tcgen05 kind; the FP64 tensor path is the warp-level
// mma.sync m8n8k4 (SASS DMMA.8x8x4, measured 37.0 TFLOP/s issue peak, profiles/r1_probe_peaks.log).
// One DMMA occupies an SM sub-partition for 16 cycles, so operand traffic is tiny next to the math
// pipe: tiles are staged by 8-byte cp.async (no alignment demands on the sub-matrix origin, zero-fill
// at the edges) through a multi-stage shared-memory ring; fragments come from conflict-free padded
// shared-memory layouts; accumulators stay in registers.
//
// Operand kinds ("MN-major": element (x,k) at g[x + k*ld];  "K-major": element (x,k) at g[k + x*ld]):
//   NT  C -= Y V^T, C -= V W^T, C -= W V^T   A MN-major, B MN-major
//   TN  W  = A^T V                           A K-major,  B K-major
//   NN  W  = A V                             A MN-major, B K-major
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int GEMM_BK = 16;

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src, bool valid)
{
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int src_size = valid ? 8 : 0;      // src_size 0 => the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_size));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool KMAJOR, int BX> struct OperandTile {
    // shared-memory footprint of one stage, in doubles; strides are == 4 (mod 16) doubles so that the
    // 16 lanes of a half-warp (4 k-values x 4 rows) hit 16 distinct 8-byte banks
    static constexpr int STRIDE = KMAJOR ? (GEMM_BK + 4) : (BX + 4);
    static constexpr int SIZE = KMAJOR ? BX * STRIDE : GEMM_BK * STRIDE;
    __device__ static __forceinline__ int offset(int x, int k) { return KMAJOR ? x * STRIDE + k : k * STRIDE + x; }
};

// loads the BX x BK tile of an operand whose logical element (x, k) lives at g[x*sx + k*sk]
template <bool KMAJOR, int BX, int NTHREADS>
__device__ __forceinline__ void load_tile(double *smem, const double *__restrict__ g, int ld,
                                          int x0, int k0, int X, int K, int tid)
{
    using OT = OperandTile<KMAJOR, BX>;
    constexpr int ELEMS = BX * GEMM_BK;
#pragma unroll
    for (int e0 = 0; e0 < ELEMS; e0 += NTHREADS) {
        int e = e0 + tid;
        if (ELEMS % NTHREADS != 0 && e >= ELEMS) break;
        int x, k;
        if (KMAJOR) { x = e / GEMM_BK; k = e % GEMM_BK; }
        else        { k = e / BX;      x = e % BX; }
        int gx = x0 + x, gk = k0 + k;
        bool valid = gx < X && gk < K;
        // clamp the source so that the address is always inside the operand
        int cx = valid ? gx : 0, ck = valid ? gk : 0;
        const double *src = KMAJOR ? g + (size_t)cx * ld + ck : g + (size_t)ck * ld + cx;
        cp_async8(smem + OT::offset(x, k), src, valid);
    }
}

// grid: (ceil(M/BM), ceil(N/BN), ksplits). With ksplits > 1 each z-slice handles k in
// [z*klen, (z+1)*klen) and writes alpha*partial to C + z*split_stride (beta must be 0).
template <bool AK, bool BKM, int WM, int WN, int MB, int NB, int STAGES>
__global__ void __launch_bounds__(WM * WN * 32)
dgemm_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int lda,
             const double *__restrict__ B, int ldb, double beta, double *__restrict__ C, int ldc,
             int klen, size_t split_stride)
{
    constexpr int BM = WM * MB * 8, BN = WN * NB * 8, NT = WM * WN * 32;
    using TA = OperandTile<AK, BM>;
    using TB = OperandTile<BKM, BN>;
    constexpr int STAGE = TA::SIZE + TB::SIZE;
    extern __shared__ double smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * klen;
    const int kend = min(K, kbeg + klen);
    const int ktiles = max(0, (kend - kbeg + GEMM_BK - 1) / GEMM_BK);
    C += (size_t)blockIdx.z * split_stride;

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < ktiles) {
            double *sa = smem + s * STAGE, *sb = sa + TA::SIZE;
            load_tile<AK, BM, NT>(sa, A, lda, m0, kbeg + s * GEMM_BK, M, kend, tid);
            load_tile<BKM, BN, NT>(sb, B, ldb, n0, kbeg + s * GEMM_BK, N, kend, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < ktiles; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch tile kt+STAGES-1 into the slot freed by tile kt-1
            int nk = kt + STAGES - 1;
            if (nk < ktiles) {
                double *sa = smem + (nk % STAGES) * STAGE, *sb = sa + TA::SIZE;
                load_tile<AK, BM, NT>(sa, A, lda, m0, kbeg + nk * GEMM_BK, M, kend, tid);
                load_tile<BKM, BN, NT>(sb, B, ldb, n0, kbeg + nk * GEMM_BK, N, kend, tid);
            }
            cp_async_commit();
        }
        const double *sa = smem + (kt % STAGES) * STAGE, *sb = sa + TA::SIZE;
        const int krem = kend - (kbeg + kt * GEMM_BK);
        const int ksteps = krem >= GEMM_BK ? GEMM_BK / 4 : (krem + 3) / 4;
#pragma unroll
        for (int ks = 0; ks < GEMM_BK / 4; ks++) {
            if (ks < ksteps) {
                double af[MB], bf[NB];
#pragma unroll
                for (int i = 0; i < MB; i++) af[i] = sa[TA::offset((wm * MB + i) * 8 + g, ks * 4 + t)];
#pragma unroll
                for (int j = 0; j < NB; j++) bf[j] = sb[TB::offset((wn * NB + j) * 8 + g, ks * 4 + t)];
#pragma unroll
                for (int i = 0; i < MB; i++)
#pragma unroll
                    for (int j = 0; j < NB; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row g][cols 2t, 2t+1] of each 8x8 block
#pragma unroll
    for (int i = 0; i < MB; i++) {
        int row = m0 + (wm * MB + i) * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NB; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int col = n0 + (wn * NB + j) * 8 + 2 * t + e;
                if (col < N) {
                    double *p = C + (size_t)col * ldc + row;
                    double v = alpha * acc[i][j][e];
                    if (beta != 0.0) v += beta * *p;
                    *p = v;
                }
            }
        }
    }
}

// W(rows x cols) = sum_z part_z   (fixed order => deterministic), part_z at part + z*stride, ld = ldp
__global__ void splitk_reduce_kernel(int rows, int cols, int splits, const double *__restrict__ part, int ldp,
                                     size_t stride, double *__restrict__ W, int ldw)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int c = blockIdx.y;
    if (r >= rows) return;
    double s = 0.0;
    for (int z = 0; z < splits; z++) s += part[(size_t)z * stride + (size_t)c * ldp + r];
    W[(size_t)c * ldw + r] = s;
}

template <bool AK, bool BKM, int WM, int WN, int MB, int NB, int STAGES>
struct GemmConfig {
    static constexpr int BM = WM * MB * 8, BN = WN * NB * 8, NT = WM * WN * 32;
    static constexpr size_t SMEM = (size_t)STAGES * (OperandTile<AK, BM>::SIZE + OperandTile<BKM, BN>::SIZE) * sizeof(double);
    static void prepare()
    {
        SB_CUDA(cudaFuncSetAttribute(dgemm_kernel<AK, BKM, WM, WN, MB, NB, STAGES>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    }
    static void launch(cudaStream_t st, int M, int N, int K, double alpha, const double *A, int lda, const double *B,
                       int ldb, double beta, double *C, int ldc, int splits, int klen, size_t split_stride)
    {
        dim3 grid(ceil_div(M, BM), ceil_div(N, BN), splits);
        dgemm_kernel<AK, BKM, WM, WN, MB, NB, STAGES><<<grid, NT, SMEM, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc,
                                                                              klen, split_stride);
    }
};

} // namespace sb200
