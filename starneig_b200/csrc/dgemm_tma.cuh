// dgemm_tma.cuh -- FP64 tensor-core (DMMA) GEMM whose operand tiles are moved by the TMA engine.
//
//   C(MxN) = alpha * op(A) * op(B) + beta * C        all operands column-major FP64 (same contract as dgemm.cuh)
//
// Same products and the same reference call sites as dgemm.cuh (src/hessenberg/cpu.c:315-316,373-375,433-435,492-494,
// 552-554; src/hessenberg/cuda.cu:183,242,303). What changes is who moves the tiles. In dgemm.cuh every thread of the CTA
// issues 8-byte cp.async (SASS LDGSTS) with its own address arithmetic: ~100 non-tensor instructions per k-tile and warp
// next to 128 DMMAs, and (ncu, profiles/r2_v1_ncu_full_fused_and_dgemm.txt) a tensor pipe that is busy only 62-72 % of the
// time. Here ONE thread of a dedicated producer warp issues cp.async.bulk.tensor.2d (SASS UTMALDG) per box; the hardware
// computes the addresses, zero-fills everything outside the operand (no edge code), writes the tile into shared memory in
// the 128-byte swizzle pattern and signals an mbarrier with the byte count. The four (or eight) consumer warps execute
// nothing but LDS.64 + DMMA.8x8x4 in their main loop, and full/empty mbarriers per ring slot replace the CTA-wide barrier
// per k-tile: a warp that is ahead does not wait for the slowest one.
//
// Shared-memory layout = what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B: rows of 16 doubles (128 bytes), the 16-byte
// chunk index of an element XORed with (row & 7).
//   MN-major operand (element (x, k) at g[x + k*ld]):  boxes of 16 x-values (row = k): one TMA per 16 rows of the tile
//   K-major  operand (element (x, k) at g[k + x*ld]):  one box of 16 k-values times BX rows (row = x): one TMA per tile
// Fragments: lane 4g+t of a DMMA holds A(x = g, k = t). With k = 4s + t (s = step inside the 16-wide k-tile) the 16 lanes
// of a half-warp would hit only 8 of the 16 eight-byte banks in either layout (2-way conflict). A DMMA sums over four k
// values and does not care which: step s takes k in {0,3,12,15}, {1,2,13,14}, {4,7,8,11}, {5,6,9,10} (same sets for both
// operands, so the products pair up correctly) and every fragment load of either layout is conflict-free (brute-force
// check: tools/swizzle_check.py). Only the order of the k-sum inside a k-tile differs from dgemm.cuh.
//
// Alignment: TMA addresses global memory in 16-byte units -- base, strides AND the start of every box. (Measured on B200: a
// box whose coordinate in the contiguous dimension is odd, i.e. 8 bytes off, raises "illegal instruction".) Leading
// dimensions must therefore be even, and an operand that starts at an odd element is handled by moving the FRAME of the
// product by one: A' = A - 1 element covers one more row (MN-major: the row frame of C moves, the extra output row is
// masked in the epilogue) or one more k (K-major: the k frame moves for both operands; the extra term must vanish, which
// the caller guarantees by keeping a zero in front of one operand -- the engine's V / VT / Y workspaces have such a guard
// row and are stored with the row parity of the panel in A, so that both K-major operands of W = A^T VT agree). Products
// that cannot be framed that way (odd leading dimension, K-major operands of different parity, no guard) use dgemm.cuh.
#pragma once
#include "dgemm.cuh"

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// tensor maps (host)
// ---------------------------------------------------------------------------------------------
#ifndef SB_CUSIM
typedef CUresult (*sb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline sb_encode_tiled_fn sb_encode_tiled()
{
    static sb_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) fatal("cuTensorMapEncodeTiled is not available in this driver", __FILE__, __LINE__);
        fn = (sb_encode_tiled_fn)p;
    }
    return fn;
}
// 2-D FP64 tensor: dim0 contiguous, dim1 strided by ld doubles; box0 x box1 elements per copy, 128-byte swizzle, zero fill
static inline void sb_make_tensor_map(SbTensorMap *map, const double *base, unsigned long long dim0, unsigned long long dim1,
                                      unsigned long long ld, unsigned box0, unsigned box1)
{
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * sizeof(double)};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = sb_encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[starneig] cuTensorMapEncodeTiled: result %d, base %p, dims %llu x %llu, ld %llu, box %u x %u\n", (int)r, (const void *)base,
                dim0, dim1, ld, box0, box1);
        fatal("cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    }
}
#endif

constexpr int TMA_BK = 16;                  // k-tile = one 128-byte swizzle row of a K-major operand
constexpr int TMA_ROW = 16;                 // doubles per swizzle row

// k of DMMA step s, fragment lane t (see the header): conflict-free fragment loads for both operand layouts
__host__ __device__ constexpr int tma_kperm(int s, int t)
{
    return s == 0 ? (t == 0 ? 0 : t == 1 ? 3 : t == 2 ? 12 : 15)
         : s == 1 ? (t == 0 ? 1 : t == 1 ? 2 : t == 2 ? 13 : 14)
         : s == 2 ? (t == 0 ? 4 : t == 1 ? 7 : t == 2 ? 8 : 11)
                  : (t == 0 ? 5 : t == 1 ? 6 : t == 2 ? 9 : 10);
}

// offset (doubles) of element (x, k) inside a swizzled tile
template <bool KMAJOR> __host__ __device__ __forceinline__ int tma_tile_offset(int x, int k)
{
    if (KMAJOR) return x * TMA_ROW + ((((k >> 1) ^ (x & 7)) << 1) | (k & 1));
    return (x >> 4) * (TMA_ROW * TMA_BK) + k * TMA_ROW + (((((x & 15) >> 1) ^ (k & 7)) << 1) | (x & 1));
}

struct GemmTmaArgs {
    int M, N, K;
    double alpha, beta;
    double *C;
    int ldc;
    int klen;                   // k range of one z-slice (multiple of TMA_BK unless there is one slice)
    size_t split_stride;        // doubles between the outputs of consecutive z-slices
    int raster;                 // 1: blockIdx.x runs over the column tiles (skinny outputs), see dgemm.cuh
    int row_min, col_min;       // first row / column of the (shifted) frame that belongs to C: 1 where the frame was moved
};

// grid: as dgemm_kernel. WM x WN consumer warps (each MB x NB blocks of 8 x 8).
// PW = true:  + one producer warp; full/empty mbarriers per ring slot, no CTA-wide barrier in the main loop. (The register
//             file is allocated in pairs of warps: the fifth warp of a 4-consumer CTA costs the registers of two.)
// PW = false: thread 0 issues the copies for tile kt + STAGES - 1 right after the CTA-wide barrier that ends tile kt - 1
//             (the ring discipline of dgemm.cuh); only the `full` barriers are used.
// KBOX: rows of a K-major operand per bulk copy (0: the whole tile in one copy).
template <bool AK, bool BKM, int WM, int WN, int MB, int NB, int STAGES, int MINB, bool PW, int KBOX>
__global__ void __launch_bounds__((WM * WN + (PW ? 1 : 0)) * 32, MINB)
dgemm_tma_kernel(const SB_GRID_CONSTANT SbTensorMap mapA, const SB_GRID_CONSTANT SbTensorMap mapB, const GemmTmaArgs p)
{
    constexpr int BM = WM * MB * 8, BN = WN * NB * 8, NCW = WM * WN;
    constexpr int A_TILE = BM * TMA_BK, B_TILE = BN * TMA_BK, STAGE = A_TILE + B_TILE;     // doubles
    constexpr unsigned STAGE_BYTES = STAGE * sizeof(double);
    static_assert(BM % 16 == 0 && BN % 8 == 0 && (AK || BM % 16 == 0) && (BKM || BN % 16 == 0), "tile shape");
    static_assert((A_TILE * 8) % 1024 == 0 && (B_TILE * 8) % 1024 == 0, "every tile starts on a swizzle atom");
    SB_DYNAMIC_SMEM(double, smem_raw);
    __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES];
    double *const smem = sb_align_shared(smem_raw, 1024);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = (p.raster ? blockIdx.y : blockIdx.x) * BM, n0 = (p.raster ? blockIdx.x : blockIdx.y) * BN;
    const int kbeg = blockIdx.z * p.klen;
    const int kend = min(p.K, kbeg + p.klen);
    const int ktiles = max(0, (kend - kbeg + TMA_BK - 1) / TMA_BK);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1u); mbar_init(&empty_bar[s], (unsigned)NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    // one ring slot: expected bytes, then the boxes of both operands
    auto issue_tile = [&](int t) {
        const int s = t % STAGES;
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        double *sa = smem + s * STAGE, *sb = sa + A_TILE;
        const int k0 = kbeg + t * TMA_BK;
        if (AK) {
            constexpr int R = KBOX > 0 ? KBOX : BM;
#pragma unroll
            for (int bx = 0; bx < BM / R; bx++) tma_load_2d(sa + bx * R * TMA_ROW, &mapA, k0, m0 + R * bx, &full_bar[s]);
        } else {
#pragma unroll
            for (int bx = 0; bx < BM / 16; bx++) tma_load_2d(sa + bx * (TMA_ROW * TMA_BK), &mapA, m0 + 16 * bx, k0, &full_bar[s]);
        }
        if (BKM) {
            constexpr int R = KBOX > 0 ? KBOX : BN;
#pragma unroll
            for (int bx = 0; bx < BN / R; bx++) tma_load_2d(sb + bx * R * TMA_ROW, &mapB, k0, n0 + R * bx, &full_bar[s]);
        } else {
#pragma unroll
            for (int bx = 0; bx < BN / 16; bx++) tma_load_2d(sb + bx * (TMA_ROW * TMA_BK), &mapB, n0 + 16 * bx, k0, &full_bar[s]);
        }
    };

    if (PW && warp == NCW) {
        // ---------------- producer: one thread feeds the ring ----------------
        if (lane == 0) {
            for (int t = 0; t < ktiles; t++) {
                mbar_wait(&empty_bar[t % STAGES], (unsigned)(((t / STAGES) & 1) ^ 1));         // slot free (first round: at once)
                issue_tile(t);
            }
        }
        return;
    }
    if (!PW && tid == 0) {
        for (int t = 0; t < STAGES - 1 && t < ktiles; t++) issue_tile(t);
    }

    // ---------------- consumers ----------------
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, t4 = lane & 3;
    const int arow = wm * MB * 8 + g, brow = wn * NB * 8 + g;
    // per-thread fragment offsets of the four DMMA steps: element (row + 8 i, k_s) = base(i) + offset of (g [+ 8], k_s)
    int offa[4], offb[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int k = t4 == 0 ? tma_kperm(s, 0) : t4 == 1 ? tma_kperm(s, 1) : t4 == 2 ? tma_kperm(s, 2) : tma_kperm(s, 3);
        offa[s] = tma_tile_offset<AK>(arow, k);
        offb[s] = tma_tile_offset<BKM>(brow, k);
    }
    double *C = p.C + (size_t)blockIdx.z * p.split_stride;

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // pull the C tile towards L2 while the main loop runs (the epilogue reads it when beta != 0)
    if (p.beta != 0.0) {
        constexpr int LINES_PER_COL = BM / 16;
        for (int e = tid; e < BN * LINES_PER_COL; e += NCW * 32) {        // (consumer threads only: tid < NCW * 32)
            const int col = e / LINES_PER_COL, r = (e % LINES_PER_COL) * 16;
            if (n0 + col < p.N && n0 + col >= p.col_min && m0 + r + p.row_min < p.M) prefetch_l2(C + (size_t)(n0 + col) * p.ldc + m0 + r + p.row_min);
        }
    }

    // rows 8 i of an operand: K-major: 8 rows further down (128 doubles); MN-major: the other half of the 16-wide box
    // (chunk index ^ 4, i.e. offset ^ 8) or the next box
    auto frag_a = [&](const double *sa, int s, int i) -> double {
        if (AK) return sa[offa[s] + i * 8 * TMA_ROW];
        return sa[(offa[s] ^ ((i & 1) << 3)) + (i >> 1) * (TMA_ROW * TMA_BK)];
    };
    auto frag_b = [&](const double *sb, int s, int j) -> double {
        if (BKM) return sb[offb[s] + j * 8 * TMA_ROW];
        return sb[(offb[s] ^ ((j & 1) << 3)) + (j >> 1) * (TMA_ROW * TMA_BK)];
    };
    // MN-major with an odd number of 8-row blocks below the warp's first row: arow may start in the upper half of a box
    // (arow & 8): offa already carries that half, the XOR flips it per block -- (i & 1) counts from the warp's own start
    // only if MB is even or the warp starts on a box boundary; guaranteed by the static_asserts in the launcher.

    for (int kt = 0; kt < ktiles; kt++) {
        const int slot = kt % STAGES;
        if (!PW) {
            // every warp is done with tile kt - 1: its slot takes tile kt + STAGES - 1
            __syncthreads();
            if (tid == 0 && kt + STAGES - 1 < ktiles) { fence_proxy_async(); issue_tile(kt + STAGES - 1); }
        }
        mbar_wait(&full_bar[slot], (unsigned)((kt / STAGES) & 1));
        const double *sa = smem + slot * STAGE, *sb = sa + A_TILE;
        double af[2][MB], bf[2][NB];
#pragma unroll
        for (int i = 0; i < MB; i++) af[0][i] = frag_a(sa, 0, i);
#pragma unroll
        for (int j = 0; j < NB; j++) bf[0][j] = frag_b(sb, 0, j);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (s + 1 < 4) {
#pragma unroll
                for (int i = 0; i < MB; i++) af[(s + 1) & 1][i] = frag_a(sa, s + 1, i);
#pragma unroll
                for (int j = 0; j < NB; j++) bf[(s + 1) & 1][j] = frag_b(sb, s + 1, j);
            }
            mma_tile<MB, NB>(acc, af[s & 1], bf[s & 1]);
        }
        if (PW) {       // this warp is done with the slot
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
        }
    }

    // Epilogue straight from the accumulator fragments (as dgemm.cuh): the eight lanes that share t cover eight
    // consecutive rows of one column (64 contiguous bytes); a read-modify-write pass streams through L2 (evict first)
    const bool use_beta = p.beta != 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int col = n0 + (wn * NB + j) * 8 + 2 * t4;
        const int rbase = m0 + wm * MB * 8 + g;
        double *Cc = C + (size_t)col * p.ldc + rbase;
        double old[MB][2];
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                old[i][e] = (use_beta && col + e < p.N && col + e >= p.col_min && rbase + i * 8 < p.M && rbase + i * 8 >= p.row_min)
                                ? __ldcs(Cc + (size_t)e * p.ldc + i * 8) : 0.0;
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                if (col + e < p.N && col + e >= p.col_min && rbase + i * 8 < p.M && rbase + i * 8 >= p.row_min) {
                    const double val = fma(p.alpha, acc[i][j][e], p.beta * old[i][e]);
                    if (use_beta) __stcs(Cc + (size_t)e * p.ldc + i * 8, val); else Cc[(size_t)e * p.ldc + i * 8] = val;
                }
    }
}

template <bool AK, bool BKM, int WM, int WN, int MB, int NB, int STAGES, int MINB, bool PW = false, int KBOX = 0>
struct GemmTmaConfig {
    static constexpr bool IS_TMA = true;
    static constexpr int BM = WM * MB * 8, BN = WN * NB * 8, NT = (WM * WN + (PW ? 1 : 0)) * 32;
    // MN-major operands are fetched in boxes of 16 rows; a warp's blocks alternate between the halves of a box
    static_assert(AK || (BM % 16 == 0 && (MB % 2 == 0)), "MN-major A: warps must start on a 16-row box");
    static_assert(BKM || (BN % 16 == 0 && (NB % 2 == 0)), "MN-major B: warps must start on a 16-row box");
    static constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * TMA_BK * sizeof(double) + 1024;      // + alignment slack
    static void prepare()
    {
        SB_CUDA(cudaFuncSetAttribute(dgemm_tma_kernel<AK, BKM, WM, WN, MB, NB, STAGES, MINB, PW, KBOX>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    }
    // Same arguments as GemmConfig::launch plus `k_guard`: the caller guarantees that, where an operand is K-major and starts
    // at an odd element, the element in front of it (k = -1) is a finite number in both operands and ZERO in at least one.
    // Returns false (nothing launched) if the product cannot be framed for TMA; the caller then uses the cp.async kernel.
    static bool launch(cudaStream_t st, int M, int N, int K, double alpha, const double *A, int lda, const double *B,
                       int ldb, double beta, double *C, int ldc, int splits, int klen, size_t split_stride, int raster = 0,
                       bool k_guard = false)
    {
        if (M < 1 || N < 1 || K < 1 || (lda & 1) || (ldb & 1) || ((uintptr_t)A & 7) || ((uintptr_t)B & 7)) return false;
        const int pa = (int)(((uintptr_t)A >> 3) & 1), pb = (int)(((uintptr_t)B >> 3) & 1);
        int dm = 0, dn = 0, dk = 0;
        if (AK) dk = pa; else dm = pa;
        if (BKM) { if (AK && pb != dk) return false; dk = pb; } else dn = pb;
        if (dk && !k_guard) return false;
        if (dk && splits > 1 && (klen % TMA_BK) != 0) return false;
        // the shifted frame: every pointer below is 16-byte aligned
        const double *Af = AK ? A - dk : A - dm - (size_t)dk * lda;
        const double *Bf = BKM ? B - dk : B - dn - (size_t)dk * ldb;
        GemmTmaArgs p;
        p.M = M + dm; p.N = N + dn; p.K = K + dk; p.alpha = alpha; p.beta = beta; p.ldc = ldc;
        p.C = C - dm - (ptrdiff_t)dn * ldc;
        p.klen = splits > 1 ? klen : p.K; p.split_stride = split_stride; p.raster = raster;
        p.row_min = dm; p.col_min = dn;
        if (splits > 1 && (long long)splits * klen < p.K) return false;       // the moved k frame needs one more k than the slices cover
        SbTensorMap mapA, mapB;
        // dim0 = the contiguous dimension (x for MN-major, k for K-major)
        static_assert(KBOX == 0 || (KBOX % 8 == 0 && (!AK || BM % KBOX == 0) && (!BKM || BN % KBOX == 0)), "K-major boxes: whole swizzle atoms");
        if (AK) sb_make_tensor_map(&mapA, Af, p.K, p.M, lda, TMA_BK, KBOX > 0 ? KBOX : BM);
        else    sb_make_tensor_map(&mapA, Af, p.M, p.K, lda, 16, TMA_BK);
        if (BKM) sb_make_tensor_map(&mapB, Bf, p.K, p.N, ldb, TMA_BK, KBOX > 0 ? KBOX : BN);
        else     sb_make_tensor_map(&mapB, Bf, p.N, p.K, ldb, 16, TMA_BK);
        M = p.M; N = p.N;
        dim3 grid(raster ? ceil_div(N, BN) : ceil_div(M, BM), raster ? ceil_div(M, BM) : ceil_div(N, BN), splits);
        SB_LAUNCH((dgemm_tma_kernel<AK, BKM, WM, WN, MB, NB, STAGES, MINB, PW, KBOX>), grid, NT, SMEM, st, mapA, mapB, p);
        return true;
    }
};

} // namespace sb200
