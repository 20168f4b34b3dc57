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

// Per-thread state of the tile loads of one operand, set up once per CTA: in every full k-tile a thread issues
// ITERS cp.async whose source addresses differ by a constant stride and whose validity (x inside the operand) does
// not depend on the k-tile, so the main loop spends ~3 integer instructions per cp.async instead of re-deriving
// (x, k), the bounds checks and the clamped address each time. Partial k-tiles (the last one) use load_tile.
// Out-of-range elements are zero-filled (src-size 0: the address is not dereferenced).
// VEC = 2: a thread moves two neighbouring elements per iteration (neighbours along x for MN-major operands, along k
// for K-major ones) -- with ONE 16-byte cp.async when the operand allows it (16-byte aligned base, even leading
// dimension: `vec16`, decided at run time per operand), else with two 8-byte ones. Half as many iterations either way.
template <bool KMAJOR, int BX, int NTHREADS, int VEC = 1> struct TileLoader {
    static constexpr int UNITS = BX * GEMM_BK / VEC;            // VEC-element units of a tile
    static constexpr int ITERS = (UNITS + NTHREADS - 1) / NTHREADS;
    static constexpr int UPL = (KMAJOR ? GEMM_BK : BX) / VEC;   // units per line (line = fixed k for MN-major, fixed x for K-major)
    static_assert((KMAJOR ? GEMM_BK : BX) % VEC == 0 && NTHREADS % UPL == 0, "thread count must tile the operand");
    // MN-major: x fixed, k = k_t + it * STEP;  K-major: k fixed, x = x_t + it * STEP
    static constexpr int STEP = NTHREADS / UPL;
    using OT = OperandTile<KMAJOR, BX>;
    const double *src;      // first element of iteration 0 in the current k-tile
    size_t it_stride;       // doubles between consecutive iterations
    size_t tile_stride;     // doubles between consecutive k-tiles
    unsigned mask;          // K-major: bit it = the x of iteration `it` is inside the operand; MN-major: valid elements of the unit
    int soff;               // shared-memory offset (doubles) of iteration 0
    bool last_in_tile;      // UNITS % NTHREADS != 0: the unit of the last iteration exists (x < BX)
    bool vec16;             // VEC == 2: one 16-byte copy per unit
    static_assert(KMAJOR || UNITS % NTHREADS == 0, "MN-major operands: the thread count must divide the tile");

    __device__ __forceinline__ void init(const double *g, int ld, int x0, int k0, int X, int tid)
    {
        int x, k;
        if (KMAJOR) { x = tid / UPL; k = (tid % UPL) * VEC; }
        else        { k = tid / UPL; x = (tid % UPL) * VEC; }
        src = KMAJOR ? g + (size_t)(x0 + x) * ld + (k0 + k) : g + (size_t)(k0 + k) * ld + (x0 + x);
        it_stride = (size_t)STEP * ld;
        tile_stride = KMAJOR ? (size_t)GEMM_BK : (size_t)GEMM_BK * ld;
        soff = OT::offset(x, k);
        vec16 = VEC == 2 && ((uintptr_t)g & 15) == 0 && (ld & 1) == 0;      // x0, k0 are even
        mask = 0u;
        last_in_tile = !KMAJOR || x + (ITERS - 1) * STEP < BX;
        if (KMAJOR) {
#pragma unroll
            for (int it = 0; it < ITERS; it++) {
                const int xi = x + it * STEP;
                if (xi < BX && x0 + xi < X) mask |= 1u << it;
            }
        } else {
            mask = (unsigned)max(0, min(VEC, X - (x0 + x)));
        }
    }
    // Full k-tile (all GEMM_BK values of k valid), one iteration at a time, so that the main loop can spread the
    // cp.async of the next stage between the DMMAs of the current one instead of issuing them as one burst in front
    // of them. Iterations must be issued in order 0 .. ITERS-1; `it` is a compile-time constant after unrolling. The
    // last iteration advances to the next k-tile.
    const double *cur;      // source of the next iteration inside the current k-tile
    __device__ __forceinline__ void load_iter(double *smem, int it)
    {
        constexpr int SSTEP = STEP * OT::STRIDE;
        if (it == 0) cur = src;
        // a zero-fill of a unit beyond the tile would land in the next stage (or past the allocation)
        if (UNITS % NTHREADS == 0 || it + 1 < ITERS || last_in_tile) {
            const int nvalid = KMAJOR ? (((mask >> it) & 1u) ? VEC : 0) : (int)mask;
            double *dst = smem + soff + it * SSTEP;
            if (VEC == 1) cp_async8(dst, cur, nvalid != 0);
            else if (vec16) cp_async16(dst, cur, 8 * nvalid);
            else { cp_async8(dst, cur, nvalid >= 1); cp_async8(dst + 1, cur + 1, nvalid >= 2); }
        }
        cur += it_stride;
        if (it == ITERS - 1) src += tile_stride;
    }
    __device__ __forceinline__ void load_full(double *smem)
    {
#pragma unroll
        for (int it = 0; it < ITERS; it++) load_iter(smem, it);
    }
    __device__ __forceinline__ void skip() { src += tile_stride; }
};

// One k-step (4 values of k) of the warp tile: fragments from shared memory, MB x NB DMMAs.
template <class TA, class TB, int MB, int NB>
__device__ __forceinline__ void load_frags(double (&af)[MB], double (&bf)[NB], const double *sa, const double *sb,
                                           int arow, int brow, int kk)
{
#pragma unroll
    for (int i = 0; i < MB; i++) af[i] = sa[TA::offset(arow + i * 8, kk)];
#pragma unroll
    for (int j = 0; j < NB; j++) bf[j] = sb[TB::offset(brow + j * 8, kk)];
}

// `between(d)` is called after the d-th DMMA of the step (d = 0 .. MB*NB-1): the hook through which the main loop
// slips the next stage's cp.async between the tensor instructions
template <int MB, int NB, class F>
__device__ __forceinline__ void mma_tile(double (&acc)[MB][NB][2], const double (&af)[MB], const double (&bf)[NB], F between)
{
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) {
            dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            between(i * NB + j);
        }
}
template <int MB, int NB>
__device__ __forceinline__ void mma_tile(double (&acc)[MB][NB][2], const double (&af)[MB], const double (&bf)[NB])
{
    mma_tile<MB, NB>(acc, af, bf, [](int) {});
}

// grid: (ceil(M/BM), ceil(N/BN), ksplits). With ksplits > 1 each z-slice handles k in
// [z*klen, (z+1)*klen) and writes alpha*partial to C + z*split_stride (beta must be 0).
// MINB resident CTAs per SM are requested so that one CTA's prologue/epilogue (global latency) is
// hidden behind another CTA's DMMA main loop.
// OPT bit 0 (ILV): the cp.async of the next stage are issued between the DMMAs of the current one (see
// TileLoader::load_iter) instead of as one burst after the barrier. OPT bit 1 (V16): two elements per thread and
// iteration, one 16-byte cp.async each where the operand is 16-byte aligned (TileLoader, VEC = 2).
constexpr int GEMM_OPT_ILV = 1, GEMM_OPT_V16 = 2;
template <bool AK, bool BKM, int WM, int WN, int MB, int NB, int STAGES, int MINB, int OPT = 0>
__global__ void __launch_bounds__(WM * WN * 32, MINB)
dgemm_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int lda,
             const double *__restrict__ B, int ldb, double beta, double *__restrict__ C, int ldc,
             int klen, size_t split_stride, int raster)
{
    constexpr int BM = WM * MB * 8, BN = WN * NB * 8, NT = WM * WN * 32;
    constexpr bool ILV = (OPT & GEMM_OPT_ILV) != 0;
    constexpr int VEC = (OPT & GEMM_OPT_V16) ? 2 : 1;
    using TA = OperandTile<AK, BM>;
    using TB = OperandTile<BKM, BN>;
    constexpr int STAGE = TA::SIZE + TB::SIZE;
    SB_DYNAMIC_SMEM(double, smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, t = lane & 3;
    // raster 1 (skinny outputs): consecutive CTAs are the column tiles of ONE row tile, so the long operand's tile is read
    // from DRAM once and from L2 by the others (with rows fastest the re-read came a whole wave later: ncu showed 3x the
    // algorithmic DRAM traffic on W = A^T VT)
    const int m0 = (raster ? blockIdx.y : blockIdx.x) * BM, n0 = (raster ? blockIdx.x : blockIdx.y) * BN;
    const int kbeg = blockIdx.z * klen;
    const int kend = min(K, kbeg + klen);
    const int ktiles = max(0, (kend - kbeg + GEMM_BK - 1) / GEMM_BK);
    C += (size_t)blockIdx.z * split_stride;
    const int arow = wm * MB * 8 + g, brow = wn * NB * 8 + g;

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // pull the C tile towards L2 while the main loop runs (the epilogue reads it when beta != 0)
    if (beta != 0.0) {
        constexpr int LINES_PER_COL = BM / 16;      // 128-byte lines per tile column
        for (int e = tid; e < BN * LINES_PER_COL; e += NT) {
            int col = e / LINES_PER_COL, r = (e % LINES_PER_COL) * 16;
            if (n0 + col < N && m0 + r < M)
                prefetch_l2(C + (size_t)(n0 + col) * ldc + m0 + r);
        }
    }

    TileLoader<AK, BM, NT, VEC> la;
    TileLoader<BKM, BN, NT, VEC> lb;
    la.init(A, lda, m0, kbeg, M, tid);
    lb.init(B, ldb, n0, kbeg, N, tid);
    // stage tile t (if it exists) into ring slot t % STAGES; one commit group per call
    auto issue_tile = [&](int t) {
        if (t < ktiles) {
            double *sa = smem + (t % STAGES) * STAGE, *sb = sa + TA::SIZE;
            const int k0 = kbeg + t * GEMM_BK;
            if (k0 + GEMM_BK <= kend) {
                la.load_full(sa);
                lb.load_full(sb);
            } else {
                load_tile<AK, BM, NT>(sa, A, lda, m0, k0, M, kend, tid);
                load_tile<BKM, BN, NT>(sb, B, ldb, n0, k0, N, kend, tid);
            }
        }
        cp_async_commit();
    };

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) issue_tile(s);

    constexpr int KSTEPS = GEMM_BK / 4;
    for (int kt = 0; kt < ktiles; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        // tile kt + STAGES - 1 goes into the ring slot freed by tile kt - 1 (all warps are past the barrier above)
        const int tn = kt + STAGES - 1;
        const bool next_full = ILV && tn < ktiles && kbeg + (tn + 1) * GEMM_BK <= kend;
        double *san = smem + (tn % STAGES) * STAGE, *sbn = san + TA::SIZE;
        if (!ILV) issue_tile(tn);
        const double *sa = smem + (kt % STAGES) * STAGE, *sb = sa + TA::SIZE;
        const int krem = kend - (kbeg + kt * GEMM_BK);
        if (krem >= GEMM_BK) {
            // full tile: branch-free, fragments of step ks+1 are fetched while step ks is multiplied, and a quarter of
            // the next stage's cp.async follows each step's DMMAs
            double af[2][MB], bf[2][NB];
            load_frags<TA, TB, MB, NB>(af[0], bf[0], sa, sb, arow, brow, t);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ks++) {
                if (ks + 1 < KSTEPS)
                    load_frags<TA, TB, MB, NB>(af[(ks + 1) & 1], bf[(ks + 1) & 1], sa, sb, arow, brow, (ks + 1) * 4 + t);
                // the C = ITERS_A + ITERS_B cp.async of the next stage are spread evenly over the D DMMAs of this tile
                if (!ILV) mma_tile<MB, NB>(acc, af[ks & 1], bf[ks & 1]);
                else mma_tile<MB, NB>(acc, af[ks & 1], bf[ks & 1], [&](int dd) {
                    constexpr int CA = decltype(la)::ITERS, CB = decltype(lb)::ITERS, C = CA + CB, D = KSTEPS * MB * NB;
                    const int d = ks * MB * NB + dd;
#pragma unroll
                    for (int c = d * C / D; c < (d + 1) * C / D; c++) {
                        if (next_full) {
                            if (c < CA) la.load_iter(san, c);
                            else        lb.load_iter(sbn, c - CA);
                        }
                    }
                });
            }
            if (ILV && tn < ktiles && !next_full) {
                const int k0 = kbeg + tn * GEMM_BK;
                load_tile<AK, BM, NT>(san, A, lda, m0, k0, M, kend, tid);
                load_tile<BKM, BN, NT>(sbn, B, ldb, n0, k0, N, kend, tid);
            }
        } else {
            const int ksteps = (krem + 3) / 4;      // the tile is zero-filled beyond kend; it is the last one
            for (int ks = 0; ks < ksteps; ks++) {
                double af[MB], bf[NB];
                load_frags<TA, TB, MB, NB>(af, bf, sa, sb, arow, brow, ks * 4 + t);
                mma_tile<MB, NB>(acc, af, bf);
            }
        }
        if (ILV) cp_async_commit();
    }
    cp_async_wait<0>();

    // Epilogue straight from the accumulator fragments: a thread holds C[row g][cols 2t, 2t+1] of each 8x8 block,
    // so the eight lanes that share t cover eight consecutive rows of one column (64 contiguous bytes): sector-
    // efficient without a trip through shared memory, no CTA-wide barrier, and all loads of a batch (one column
    // block: 2*MB values) are in flight together.
    const bool use_beta = beta != 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int col = n0 + (wn * NB + j) * 8 + 2 * t;
        double *Cc = C + (size_t)col * ldc + m0 + wm * MB * 8 + g;
        const int rbase = m0 + wm * MB * 8 + g;
        double old[MB][2];
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                old[i][e] = (use_beta && col + e < N && rbase + i * 8 < M) ? __ldcs(Cc + (size_t)e * ldc + i * 8) : 0.0;
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                if (col + e < N && rbase + i * 8 < M) {
                    // a read-modify-write pass over C streams through L2 once (evict first): the operands stay resident
                    const double val = fma(alpha, acc[i][j][e], beta * old[i][e]);
                    if (use_beta) __stcs(Cc + (size_t)e * ldc + i * 8, val); else Cc[(size_t)e * ldc + i * 8] = val;
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

template <bool AK, bool BKM, int WM, int WN, int MB, int NB, int STAGES, int MINB, int OPT = 0>
struct GemmConfig {
    static constexpr bool IS_TMA = false;
    static constexpr int BM = WM * MB * 8, BN = WN * NB * 8, NT = WM * WN * 32;
    static constexpr size_t SMEM = (size_t)STAGES * (OperandTile<AK, BM>::SIZE + OperandTile<BKM, BN>::SIZE) * sizeof(double);
    static void prepare()
    {
        SB_CUDA(cudaFuncSetAttribute(dgemm_kernel<AK, BKM, WM, WN, MB, NB, STAGES, MINB, OPT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    }
    static void launch(cudaStream_t st, int M, int N, int K, double alpha, const double *A, int lda, const double *B,
                       int ldb, double beta, double *C, int ldc, int splits, int klen, size_t split_stride, int raster = 0)
    {
        dim3 grid(raster ? ceil_div(N, BN) : ceil_div(M, BM), raster ? ceil_div(M, BM) : ceil_div(N, BN), splits);
        SB_LAUNCH((dgemm_kernel<AK, BKM, WM, WN, MB, NB, STAGES, MINB, OPT>), grid, NT, SMEM, st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc,
                  klen, split_stride, raster);
    }
};

} // namespace sb200
