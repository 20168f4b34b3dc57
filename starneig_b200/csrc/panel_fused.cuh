// panel_fused.cuh -- the whole panel factorisation (w columns) as ONE persistent cooperative kernel.
//
// Same arithmetic as the three per-column kernels of panel.cuh (which stay as the reference path and for
// panels wider than FUSED_MAX_NB), same reference codelets replaced (src/hessenberg/cpu.c:50-285, the
// prepare_column / compute_column / finish_column chain of src/hessenberg/core.c:461-517), but the 3*w
// dependent launches of a panel become one launch: one CTA per SM stays resident, the column loop runs on
// the device and the grid-wide dependencies are device-side barriers (~1 us) instead of kernel boundaries
// plus "last block" reductions. This matters twice: the level-2 side work of a column (re-reading Y, V, VT
// from L2) is latency-bound, and on P GPUs it is replicated on every rank, so it bounds the scaling.
//
// Per column j (c = i + j), all CTAs:
//   [DIST: push own rows of the local GEMV sum to every rank, NVLink peer stores + flags]
//   A   finish column j-1 (V, Y, VT columns, H entries of the panel column), right-update column j,
//       partial w2 = VT^T p'                                            -> grid barrier
//   A'  w2 = sum of the partials (one warp per entry, fixed order)       -> grid barrier
//   R   p'' = p' - V w2, partial ||x||^2 and z = V^T x                   -> grid barrier (+ vote: linear column?)
//   R'  DLARFG scalars, s = scale*z + V(j,:)  (one warp per entry). Linear column (the usual case): look-ahead warps
//       only, off the critical path -- the GEMV warps are already streaming against the unscaled x (FusedArgs::linear)
//   G   trailing GEMV partials, perfectly balanced 1-D split of (row block, column) items over all
//       128-thread groups of the grid                                    -> grid barrier
// Row ownership: CTA b owns rows [b*rpc, (b+1)*rpc) of the panel in every phase except G (rpc = 32*nsub), so V, Y,
// VT columns are written and re-read by the same SM; everything that crosses CTAs inside the launch (pcol, s,
// w2, partials, scalars, row j of V) is read with ld.global.cg.
// (Round 1 carried opt-in variants of this kernel -- LL-entry reductions instead of grid barriers, a single-pass phase R,
// rows spread over all CTAs, L2 prefetch of the GEMV's head, "evict last" loads for columns that every GEMV of a panel
// reads, a 64-register build for co-resident DMMA tiles. All of them
// were timed on B200 at n = 6000 and n = 20000 and lost to this kernel (profiles/r2_v1_switch_sweep_n20000.txt,
// profiles/r2_v4_visit8_sweep_parity_n20000_n50000_8gpu.log, profiles/r1_s8_variant_smoke*_n6000.log); they were removed in round 2.)
#pragma once
#include "panel.cuh"

namespace sb200 {

// Geometry of the CTA (build-time, for A/B experiments on hardware: csrc/Makefile `exp` target): threads, 128-thread GEMV
// groups, 16-byte loads in flight per GEMV thread (U being accumulated + U being fetched).
#ifndef SB_FUSED_THREADS
#define SB_FUSED_THREADS 640
#endif
#ifndef SB_FUSED_VB
#define SB_FUSED_VB 4
#endif
#ifndef SB_GEMV_U
#define SB_GEMV_U 8
#endif
constexpr int FUSED_THREADS = SB_FUSED_THREADS;
constexpr int FUSED_WARPS = FUSED_THREADS / 32;
constexpr int FUSED_MAX_NB = 512;
constexpr int FUSED_VB = SB_FUSED_VB;                // 128-thread GEMV groups per CTA (warps 0..15); the other warps (16..19) look ahead
constexpr int FUSED_LA_BARRIER = FUSED_VB + 1;       // named barrier of the look-ahead warps (1 .. FUSED_VB belong to the GEMV groups)
static_assert(FUSED_THREADS % 32 == 0 && FUSED_THREADS <= 1024 && FUSED_THREADS > 128 * FUSED_VB && FUSED_VB + 1 <= 15, "CTA geometry");
constexpr int FUSED_KC = 512;                        // columns of v staged per group at a time (default; FusedArgs::kc)
constexpr int FUSED_MINSEG = 16;                     // fewest (row block, column) items per group

struct FusedArgs {
    PanelArgs a;
    int w;                      // columns of the panel
    int i;                      // first column (global index)
    double *pan;                // row i+1 of panel column 0 (in A itself or in the replicated panel buffer)
    int ldpan;
    const double *Aloc;         // the rank's column storage
    int lda;
    ColMap cm;
    int lc_end;                 // local index one past the last local column of the reduced block
    int nsub;                   // 32-row sub-tiles per CTA
    int rpc;                    // rows owned by a CTA (32 * nsub)
    int kc;                     // columns of v a GEMV group stages in shared memory at a time (each refill drains its load pipeline)
    unsigned *gbar;             // grid barrier counter, zero at launch
    unsigned long long *rbar;   // one arrival word per panel column for the barrier that ends phase R, zero at launch: arrivals in
                                // the low 32 bits, and above them how many CTAs saw a medium-range (bits 32-47) / a huge (bits
                                // 48-63) entry in their part of x -- the grid-wide vote that decides `linear` for the column
    int linear;                 // 1: GEMV linearity. v = (1, scale x), so A v = A(:, c+1) + scale (A(:, c+2:) x): the GEMV
                                // streams against the UNSCALED x as soon as x is complete (the barrier after phase R) while the
                                // look-ahead warps alone derive beta, tau, scale and s = V^T v; the row owner forms
                                // y(r) = A(r, c+1) + scale g(r) when it sums the partials. Takes phase R' (two dependent L2
                                // round trips, a division and a square root per column) off the critical path. Only for columns
                                // whose x has a medium-range entry and no huge one (so that neither the DLARFG rescaling branch
                                // nor an overflow of A x can occur); other columns (zero x: tau = 0, denormal or huge
                                // entries) take the sequential path.
    unsigned long long *timers; // ns on CTA 0: [0] GEMV phases, [1] whole kernel, [2..5] phases A, A', R, R' (each incl. its barrier)
    Xchg x;                     // DIST: x.epoch = sequence number of the panel's first column (tags of the LL entries of the exchange)
};

// All CTAs of the (cooperative, co-resident) grid: one arrival counter (red.release) that thread 0 of every CTA
// polls. Measured on B200 (tools/bar_bench.cu, 148 CTAs): 1.4 us per barrier, faster than slot arrays with a
// gathering master CTA (2.1 us) or two-level schemes (2.2-2.9 us). `gen` counts the arrivals expected so far.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += gridDim.x;
        red_release_gpu_add(bar, 1u);
        while ((int)(ld_acquire_gpu(bar) - gen) < 0) { }
    }
    __syncthreads();
}

// "LL" exchange entry (the protocol NCCL uses for latency-bound messages): a double travels as two 8-byte words
// {low 32 bits, tag} {high 32 bits, tag}; 8-byte stores are atomic over NVLink, so a reader that sees the expected
// tag in both words has the value, without any fence or separate flag on the critical path.
__device__ __forceinline__ void ll_store(uint4 *dst, double v, unsigned tag)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_volatile_v4(dst, make_uint4((unsigned)bits, tag, (unsigned)(bits >> 32), tag));
}
// `count` (<= N) LL entries p[u * stride], polled TOGETHER until every one of them carries `tag`: a poll round costs one L2
// round trip whatever the number of entries (polling them one after the other costs one round trip EACH -- the first B200
// timing of the LL variant lost ~3 us per column that way). Entries u >= count give 0.
template <int N>
__device__ __forceinline__ void ll_load_batch(const uint4 *p, size_t stride, int count, unsigned tag, unsigned *status, double (&x)[N])
{
    long long t0 = 0;
    for (;;) {
        uint4 e[N];
#pragma unroll
        for (int u = 0; u < N; u++) e[u] = ld_volatile_v4(p + (size_t)min(u, max(count - 1, 0)) * stride);
        bool all = true;
#pragma unroll
        for (int u = 0; u < N; u++) {
            const bool ok = u >= count || (e[u].y == tag && e[u].w == tag);
            all = all && ok;
            x[u] = u < count ? __longlong_as_double((long long)(((unsigned long long)e[u].z << 32) | e[u].x)) : 0.0;
        }
        if (all || count <= 0) return;
        if (t0 == 0) { t0 = clock64(); if (*(volatile unsigned *)status != 0u) return; }
        else if (clock64() - t0 > 8000000000ll) { atomicExch(status, 2u); return; }
    }
}

// sum over the CTAs that own rows (nblk <= 32*5) of part[bb*ldt], fixed order, by one warp: the (up to) five loads
// of a lane are independent
__device__ __forceinline__ double sum_over_ctas(const double *part, size_t ldt, int nblk, int lane)
{
    double acc = 0.0;
    for (int b0 = 0; b0 < nblk; b0 += 160) {
        double x[5];
#pragma unroll
        for (int u = 0; u < 5; u++) {
            const int bb = b0 + lane + 32 * u;
            x[u] = bb < nblk ? __ldcg(part + (size_t)bb * ldt) : 0.0;
        }
        acc += ((x[0] + x[1]) + (x[2] + x[3])) + x[4];
    }
    return warp_sum(acc);
}

// the same for three arrays part + k*stride (unit stride inside each), all loads in flight together
__device__ __forceinline__ void sum3_over_ctas(const double *part, int stride, int nblk, int lane, double (&out)[3])
{
    double acc[3] = {0.0, 0.0, 0.0};
    for (int b0 = 0; b0 < nblk; b0 += 160) {
        double x[3][5];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int bb = b0 + lane + 32 * u;
                x[k][u] = bb < nblk ? __ldcg(part + k * stride + bb) : 0.0;
            }
#pragma unroll
        for (int k = 0; k < 3; k++) acc[k] += ((x[k][0] + x[k][1]) + (x[k][2] + x[k][3])) + x[k][4];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) out[k] = warp_sum(acc[k]);
}

// split of the GEMV over 128-thread groups: items = (row block, local column) pairs in row-block-major order
struct GemvSplit {
    int skip, RB, nloc, per;
    __device__ __forceinline__ int first_group(int rb) const { return (int)(((long long)rb * nloc) / per); }
    __device__ __forceinline__ int last_group(int rb) const { return (int)((((long long)(rb + 1)) * nloc - 1) / per); }
};

// sum of S partial values p[z*ld], z < S, in a fixed order; the loads of a batch are independent
__device__ __forceinline__ double sum_partials(const double *p, int ld, int S)
{
    double e = 0.0;
    int z = 0;
    for (; z + 8 <= S; z += 8) {
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = __ldcg(p + (size_t)(z + u) * ld);
        e += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    }
    for (; z + 4 <= S; z += 4) {
        const double x0 = __ldcg(p + (size_t)z * ld), x1 = __ldcg(p + (size_t)(z + 1) * ld);
        const double x2 = __ldcg(p + (size_t)(z + 2) * ld), x3 = __ldcg(p + (size_t)(z + 3) * ld);
        e += (x0 + x1) + (x2 + x3);
    }
    for (; z < S; z++) e += __ldcg(p + (size_t)z * ld);
    return e;
}

// Column-wise dots of one CTA: out[t] = sum over the CTA's rows of M(r, t) * p(r) for t in [tb, tb+NC), t < tmax.
// One warp; NC columns x NS sub-tiles = 32 independent loads are in flight per lane. p(r) comes from shared
// memory (pv[sub*32 + lane], zero for rows >= m). Result: see transpose_reduce8 / transpose_reduce32.
// `rbase`: row index of M's first stored row (0: a matrix in global memory; row0: the CTA's slab in shared memory).
template <int NC, int NS>
__device__ __forceinline__ void coldots(const double *__restrict__ M, int ld, int rbase, int tb, int tmax, int row0, int m, int nsub,
                                        const double *pv, int lane, double (&acc)[NC])
{
#pragma unroll
    for (int q = 0; q < NC; q++) acc[q] = 0.0;
    for (int sub0 = 0; sub0 < nsub; sub0 += NS) {
        double v[NS][NC];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int r = row0 + (sub0 + s) * 32 + lane;
            const bool okr = sub0 + s < nsub && r < m;
            const double *Mr = M + (size_t)tb * ld + (r - rbase);
#pragma unroll
            for (int q = 0; q < NC; q++) v[s][q] = (okr && tb + q < tmax) ? Mr[(size_t)q * ld] : 0.0;
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const double p = sub0 + s < nsub ? pv[(sub0 + s) * 32 + lane] : 0.0;
#pragma unroll
            for (int q = 0; q < NC; q++) acc[q] = fma(v[s][q], p, acc[q]);
        }
    }
}

// Sum over the 32 lanes of a warp of 16 per-lane values: on return lane l holds the total of column l & 15.
__device__ __forceinline__ double transpose_reduce16(double (&v)[16], int lane)
{
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; k++) {
            const double send = upper ? v[k] : v[k + off];
            const double keep = upper ? v[k + off] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// partial column dots of the CTA -> colpart[b][t], t < tmax; warps [0, T) share the column batches
__device__ __forceinline__ void coldots_all(const double *__restrict__ M, int ld, int rbase, int tmax, int row0, int m, int nsub,
                                            const double *pv, int wp, int T, int lane, double *out)
{
    if (nsub >= 3) {
        for (int tb = 8 * wp; tb < tmax; tb += 8 * T) {
            double acc[8];
            coldots<8, 4>(M, ld, rbase, tb, tmax, row0, m, nsub, pv, lane, acc);
            const double xx = transpose_reduce8(acc, lane);
            const int t = tb + (lane >> 2);
            if ((lane & 3) == 0 && t < tmax) out[t] = xx;
        }
    } else {
        for (int tb = 16 * wp; tb < tmax; tb += 16 * T) {
            double acc[16];
            coldots<16, 2>(M, ld, rbase, tb, tmax, row0, m, nsub, pv, lane, acc);
            const double xx = transpose_reduce16(acc, lane);
            if (lane < 16 && tb + lane < tmax) out[tb + lane] = xx;
        }
    }
}

constexpr int FUSED_GEMV_WARPS = 4 * FUSED_VB;                       // warps 0..15: the GEMV groups of phase G
constexpr int FUSED_SHADOW_THREADS = FUSED_THREADS - 32 * FUSED_GEMV_WARPS;      // warps 16..19: look-ahead during phase G

// shared-memory layout (doubles), fixed for the whole launch
// `slabs` (0 or 1): the CTA keeps its own rows of V for the whole panel in shared memory (w columns of 32 * nsub rows) next
// to the copy in global memory (which the level-3 updates and the other CTAs read): phase R reads those rows twice per
// column, from L2 at ~0.8 us per dependent round trip. Measured on B200 (profiles/r2_v14_sweep_slabs.txt,
// r2_v15_sweep_l1_split.txt): -2 % at AED-window sizes (n = 1000 ... 4000, width 224) as long as the launch stays at or below
// ~130 KB of shared memory; slabs of VT and Y as well (read by phase A and by the look-ahead warps) LOSE, and so does any
// layout beyond ~190 KB: the GEMV keeps up to 128 KB of loads in flight per SM and L1 is what is left of 256 KB (a launch
// padded to 224 KB streams 21 % slower, one at 190 KB 1 % slower, below 150 KB no difference).
struct FusedSmem {
    int vs, s, vrow, w2, red, pv, ysm, ysum, sqred, scal, slab, slab_doubles, total;
    __host__ __device__ FusedSmem(int w, int nsub, int kc = FUSED_KC, int slabs = 0)
    {
        const int NW = (w + 31) / 32 > 1 ? (w + 31) / 32 : 1;
        const int wp8 = (w + 8) / 8 * 8;
        int o = 0;
        vs = o;    o += FUSED_VB * kc;
        s = o;     o += wp8;
        vrow = o;  o += wp8;
        w2 = o;    o += wp8;
        red = o;   o += nsub * 3 * NW * 32;
        pv = o;    o += nsub * 32;
        ysm = o;   o += nsub * 32;
        ysum = o;  o += FUSED_THREADS;          // per-thread shares of the GEMV partial sums of a row (phase A)
        sqred = o; o += 3 * FUSED_WARPS;
        scal = o;  o += 4;
        slab = o;  slab_doubles = w * nsub * 32; o += slabs * slab_doubles;
        total = o;
    }
};

// LAPACK's DLARFG rescaling branch inside the persistent kernel (a column in the denormal range; every CTA takes it
// together): the row owners multiply x by 2^969 before anybody forms v = x * scale, and z = V^T x is taken again from the
// rescaled x (the first one was summed in denormal arithmetic, too coarse for T). Two more grid barriers in such a
// column. Kept out of line: it must not cost the common path registers. Returns the z entry of this warp.
// (All arguments by value: taking the address of the kernel's parameter block or of `gen` would move them to local memory.)
// The caller advances its barrier generation by two barriers.
__device__ __noinline__ double fused_rescale_x(double *pcol, const double *V, int ld, double *colpart, int ldt, int m, int nsub, int rpc,
                                               unsigned *gbar, unsigned gen, double *pv, int j, double xmul, int t_first)
{
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int b = blockIdx.x;
    const int row0 = b * rpc;
    const int row_end = min(m, row0 + rpc);
    const int rows_here = max(0, row_end - row0);
    const int nblk = (m + rpc - 1) / rpc;
    grid_barrier(gbar, gen);                // nobody reads the old z partials any more
    for (int rr = tid; rr < rows_here; rr += FUSED_THREADS)
        if (row0 + rr > j) pcol[row0 + rr] *= xmul;
    for (int rr = tid; rr < nsub * 32; rr += FUSED_THREADS) pv[rr] *= xmul;
    __syncthreads();
    if (j > 0) coldots_all(V, ld, 0, j, row0, row_end, nsub, pv, wp, FUSED_WARPS, lane, colpart + (size_t)b * ldt);
    grid_barrier(gbar, gen);
    return t_first < j ? sum_over_ctas(colpart + t_first, ldt, nblk, lane) : 0.0;
}

// The kernel. 640 threads x 96 registers: the CTA owns the register file of its SM (one CTA per SM, cooperative launch).
// SLABS = 1: the CTA's rows of V resident in shared memory (FusedSmem); chosen per launch by what fits.
template <bool DIST, int SLABS>
__global__ void __launch_bounds__(FUSED_THREADS, 1) k_panel_fused(FusedArgs f)
{
    // 16-byte loads in flight per GEMV thread: U columns being accumulated + U being fetched
    constexpr int GEMV_U = SB_GEMV_U;
    SB_DYNAMIC_SMEM(double, sh);
    const PanelArgs &a = f.a;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int m = a.m, ld = a.ld, nsub = f.nsub;
    const int G = gridDim.x, b = blockIdx.x;
    // CTA b owns rows [row0, row_end): rpc rows (rpc = 32 * nsub, or fewer so that every CTA of the grid owns rows)
    const int row0 = b * f.rpc;
    const int row_end = min(m, row0 + f.rpc);
    const int rows_here = max(0, row_end - row0);
    const int nblk = (m + f.rpc - 1) / f.rpc;                   // CTAs that own rows
    const FusedSmem L(f.w, nsub, f.kc, SLABS);
    // the CTA's slabs: element (row row0 + rr, panel column t) at [t * rpc + rr]
    double *const Vs = sh + L.slab;
    double *const vs_all = sh + L.vs, *const s_sh = sh + L.s, *const vrow_sh = sh + L.vrow, *const w2_sh = sh + L.w2;
    double *const red = sh + L.red, *const pv = sh + L.pv, *const ysm = sh + L.ysm, *const sqred = sh + L.sqred;
    double *const ysum = sh + L.ysum;
    unsigned gen = 0, gen2 = 0;
    bool lin = false, lin_prev = false;     // the GEMV of this / of the previous column runs (ran) against the unscaled x
    unsigned *const bar2 = f.gbar + 32;         // arrival counter "s of this column is complete" (look-ahead warps only)
    unsigned long long t_gemv = 0, t_begin = 0, t_ph[4] = {0, 0, 0, 0}, t_mark = 0;
    const bool timer = (b == 0 && tid == 0);
    if (timer) t_begin = t_mark = globaltimer_ns();
#define SB_PHASE_MARK(k) do { if (timer) { const unsigned long long now_ = globaltimer_ns(); t_ph[k] += now_ - t_mark; t_mark = now_; } } while (0)

    // GEMV geometry that does not depend on the column
    GemvSplit gs;
    {
        // Row blocks start on a 128-byte line of the matrix (up to 15 rows above the panel's first row; those rows are
        // read and dropped), so that a warp's 512 contiguous bytes of a column are four whole lines: measured on B200,
        // blocks that start mid-line (panels at i = 8 mod 16, every other panel of the default width 312) stream 4 % slower.
        // Needs columns that all start on a line (128-byte aligned storage, leading dimension a multiple of 16);
        // otherwise only the 16-byte alignment of the loads is kept (parity of the row offset).
        const double *base = f.Aloc + (size_t)f.i + 1;
        const bool lines = ((uintptr_t)f.Aloc & 127) == 0 && (f.lda & 15) == 0;
        gs.skip = (int)(((uintptr_t)base / sizeof(double)) & (lines ? 15 : 1));
        gs.RB = (m + gs.skip + 255) >> 8;
    }
    double *const scal_sh = sh + L.scal;    // tau, beta, scale of the current column

    // look-ahead results for column 0 -> 1 do not exist yet: phase A(1) must see zeros (no previous columns)
    for (int t = tid; t < nsub * 3 * 32; t += FUSED_THREADS) red[t] = 0.0;

    for (int j = 0; j <= f.w; j++) {
        const int jm1 = j - 1;
        lin_prev = lin;
        double *acol = f.pan + (size_t)j * f.ldpan;
        const int NW = max(1, (j + 31) >> 5);
        double *const pc_cur = a.pcol;              // the column being reduced: p', then p''
        const double *const pc_prev = a.pcol;

        if (j > 0) {
            // ================= phase A: finish column j-1, start column j =================
            // The row-wise products with Y(:, :j-1) and VT(:, :j-1) were computed by the look-ahead warps while the
            // GEMV of column j-1 was streaming (`red`); what is left on the critical path is y itself.
            const int NWa = max(1, (jm1 + 31) >> 5);
            const bool do_update = j < f.w;
            const double tau = __ldcg(&a.scal[jm1].tau), beta_prev = __ldcg(&a.scal[jm1].beta), scale_prev = __ldcg(&a.scal[jm1].scale);
            // the epilogue's inputs that do not depend on the GEMV (p'' of column j-1 and column j of the panel, own rows of the
            // warp's first sub-tile) are fetched before the partial sums instead of after them: one L2 round trip less
            double pre_p = 0.0, pre_a = 0.0;
            if (wp < nsub) {
                const int r = row0 + wp * 32 + lane;
                if (r < row_end) { pre_p = pc_prev[r]; pre_a = acol[r]; }       // (j == w: the column right of the panel)
            }
            double *acol_prev = f.pan + (size_t)jm1 * f.ldpan;
            {
                // y(r) of the CTA's rows: sum of the local GEMV partials of column j-1 (fixed order); on P GPUs the
                // sums travel to every rank as self-validating 16-byte entries and the P contributions are added
                const int nloc_prev = f.lc_end - f.cm.lower(f.i + jm1 + 1);
                GemvSplit gp = gs;
                gp.nloc = nloc_prev;
                gp.per = max(FUSED_MINSEG, (int)(((long long)gs.RB * nloc_prev + G * FUSED_VB - 1) / (G * FUSED_VB)));
                const unsigned epoch = f.x.epoch + jm1;
                const int par = epoch & 1;
                // The S partials of a row (one per GEMV group that touched its row block; S = 8 at m = 20000 on one GPU, but 15-25
                // on 8 GPUs and 70+ for matrices of order 2000) are shared out over the `parts` threads that the CTA has per
                // owned row, so that they are fetched in ONE batch of independent loads instead of S / 8 dependent batches;
                // the shares meet in shared memory and are added in a fixed order.
                const int RP = (rows_here + 31) & ~31;
                const int parts = RP > 0 ? FUSED_THREADS / RP : 0;
                if (parts >= 2) {
                    double e = 0.0;
                    if (tid < parts * RP) {
                        const int part = tid / RP, rr = tid - part * RP, r = row0 + rr;
                        if (rr < rows_here && nloc_prev > 0) {
                            const int rb = (r + gs.skip) >> 8;
                            const int S = gp.last_group(rb) - gp.first_group(rb) + 1;
                            const double *p = a.ypart + r;
                            int z = part;
                            for (; z + 7 * parts < S; z += 8 * parts) {
                                double x[8];
#pragma unroll
                                for (int u = 0; u < 8; u++) x[u] = __ldcg(p + (size_t)(z + u * parts) * a.ldp);
                                e += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
                            }
                            double x[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) x[u] = (z + u * parts < S) ? __ldcg(p + (size_t)(z + u * parts) * a.ldp) : 0.0;
                            e += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
                        }
                        ysum[tid] = e;
                    }
                    __syncthreads();
                }
                for (int rr = tid; rr < rows_here; rr += FUSED_THREADS) {
                    const int r = row0 + rr;
                    const int rb = (r + gs.skip) >> 8;
                    double sum = 0.0;
                    if (parts >= 2) {
                        for (int q = 0; q < parts; q++) sum += ysum[q * RP + rr];
                    } else if (nloc_prev > 0) {
                        const int S = gp.last_group(rb) - gp.first_group(rb) + 1;
                        sum = sum_partials(a.ypart + r, a.ldp, S);
                    }
                    if (DIST) {
                        // all-reduce over the ranks: the rank's sum goes to every inbox (the own one included, so that all
                        // P entries of a row are polled alike and together), the P contributions are added in rank order
                        const size_t slot = ((size_t)par * f.x.P + f.x.g) * a.ldp + r;
                        for (int d = 0; d < f.x.P; d++) ll_store((uint4 *)f.x.inbox[(f.x.g + d) % f.x.P] + slot, sum, epoch);
                        const uint4 *in = (const uint4 *)f.x.inbox[f.x.g] + (size_t)par * f.x.P * a.ldp + r;
                        double val[MAX_RANKS];
                        ll_load_batch<MAX_RANKS>(in, (size_t)a.ldp, f.x.P, epoch, f.x.status, val);
                        sum = 0.0;
#pragma unroll
                        for (int q = 0; q < MAX_RANKS; q++)
                            if (q < f.x.P) sum += val[q];
                    }
                    ysm[rr] = sum;
                }
            }
            __syncthreads();

            // ---- per-row epilogue, one warp per sub-tile
            for (int sub = wp; sub < nsub; sub += FUSED_WARPS) {
                const int r = row0 + sub * 32 + lane;
                const bool valid = r < row_end;
                double pp = 0.0;
                if (valid) {
                    const bool pre = sub == wp;
                    const double pprev = pre ? pre_p : pc_prev[r];
                    const double ac_raw = pre ? pre_a : ((do_update || lin_prev) ? acol[r] : 0.0);
                    const double ac = do_update ? ac_raw : 0.0;
                    // the GEMV of column j-1 was linear: its partials are g = A(:, c+1:) x and column j of the panel (not
                    // yet updated) is A(:, c), the column that belongs to the leading one of v
                    double D3 = ysm[sub * 32 + lane];
                    // (j == w: column i + w, the first one right of the panel; it exists -- the last panel ends at end - 2)
                    if (lin_prev) D3 = fma(scale_prev, D3, ac_raw);
                    const double *rd = red + (size_t)sub * 3 * NWa * 32 + lane;
                    double D0 = 0.0, D1 = 0.0, D2 = 0.0;
                    for (int q = 0; q < NWa; q++) { D0 += rd[q * 32]; D1 += rd[(NWa + q) * 32]; D2 += rd[(2 * NWa + q) * 32]; }
                    // column j-1 of V and of the reduced matrix (the row owner writes: no cross-CTA traffic)
                    const double vr = (r < jm1) ? 0.0 : (r == jm1 ? 1.0 : pprev * scale_prev);
                    a.V[(size_t)jm1 * ld + r] = vr;
                    if (SLABS >= 1) Vs[(size_t)jm1 * f.rpc + sub * 32 + lane] = vr;
                    if (r == jm1) acol_prev[r] = beta_prev;
                    else if (r > jm1) acol_prev[r] = 0.0;
                    const double ynew = tau * (D3 - D0);                 // finish_column: Y(:,j-1)
                    a.Y[(size_t)jm1 * ld + r] = ynew;
                    a.VT[(size_t)jm1 * ld + r] = tau * (vr - D2);        // VT(:,j-1) = V * T(:,j-1)
                    if (do_update) {
                        pp = ac - (D1 + ynew);                           // prepare_column: p - Y V(j-1,:)^T, V(j-1,j-1) = 1
                        pc_cur[r] = pp;
                    }
                }
                pv[sub * 32 + lane] = pp;
            }
            if (!do_update) break;          // j == w: the panel is complete
            __syncthreads();

            // ---- w2part[t] = sum over the CTA's rows of VT(r,t) * p'(r), t < j (column j-1 was stored just above)
            coldots_all(a.VT, ld, 0, j, row0, row_end, nsub, pv, wp, FUSED_WARPS, lane, a.colpart + (size_t)b * a.ldt);
            grid_barrier(f.gbar, gen);
            SB_PHASE_MARK(0);

            // ================= A': w2[t] = sum over CTAs (one warp per entry) =================
            for (int t = b * FUSED_WARPS + wp; t < j; t += G * FUSED_WARPS) {
                const double acc = sum_over_ctas(a.colpart + t, a.ldt, nblk, lane);
                if (lane == 0) a.w2[t] = acc;
            }
            grid_barrier(f.gbar, gen);
            SB_PHASE_MARK(1);
        }

        // ================= phase R: p'' = p' - V w2; ||x||^2, z = V^T x =================
        {
            // (Loading the warp's first tile of V before w2 arrives was measured: the 64 registers it holds across the barrier
            // spill and phase R doubles, profiles/r2_v13_sweep_ysum.txt.)
            for (int t = tid; t < j; t += FUSED_THREADS) w2_sh[t] = __ldcg(a.w2 + t);
            __syncthreads();
            if (j > 0) {
                // d(r) = V(r, :j) w2: 32x32 tiles, 32 loads in flight per lane
                for (int item = wp; item < nsub * NW; item += FUSED_WARPS) {
                    const int sub = item / NW, g = item - sub * NW, t0 = g * 32;
                    const int r = row0 + sub * 32 + lane;
                    const bool valid = r < row_end;
                    const double *Vr = SLABS >= 1 ? Vs + (size_t)t0 * f.rpc + sub * 32 + lane : a.V + (size_t)t0 * ld + r;
                    const size_t vstep = SLABS >= 1 ? (size_t)f.rpc : (size_t)ld;
                    double v32[32];
#pragma unroll
                    for (int q = 0; q < 32; q++) v32[q] = (valid && t0 + q < j) ? Vr[(size_t)q * vstep] : 0.0;
                    double d = 0.0;
#pragma unroll
                    for (int q = 0; q < 32; q++) d = fma(v32[q], w2_sh[min(t0 + q, j - 1)], d);
                    red[((size_t)sub * NW + g) * 32 + lane] = d;
                }
            }
            __syncthreads();
            SumSq sq;
            sq.clear();
            for (int sub = wp; sub < nsub; sub += FUSED_WARPS) {
                const int r = row0 + sub * 32 + lane;
                const bool valid = r < row_end;
                double xx = 0.0;
                if (valid) {
                    // p' of the row: this thread left it in pv at the end of phase A (one L2 round trip less than pc_cur[r])
                    double pp = j > 0 ? pv[sub * 32 + lane] : acol[r];
                    if (j > 0) {
                        double D = 0.0;
                        for (int q = 0; q < NW; q++) D += red[((size_t)sub * NW + q) * 32 + lane];
                        pp -= D;
                    }
                    pc_cur[r] = pp;
                    if (r < j) acol[r] = pp;            // final entries of H above the sub-diagonal
                    if (r == j) a.scal[j].alpha = pp;
                    if (r > j) xx = pp;
                }
                sq.add(xx);
                pv[sub * 32 + lane] = xx;
            }
            // huge / tiny entries (SumSq) are rare: their sums only cost a vote when there are none
            sq.med = warp_sum(sq.med);
            if (__any_sync(0xffffffffu, sq.big != 0.0 || sq.sml != 0.0)) { sq.big = warp_sum(sq.big); sq.sml = warp_sum(sq.sml); }
            if (lane == 0) { sqred[wp] = sq.med; sqred[FUSED_WARPS + wp] = sq.big; sqred[2 * FUSED_WARPS + wp] = sq.sml; }
            __syncthreads();
            if (j > 0) {
                if (SLABS >= 1) coldots_all(Vs, f.rpc, row0, j, row0, row_end, nsub, pv, wp, FUSED_WARPS, lane, a.colpart + (size_t)b * a.ldt);
                else            coldots_all(a.V, ld, 0, j, row0, row_end, nsub, pv, wp, FUSED_WARPS, lane, a.colpart + (size_t)b * a.ldt);
            }
            if (tid < 3) {
                double sum = 0.0;
                for (int q = 0; q < FUSED_WARPS; q++) sum += sqred[tid * FUSED_WARPS + q];
                a.sqpart[tid * PANEL_LDB + b] = sum;
            }
            // grid barrier on the column's own arrival word; the arrivals carry the vote on `linear`
            __syncthreads();
            if (wp == 0) {
                // lane q looks at the sums of warp q: !(x == 0) so that a NaN counts as an entry
                const bool any_med = __any_sync(0xffffffffu, lane < FUSED_WARPS && !(sqred[min(lane, FUSED_WARPS - 1)] == 0.0));
                const bool any_big = __any_sync(0xffffffffu, lane < FUSED_WARPS && !(sqred[FUSED_WARPS + min(lane, FUSED_WARPS - 1)] == 0.0));
                if (lane == 0) {
                    red_release_gpu_add_u64(f.rbar + j, 1ull | (any_med ? 1ull << 32 : 0ull) | (any_big ? 1ull << 48 : 0ull));
                    unsigned long long seen;
                    while ((unsigned)((seen = ld_acquire_gpu_u64(f.rbar + j)) & 0xffffffffull) < (unsigned)G) { }
                    scal_sh[3] = (f.linear && ((seen >> 32) & 0xffffull) != 0ull && (seen >> 48) == 0ull) ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            lin = scal_sh[3] != 0.0;
            SB_PHASE_MARK(2);
        }

        // ================= R': DLARFG scalars (every warp, identical), s = scale*z + V(j,:) =================
        // `lin`: only the look-ahead warps run it (the GEMV warps are already streaming A against the unscaled x)
        if (!lin || wp >= FUSED_GEMV_WARPS) {
            // every warp derives the scalars itself (no CTA-wide wait); the z entries it owns are loaded alongside
            const int nwr = lin ? FUSED_WARPS - FUSED_GEMV_WARPS : FUSED_WARPS;
            const int wpr = lin ? wp - FUSED_GEMV_WARPS : wp;
            const int t_first = b * nwr + wpr;
            double zsum = 0.0, vjt = 0.0;
            if (t_first < j) { zsum = sum_over_ctas(a.colpart + t_first, a.ldt, nblk, lane); vjt = __ldcg(a.V + (size_t)t_first * ld + j); }
            double ssq[3];
            sum3_over_ctas(a.sqpart, PANEL_LDB, nblk, lane, ssq);
            const double alpha = __ldcg(&a.scal[j].alpha);
            const Reflector rf = dlarfg_scalars(alpha, ssq[0], ssq[1], ssq[2], m - j > 1);
            const double tau = rf.tau, beta = rf.beta, scale = rf.scale;
            if (wpr == 0 && lane == 0) {
                scal_sh[0] = tau; scal_sh[1] = beta; scal_sh[2] = scale;
                if (b == 0) { a.scal[j].tau = tau; a.scal[j].beta = beta; a.scal[j].scale = scale; }
            }
            if (rf.xmul != 1.0) {       // denormal-range column (never `lin`: all warps are here)
                zsum = fused_rescale_x(pc_cur, a.V, ld, a.colpart, a.ldt, m, nsub, f.rpc, f.gbar, gen, pv, j, rf.xmul, t_first);
                gen += 2 * G;
            }
            if (t_first < j && lane == 0) a.s[t_first] = fma(scale, zsum, vjt);
            for (int t = t_first + G * nwr; t < j; t += G * nwr) {
                const double acc = sum_over_ctas(a.colpart + t, a.ldt, nblk, lane);
                if (lane == 0) a.s[t] = fma(scale, acc, __ldcg(a.V + (size_t)t * ld + j));
            }
            if (lin) group_barrier(FUSED_LA_BARRIER, FUSED_SHADOW_THREADS); else __syncthreads();
            // "my part of s is written": only the look-ahead warps wait for this, the GEMV starts at once
            if (wpr == 0 && lane == 0) red_release_gpu_add(bar2, 1u);
        }

        // ================= phase G: GEMV partials (warps 0..11) + look-ahead for column j+1 (warps 12..15) ==========
        {
            SB_PHASE_MARK(3);
            if (timer) t_gemv -= globaltimer_ns();
            // `lin`: the GEMV runs against x itself; y(r) = A(r, c+1) + scale g(r) is formed by the row owner in phase A
            const double scale = lin ? 1.0 : scal_sh[2];
            const double v_first = lin ? 0.0 : 1.0;
            if (wp < FUSED_GEMV_WARPS) {
                const int c = f.i + j;
                const int lc0 = f.cm.lower(c + 1), nloc = f.lc_end - lc0;
                const int gc0 = c + 1;
                GemvSplit gq = gs;
                gq.nloc = nloc;
                const long long items = (long long)gs.RB * nloc;
                gq.per = max(FUSED_MINSEG, (int)((items + G * FUSED_VB - 1) / (G * FUSED_VB)));
                const int vb = tid >> 7, vt = tid & 127;
                // groups are spread over the CTAs first (group v lives on CTA v % G), so that a GEMV smaller than the
                // grid still touches every SM
                const int v = vb * G + b;
                double *vs = vs_all + vb * f.kc;
                long long it = (long long)v * gq.per;
                const long long it_end = min(items, it + gq.per);
                const int mp = m + gs.skip;
                while (it < it_end) {
                    const int rb = (int)(it / nloc);
                    const int cbeg = (int)(it - (long long)rb * nloc);
                    const int cend = (int)min((long long)nloc, cbeg + (it_end - it));
                    const int rp = rb * 256 + vt * 2;
                    const bool rows_ok = rp < mp;
                    double2 acc = make_double2(0.0, 0.0);
                    const double *Ap = f.Aloc + (size_t)lc0 * f.lda + f.i + 1 - gs.skip + rp;
                    for (int k0 = cbeg, nk = 0; k0 < cend; k0 += nk) {
                        nk = min(f.kc, cend - k0);
                        constexpr int U = GEMV_U;
                        const size_t step = (size_t)f.lda;
                        const double *P0 = Ap + (size_t)k0 * step;
                        double2 cur[U], nxt[U];
                        // (Issuing the first loads of the block here, before v is staged, was measured: the registers they hold
                        // across the staging code spill and the kernel loses 2 %, profiles/r2_v10_sweep_microopts.txt.)
                        group_barrier(1 + vb, 128);          // previous chunk's vs fully consumed
                        for (int k = vt; k < nk; k += 128) {
                            const int kk = f.cm.l2g(lc0 + k0 + k) - gc0;
                            vs[k] = (kk == 0) ? v_first : __ldcg(pc_cur + j + kk) * scale;
                        }
                        group_barrier(1 + vb, 128);
                        if (rows_ok) {
                            int k = 0;
                            if (nk >= U) {
#pragma unroll
                                for (int u = 0; u < U; u++) cur[u] = __ldcs((const double2 *)(P0 + u * step));
                                const double *Pn = P0 + U * step;
                                for (; k + 2 * U <= nk; k += U) {
#pragma unroll
                                    for (int u = 0; u < U; u++) nxt[u] = __ldcs((const double2 *)(Pn + u * step));
                                    Pn += U * step;
#pragma unroll
                                    for (int u = 0; u < U; u++) {
                                        const double vk = vs[k + u];
                                        acc.x = fma(cur[u].x, vk, acc.x);
                                        acc.y = fma(cur[u].y, vk, acc.y);
                                    }
#pragma unroll
                                    for (int u = 0; u < U; u++) cur[u] = nxt[u];
                                }
#pragma unroll
                                for (int u = 0; u < U; u++) {
                                    const double vk = vs[k + u];
                                    acc.x = fma(cur[u].x, vk, acc.x);
                                    acc.y = fma(cur[u].y, vk, acc.y);
                                }
                                k += U;
                            }
                            for (; k < nk; k++) {
                                const double vk = vs[k];
                                const double2 xx = __ldcs((const double2 *)(P0 + (size_t)k * step));
                                acc.x = fma(xx.x, vk, acc.x);
                                acc.y = fma(xx.y, vk, acc.y);
                            }
                        }
                    }
                    if (rows_ok) {
                        const size_t slot = (size_t)(v - gq.first_group(rb)) * a.ldp;
                        const int r = rp - gs.skip;
                        double *yp = a.ypart + slot;
                        if (r >= 0) yp[r] = acc.x;
                        if (r + 1 >= 0 && r + 1 < m) yp[r + 1] = acc.y;
                    }
                    it += cend - cbeg;
                }
            } else {
                // ---- look-ahead for phase A of column j+1: D0 = Y(:, :j) s_j, D1 = Y(:, :j) V(j, :j)^T, D2 = VT(:, :j) s_j.
                // None of them needs the GEMV result, so they run in the shadow of the HBM-bound GEMV.
                const int xt = tid - 32 * FUSED_GEMV_WARPS, xw = wp - FUSED_GEMV_WARPS;
                gen2 += G;
                if (xt == 0) while ((int)(ld_acquire_gpu(bar2) - gen2) < 0) { }
                group_barrier(FUSED_LA_BARRIER, FUSED_SHADOW_THREADS);
                for (int t = xt; t < j; t += FUSED_SHADOW_THREADS) {
                    s_sh[t] = __ldcg(a.s + t);
                    vrow_sh[t] = __ldcg(a.V + (size_t)t * ld + j);
                }
                group_barrier(FUSED_LA_BARRIER, FUSED_SHADOW_THREADS);
                const int NWn = max(1, (j + 31) >> 5);
                for (int item = xw; item < nsub * NWn; item += FUSED_WARPS - FUSED_GEMV_WARPS) {
                    const int sub = item / NWn, g = item - sub * NWn, t0 = g * 32;
                    const int r = row0 + sub * 32 + lane;
                    const bool valid = r < row_end;
                    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
                    const double *VTr = a.VT + (size_t)t0 * ld + r;
                    const double *Yr = a.Y + (size_t)t0 * ld + r;
#pragma unroll
                    for (int bt = 0; bt < 2; bt++) {
                        const int tb = t0 + 16 * bt;
                        if (tb < j) {
                            double y16[16], v16[16];
#pragma unroll
                            for (int q = 0; q < 16; q++) {
                                const bool ok = valid && tb + q < j;
                                y16[q] = ok ? Yr[(size_t)(16 * bt + q) * ld] : 0.0;
                                v16[q] = ok ? VTr[(size_t)(16 * bt + q) * ld] : 0.0;
                            }
#pragma unroll
                            for (int q = 0; q < 16; q++) {
                                const int t = min(tb + q, j - 1);        // values beyond j are zero
                                const double sv = s_sh[t];
                                d0 = fma(y16[q], sv, d0);
                                d2 = fma(v16[q], sv, d2);
                                d1 = fma(y16[q], vrow_sh[t], d1);
                            }
                        }
                    }
                    double *rd = red + ((size_t)sub * 3 * NWn + g) * 32 + lane;
                    rd[0] = d0; rd[NWn * 32] = d1; rd[2 * NWn * 32] = d2;
                }
            }
            grid_barrier(f.gbar, gen);
            if (timer) { t_mark = globaltimer_ns(); t_gemv += t_mark; }
        }
    }
    if (timer) {
        f.timers[0] += t_gemv;
        f.timers[1] += globaltimer_ns() - t_begin;
        for (int k = 0; k < 4; k++) f.timers[2 + k] += t_ph[k];
    }
#undef SB_PHASE_MARK
}

// dynamic shared memory of k_panel_fused for a panel of w columns with nsub sub-tiles per CTA
static inline size_t fused_smem_bytes(int w, int nsub, int kc = FUSED_KC, int slabs = 0) { return (size_t)FusedSmem(w, nsub, kc, slabs).total * sizeof(double); }

} // namespace sb200
