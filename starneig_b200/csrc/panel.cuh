// panel.cuh -- the per-column kernels of the panel factorisation (the critical path).
//
// Reference codelets replaced (all CPU-only or naive-CUDA in the reference):
//   prepare_column  src/hessenberg/cpu.c:50-161   (column update, DLARFG, zeroing)
//   compute_column  src/hessenberg/cpu.c:163-224, cuda.cu:62-150   (the trailing GEMV  y = A v)
//   finish_column   src/hessenberg/cpu.c:226-285  (Y(:,j), T(:,j))
//
// Formulation. The reference keeps the compact-WY factor T explicitly and applies it with small
// triangular matrix-vector products on the critical path. Here the product VT := V*T (m x w) is kept
// instead. It obeys the same recurrence as Y,
//     s        = V(:, :j)^T v_j
//     Y(:,j)   = tau_j * ( A v_j - Y(:, :j)  s )
//     VT(:,j)  = tau_j * (   v_j - VT(:, :j) s )          ( == V * T(:, j) )
// and gives the column update   p <- p - V * ( VT^T p )   ( == (I - V T V^T)^T p )  without any
// triangular solve or T mat-vec, and the block updates as  W = X * VT  ( == X V T ). Mathematically
// identical to the reference; only rounding differs.
//
// Per column j of a panel three kernels run back to back on one stream:
//   k_col_finish_update  (j >= 1) finishes column j-1 (Y, VT), applies the right update
//                        p' = p - Y V(j-1,:)^T and reduces  w2 = VT^T p'
//   k_col_reflector      p'' = p' - V w2; reduces ||p''(j+1:)||^2 and z = V(j+1:,:)^T p''(j+1:); the last
//                        block does the DLARFG scalar work and s = scale*z + V(j,:)^T  ( == V^T v )
//   k_col_gemv           the HBM-bound GEMV over the trailing matrix with v formed on the fly from
//                        p'' and the scale; also writes V(:,j), the exact zeros and beta into A
//
// Tile scheme of the first two kernels: a block owns 32*nsub rows, split into sub-tiles of 32 rows (one
// row per lane); warp (g, h) owns the 32 columns [32g, 32g+32) of the sub-tiles h, h+RS, ... Phase A
// streams the tiles once for the row-wise dot products (reduced across column groups through shared
// memory), a per-row epilogue forms the new column entries, phase B streams the tile a second time (L2)
// for the column-wise dot products, reduced across lanes by a shuffle transpose-butterfly. Loads are
// issued in independent batches of 8 columns with no barrier inside a phase.
// Cross-block reductions write per-block partials; the block that finishes last adds them in a fixed
// order, so results are bitwise reproducible run to run.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int PANEL_MAX_BLOCKS = 148;   // row blocks of the panel kernels (<= one per SM)
constexpr int PANEL_LDB = 160;          // leading dimension of the per-block partial arrays
constexpr int PANEL_MAX_NB = 1024;      // widest panel the warp-per-32-columns scheme supports
constexpr int GEMV_THREADS = 128;

constexpr int MAX_RANKS = 8;            // GPUs of one NVSwitch box
constexpr int RB_MAX = 512;             // row blocks (256 rows) of the exchanged GEMV result: n <= 131000

struct ColScal {        // DLARFG results for one column
    double tau, beta, scale, alpha;
};

// 1-D block-cyclic column map: global column c lives on rank (c / cb) % P. The local columns
// of a rank are ordered by global index, so any global column range [a, b) is the contiguous local range
// [lower(a), lower(b)). P == 1 is the identity map.
// ||x|| without overflow / underflow of the squares: Blue's three-accumulator scheme, the one LAPACK's dnrm2 uses
// (the reference reaches dnrm2 through dlarfg_, src/hessenberg/cpu.c:140). Entries in the "medium" range
// [2^-511, 2^486] -- all that occurs for sanely scaled matrices -- accumulate exactly like a plain sum of squares;
// huge / tiny entries are scaled before they are squared and summed separately.
constexpr double SUMSQ_TSML = 1.4916681462400413e-154, SUMSQ_TBIG = 1.9979190722022350e+146;
constexpr double SUMSQ_SSML = 4.4989137945431964e+161, SUMSQ_SBIG = 1.1113793747425387e-162;
struct SumSq {
    double med, big, sml;
    __device__ __forceinline__ void clear() { med = big = sml = 0.0; }
    __device__ __forceinline__ void add(double x)
    {
        const double ax = fabs(x);
        if (ax > SUMSQ_TBIG) { const double t = x * SUMSQ_SBIG; big = fma(t, t, big); }
        else if (ax < SUMSQ_TSML) { const double t = x * SUMSQ_SSML; sml = fma(t, t, sml); }
        else med = fma(x, x, med);          // NaN lands here as well
    }
};
// the norm from the three totals (LAPACK 3.10 dnrm2)
__device__ __forceinline__ double sumsq_norm(double med, double big, double sml)
{
    if (big > 0.0) {
        if (med > 0.0 || med != med) big += (med * SUMSQ_SBIG) * SUMSQ_SBIG;
        return sqrt(big) / SUMSQ_SBIG;
    }
    if (sml > 0.0) {
        if (med > 0.0 || med != med) {
            const double a = sqrt(med), b = sqrt(sml) / SUMSQ_SSML;
            const double ymin = fmin(a, b), ymax = fmax(a, b);
            const double q = ymin / ymax;
            return ymax * sqrt(1.0 + q * q);
        }
        return sqrt(sml) / SUMSQ_SSML;
    }
    return sqrt(med);
}

// DLARFG scalars (LAPACK dlarfg, called by the reference at src/hessenberg/cpu.c:140) from alpha and the three sums
// of squares of x: beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = x * scale, scale = 1 / (alpha - beta).
// LAPACK's rescaling branch (|beta| < safmin = 2^-969: x, alpha are multiplied by 1/safmin = 2^969 and the norm is taken
// again; in IEEE double one round always suffices) is reproduced without a second pass over x: such an x lives entirely
// in the small-range accumulator, sml = sum (x 2^537)^2 exactly scaled, so ||2^969 x|| = sqrt(sml) 2^432. Then `xmul`
// = 2^969 is returned: the caller multiplies x (its copy in pcol, and z = V^T x) by it before v = x * scale is formed,
// because 2^969 * scale itself may overflow. beta is scaled back (it may be denormal, as in LAPACK).
struct Reflector { double tau, beta, scale, xmul; };
constexpr double DLARFG_SAFMIN = 2.0041683600089728e-292;      // dlamch('S') / dlamch('E') = 2^-969
constexpr double DLARFG_RSAFMN = 4.9896007738368e+291;         // 2^969
__device__ __forceinline__ Reflector dlarfg_scalars(double alpha, double med, double big, double sml, bool has_x)
{
    Reflector f;
    f.tau = 0.0; f.beta = alpha; f.scale = 0.0; f.xmul = 1.0;
    const double xnorm = sumsq_norm(med, big, sml);
    if (!has_x || xnorm == 0.0) return f;
    double beta = -copysign(hypot(alpha, xnorm), alpha);
    if (fabs(beta) < DLARFG_SAFMIN) {
        const double as = alpha * DLARFG_RSAFMN;
        const double xs = sqrt(sml) * 1.1090678776483259e+130;                 // 2^432 = 2^969 / 2^537
        beta = -copysign(hypot(as, xs), as);
        f.tau = (beta - as) / beta;
        f.scale = 1.0 / (as - beta);
        f.beta = beta * DLARFG_SAFMIN;
        f.xmul = DLARFG_RSAFMN;
        return f;
    }
    f.tau = (beta - alpha) / beta;
    f.scale = 1.0 / (alpha - beta);
    f.beta = beta;
    return f;
}

struct ColMap {
    int P, g, cb;
    __host__ __device__ int l2g(int lc) const { return ((lc / cb) * P + g) * cb + lc % cb; }
    __host__ __device__ int lower(int x) const      // number of owned columns with global index < x
    {
        const int B = x / cb, r = B % P;
        return (B / P) * cb + (r > g ? cb : (r == g ? x % cb : 0));
    }
    __host__ __device__ int owner(int c) const { return (c / cb) % P; }
};

// Cross-GPU exchange of the GEMV result (P > 1): every rank pushes its partial y (summed over its column
// chunks) into inbox[parity][g] of every rank over NVLink peer stores, 256 rows at a time, and then raises
// yflag[parity][g][row block] = epoch on that rank. Consumers (k_col_finish_update) poll their own flags.
struct Xchg {
    int P, g;
    unsigned epoch;             // sequence number of this column, monotonic over the life of the arena
    double *inbox[MAX_RANKS];   // [2][P][ldp] on every rank
    unsigned *yflag[MAX_RANKS]; // [2][P][RB_MAX] on every rank
    unsigned *rbcount;          // local, RB_MAX arrival counters (zero between launches)
    unsigned *status;           // local: set to non-zero when a wait timed out
};

// spin until *p has reached `want` (acquire; sequence numbers, wrap-around safe); gives up after ~4 s and
// reports through *status. ">=": a peer that already left a barrier may have announced the next one.
__device__ __forceinline__ bool flag_reached(const unsigned *p, unsigned want) { return (int)(ld_acquire_sys(p) - want) >= 0; }
__device__ __forceinline__ void wait_flag(const unsigned *p, unsigned want, unsigned *status, unsigned code = 1u)
{
    if (flag_reached(p, want)) return;
    if (*(volatile unsigned *)status != 0u) return;      // an earlier wait already failed: do not stall again
    const long long t0 = clock64();
    while (!flag_reached(p, want)) {
        if (clock64() - t0 > 8000000000ll) { atomicExch(status, code); break; }
        __nanosleep(20);
    }
}

struct PanelArgs {
    int m;              // rows of the panel (= end - i - 1); local row r <-> global row i+1+r
    int ld;             // leading dimension of V, Y, VT
    double *V, *Y, *VT;
    double *pcol;       // m   scratch: the column being reduced
    double *ypart;      // S x ldp  GEMV partial sums
    int ldp;
    double *s;          // nb  s = V^T v
    double *w2;         // nb
    double *colpart;    // PANEL_LDB x ldt  per-block partials of the column-wise dot products (column index contiguous)
    int ldt;
    double *sqpart;     // 3 x PANEL_LDB: per-block sums of squares (medium, big, small range; see SumSq)
    ColScal *scal;      // nb
    unsigned *counter;  // zero-initialised
};

// Sum over the 32 lanes of a warp of 32 per-lane values each: on return lane l holds
// sum_over_lanes(v[l]). 16+8+4+2+1 = 31 shuffles instead of 32 five-step butterflies.
__device__ __forceinline__ double transpose_reduce32(double (&v)[32], int lane)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; k++) {
            double send = upper ? v[k] : v[k + off];
            double keep = upper ? v[k + off] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// out[t] = sum_{b < nblocks} part[b*ldt + t] for t < ncols, by the calling block; fixed summation order.
// Warp g owns columns [32g, 32g+32) (+ multiples of 32*nwarps); lane = block index within a chunk of 32
// blocks: 32 independent loads per lane, then the transpose-butterfly. `apply(t, sum)` is called by the
// lane that ends up owning column t.
template <typename F>
__device__ __forceinline__ void reduce_block_partials(const double *part, int ldt, int ncols, int nblocks, F apply)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int t0 = warp * 32; t0 < ncols; t0 += nwarps * 32) {
        double acc = 0.0;
        for (int b0 = 0; b0 < nblocks; b0 += 32) {
            const int b = b0 + lane;
            const double2 *p = (const double2 *)(part + (size_t)b * ldt + t0);
            double v[32];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                double2 x = (b < nblocks && t0 + 2 * q < ncols) ? __ldcg(p + q) : make_double2(0.0, 0.0);
                v[2 * q] = x.x;
                v[2 * q + 1] = (t0 + 2 * q + 1 < ncols) ? x.y : 0.0;
            }
            acc += transpose_reduce32(v, lane);
        }
        if (t0 + lane < ncols) apply(t0 + lane, acc);
    }
}

// Sum over the 32 lanes of a warp of 8 per-lane values: on return every lane holds the total of
// column (lane >> 2) & 7.   4+2+1+2 = 9 shuffles.
__device__ __forceinline__ double transpose_reduce8(double (&v)[8], int lane)
{
    // fully unrolled by hand so that v[] stays in registers
    {
        const bool up = (lane & 16) != 0;
        double s0 = up ? v[0] : v[4], s1 = up ? v[1] : v[5], s2 = up ? v[2] : v[6], s3 = up ? v[3] : v[7];
        double k0 = up ? v[4] : v[0], k1 = up ? v[5] : v[1], k2 = up ? v[6] : v[2], k3 = up ? v[7] : v[3];
        v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
        v[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
        v[2] = k2 + __shfl_xor_sync(0xffffffffu, s2, 16);
        v[3] = k3 + __shfl_xor_sync(0xffffffffu, s3, 16);
    }
    {
        const bool up = (lane & 8) != 0;
        double s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];
        double k0 = up ? v[2] : v[0], k1 = up ? v[3] : v[1];
        v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
        v[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
    }
    double x;
    {
        const bool up = (lane & 4) != 0;
        double s0 = up ? v[0] : v[1];
        double k0 = up ? v[1] : v[0];
        x = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
    }
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    return x;
}

// Geometry shared by the two row-block kernels: a block owns 32*nsub consecutive rows (sub-tiles of 32
// rows, one row per lane). Its NW*RS warps are indexed (g, h): g = column group (32 columns), h = row
// slice; slice h works on sub-tiles h, h+RS, ...  All global loads are issued in independent batches of
// 8 columns, there is no barrier inside a phase, so many sub-tiles are in flight per SM.
struct TileGeom {
    int NW, RS, nsub;
};

// ------------------------------------------------------------------------------------------------
// k_col_finish_update: finish column j-1, start column j.
//   grid = row blocks (<= PANEL_MAX_BLOCKS), block = 32*NW*RS threads
//   acol = &A[i+1, i+j] (unused when do_update == 0), S = number of GEMV partials of column j-1
//   phase A  row-wise dots: Y(r,:)s, Y(r,:)vrow, VT(r,:)s, sum of the GEMV partials
//   epilogue Y(r,j-1), VT(r,j-1), p'(r)                       (one warp per sub-tile)
//   phase B  column-wise dots w2part = VT(rows,:)^T p'(rows)  (second read of VT, from L2)
//   dynamic smem (doubles): 2j + nsub*4*NW*32 + 2*nsub*32 + RS*NW*32
// ------------------------------------------------------------------------------------------------
// The GEMV result of column j-1 arrives as `S` partial vectors yin[z*ldp + r] that are summed here in fixed
// order: the column chunks of the local GEMV (P == 1) or the inbox slots of the P ranks (P > 1, `yw` non-null:
// the block first waits until every rank has raised the flags of the 256-row blocks covering its rows).
struct YWait {
    const unsigned *flags;      // local yflag + parity*P*RB_MAX, or nullptr (no wait)
    unsigned epoch;
    int P, skip;
    unsigned *status;
};

template <int MAXT>
__global__ void __launch_bounds__(MAXT) k_col_finish_update(PanelArgs a, int j, int S, const double *__restrict__ yin,
                                                            double *__restrict__ acol, int do_update, TileGeom tg, YWait yw)
{
    SB_DYNAMIC_SMEM(double, sh);
    const int jm1 = j - 1;
    const int NW = tg.NW, RS = tg.RS, nsub = tg.nsub;
    const int nwarps = NW * RS;
    double *s_sh = sh;                             // j   (jm1 used)
    double *vrow_sh = s_sh + j;                    // j
    double *red = vrow_sh + j;                     // nsub * 4 * NW * 32
    double *pv = red + (size_t)nsub * 4 * NW * 32; // nsub * 32
    double *vtn = pv + nsub * 32;                  // nsub * 32
    double *colred = vtn + nsub * 32;              // RS * NW * 32

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = w % NW, h = w / NW;
    const int t0 = g * 32;
    const int ld = a.ld, m = a.m;
    const int row0 = blockIdx.x * nsub * 32;

    for (int t = tid; t < jm1; t += blockDim.x) s_sh[t] = a.s[t];
    if (do_update)
        for (int t = tid; t < j; t += blockDim.x) vrow_sh[t] = a.V[(size_t)t * ld + jm1];
    __syncthreads();

    // ---- phase A
    for (int sub = h; sub < nsub; sub += RS) {
        const int r = row0 + sub * 32 + lane;
        const bool valid = r < m;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
        const double *VTr = a.VT + (size_t)t0 * ld + r;
        const double *Yr = a.Y + (size_t)t0 * ld + r;
#pragma unroll
        for (int bt = 0; bt < 4; bt++) {
            const int tb = t0 + 8 * bt;
            if (tb < jm1) {
                double y8[8], v8[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const bool ok = valid && tb + q < jm1;
                    y8[q] = ok ? Yr[(size_t)(8 * bt + q) * ld] : 0.0;
                    v8[q] = ok ? VTr[(size_t)(8 * bt + q) * ld] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int t = min(tb + q, jm1 - 1);        // values beyond jm1 are zero
                    const double sv = s_sh[t];
                    d0 = fma(y8[q], sv, d0);
                    d2 = fma(v8[q], sv, d2);
                    if (do_update) d1 = fma(y8[q], vrow_sh[t], d1);
                }
            }
        }
        double *rd = red + ((size_t)sub * 4 * NW + g) * 32 + lane;
        rd[0] = d0; rd[NW * 32] = d1; rd[2 * NW * 32] = d2;
    }
    if (yw.flags != nullptr) {
        // multi-GPU: the partial results of the other ranks arrive over NVLink while the products above run
        const int rows_here = min(nsub * 32, m - row0);
        if (rows_here > 0) {
            const int rb_lo = (row0 + yw.skip) >> 8, rb_hi = (row0 + rows_here - 1 + yw.skip) >> 8;
            const int nrb = rb_hi - rb_lo + 1;
            for (int t = tid; t < yw.P * nrb; t += blockDim.x)
                wait_flag(yw.flags + (size_t)(t / nrb) * RB_MAX + rb_lo + t % nrb, yw.epoch, yw.status, 2u);
        }
        __syncthreads();
    }
    for (int sub = h; sub < nsub; sub += RS) {
        const int r = row0 + sub * 32 + lane;
        double d3 = 0.0;
        if (r < m) {
            // GEMV partials: independent loads, four accumulators
            const double *yp = yin + r;
            const size_t zs = (size_t)a.ldp * NW;
            double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
            int z = g;
            for (; z + 3 * NW < S; z += 4 * NW) {
                const double *q0 = yp + (size_t)z * a.ldp;
                e0 += __ldcg(q0); e1 += __ldcg(q0 + zs); e2 += __ldcg(q0 + 2 * zs); e3 += __ldcg(q0 + 3 * zs);
            }
            for (; z < S; z += NW) e0 += __ldcg(yp + (size_t)z * a.ldp);
            d3 = (e0 + e1) + (e2 + e3);
        }
        red[((size_t)sub * 4 * NW + 3 * NW + g) * 32 + lane] = d3;
    }
    __syncthreads();

    // ---- per-row epilogue, one warp per sub-tile
    const double tau = a.scal[jm1].tau;
    for (int sub = w; sub < nsub; sub += nwarps) {
        const int r = row0 + sub * 32 + lane;
        const bool valid = r < m;
        double vr = 0.0, ac = 0.0;
        if (valid) {
            vr = a.V[(size_t)jm1 * ld + r];
            if (do_update) ac = acol[r];
        }
        const double *rd = red + (size_t)sub * 4 * NW * 32 + lane;
        double D0 = 0.0, D1 = 0.0, D2 = 0.0, D3 = 0.0;
        for (int q = 0; q < NW; q++) {
            D0 += rd[q * 32]; D1 += rd[(NW + q) * 32]; D2 += rd[(2 * NW + q) * 32]; D3 += rd[(3 * NW + q) * 32];
        }
        double pp = 0.0, vtnew = 0.0;
        if (valid) {
            double ynew = tau * (D3 - D0);                       // finish_column: Y(:,j-1)
            a.Y[(size_t)jm1 * ld + r] = ynew;
            vtnew = tau * (vr - D2);                             // VT(:,j-1) = V * T(:,j-1)
            a.VT[(size_t)jm1 * ld + r] = vtnew;
            if (do_update) {
                pp = ac - (D1 + ynew * vrow_sh[jm1]);            // prepare_column: p - Y V(j-1,:)^T
                a.pcol[r] = pp;
            }
        }
        pv[sub * 32 + lane] = pp;
        vtn[sub * 32 + lane] = vtnew;
    }
    if (!do_update) return;
    __syncthreads();

    // ---- phase B: w2part[t] = sum over the block's rows of VT(r,t) * p'(r)
#pragma unroll
    for (int bt = 0; bt < 4; bt++) {
        const int tb = t0 + 8 * bt;
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = 0.0;
        if (tb < j) {
            for (int sub = h; sub < nsub; sub += RS) {
                const int r = row0 + sub * 32 + lane;
                const bool valid = r < m;
                const double p = pv[sub * 32 + lane], vn = vtn[sub * 32 + lane];
                const double *VTr = a.VT + (size_t)tb * ld + r;
                double v8[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v8[q] = (valid && tb + q < jm1) ? VTr[(size_t)q * ld] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = fma(tb + q == jm1 ? vn : v8[q], p, acc[q]);
            }
        }
        const double x = transpose_reduce8(acc, lane);
        if ((lane & 3) == 0) colred[(size_t)h * NW * 32 + t0 + 8 * bt + (lane >> 2)] = x;
    }
    __syncthreads();
    if (h == 0 && t0 + lane < j) {
        double sum = 0.0;
        for (int q = 0; q < RS; q++) sum += colred[(size_t)q * NW * 32 + t0 + lane];
        a.colpart[(size_t)blockIdx.x * a.ldt + t0 + lane] = sum;
    }
    if (last_block_done(a.counter, gridDim.x)) {
        double *w2 = a.w2;
        reduce_block_partials(a.colpart, a.ldt, j, gridDim.x, [w2](int t, double sum) { w2[t] = sum; });
    }
}

// ------------------------------------------------------------------------------------------------
// k_col_reflector: p'' = p' - V(:, :j) w2 ; DLARFG scalars ; s = V^T v.
//   grid = row blocks, block = 32*NW*RS threads; acol = &A[i+1, i+j]; for j == 0 the column is read from
//   acol directly.   dynamic smem (doubles): j + nsub*NW*32 + nsub*32 + RS*NW*32 + 32
// DLARFG (LAPACK, called at src/hessenberg/cpu.c:140): beta = -sign(alpha) * hypot(alpha, ||x||),
// tau = (beta - alpha) / beta, v = x / (alpha - beta); ||x|| == 0 (or an empty x) gives tau = 0.
// With v = (1, scale * x):  s = V(j:, :j)^T v = V(j, :j)^T + scale * V(j+1:, :j)^T x.
// ------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT) k_col_reflector(PanelArgs a, int j, double *__restrict__ acol, TileGeom tg)
{
    SB_DYNAMIC_SMEM(double, sh);
    const int NW = tg.NW, RS = tg.RS, nsub = tg.nsub;
    const int nwarps = NW * RS;
    double *w2_sh = sh;                             // j
    double *red = w2_sh + j;                        // nsub * NW * 32
    double *pv = red + (size_t)nsub * NW * 32;      // nsub * 32
    double *colred = pv + nsub * 32;                // RS * NW * 32
    double *sqred = colred + (size_t)RS * NW * 32;  // 3 x 32 (one triple per warp)
    __shared__ double scale_sh, xmul_sh;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = w % NW, h = w / NW;
    const int t0 = g * 32;
    const int ld = a.ld, m = a.m;
    const int row0 = blockIdx.x * nsub * 32;

    for (int t = tid; t < j; t += blockDim.x) w2_sh[t] = a.w2[t];
    __syncthreads();

    // ---- phase A: d(r) = V(r, :j) w2
    for (int sub = h; sub < nsub; sub += RS) {
        const int r = row0 + sub * 32 + lane;
        const bool valid = r < m;
        double d = 0.0;
        const double *Vr = a.V + (size_t)t0 * ld + r;
#pragma unroll
        for (int bt = 0; bt < 4; bt++) {
            const int tb = t0 + 8 * bt;
            if (tb < j) {
                double v8[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v8[q] = (valid && tb + q < j) ? Vr[(size_t)(8 * bt + q) * ld] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) d = fma(v8[q], w2_sh[min(tb + q, j - 1)], d);
            }
        }
        red[((size_t)sub * NW + g) * 32 + lane] = d;
    }
    __syncthreads();

    // ---- per-row epilogue
    SumSq sq;
    sq.clear();
    for (int sub = w; sub < nsub; sub += nwarps) {
        const int r = row0 + sub * 32 + lane;
        const bool valid = r < m;
        double base = 0.0;
        if (valid) base = j > 0 ? a.pcol[r] : acol[r];
        double D = 0.0;
        for (int q = 0; q < NW; q++) D += red[((size_t)sub * NW + q) * 32 + lane];
        double x = 0.0;
        if (valid) {
            double pp = base - D;
            a.pcol[r] = pp;
            if (r < j) acol[r] = pp;            // final entries of H above the sub-diagonal
            if (r == j) a.scal[j].alpha = pp;
            if (r > j) x = pp;
        }
        sq.add(x);
        pv[sub * 32 + lane] = x;
    }
    sq.med = warp_sum(sq.med);
    if (__any_sync(0xffffffffu, sq.big != 0.0 || sq.sml != 0.0)) { sq.big = warp_sum(sq.big); sq.sml = warp_sum(sq.sml); }
    if (lane == 0) { sqred[w] = sq.med; sqred[32 + w] = sq.big; sqred[64 + w] = sq.sml; }
    __syncthreads();

    // ---- phase B: zpart[t] = sum over the block's rows > j of V(r,t) * p''(r)
    if (j > 0) {
#pragma unroll
        for (int bt = 0; bt < 4; bt++) {
            const int tb = t0 + 8 * bt;
            double acc[8];
#pragma unroll
            for (int q = 0; q < 8; q++) acc[q] = 0.0;
            if (tb < j) {
                for (int sub = h; sub < nsub; sub += RS) {
                    const int r = row0 + sub * 32 + lane;
                    const bool valid = r < m;
                    const double p = pv[sub * 32 + lane];
                    const double *Vr = a.V + (size_t)tb * ld + r;
                    double v8[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) v8[q] = (valid && tb + q < j) ? Vr[(size_t)q * ld] : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma(v8[q], p, acc[q]);
                }
            }
            const double x = transpose_reduce8(acc, lane);
            if ((lane & 3) == 0) colred[(size_t)h * NW * 32 + t0 + 8 * bt + (lane >> 2)] = x;
        }
        __syncthreads();
        if (h == 0 && t0 + lane < j) {
            double sum = 0.0;
            for (int q = 0; q < RS; q++) sum += colred[(size_t)q * NW * 32 + t0 + lane];
            a.colpart[(size_t)blockIdx.x * a.ldt + t0 + lane] = sum;
        }
    }
    if (tid < 3) {
        double sum = 0.0;
        for (int q = 0; q < nwarps; q++) sum += sqred[32 * tid + q];
        a.sqpart[tid * PANEL_LDB + blockIdx.x] = sum;
    }
    if (last_block_done(a.counter, gridDim.x)) {
        if (w == 0) {
            double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int q = 0; q < PANEL_LDB / 32; q++) {
                    int b = lane + 32 * q;
                    acc[k] += b < (int)gridDim.x ? __ldcg(a.sqpart + k * PANEL_LDB + b) : 0.0;
                }
#pragma unroll
            for (int k = 0; k < 3; k++) acc[k] = warp_sum(acc[k]);
            if (lane == 0) {
                const double alpha = __ldcg(&a.scal[j].alpha);
                const Reflector f = dlarfg_scalars(alpha, acc[0], acc[1], acc[2], m - j > 1);
                a.scal[j].tau = f.tau;
                a.scal[j].beta = f.beta;
                a.scal[j].scale = f.scale;
                scale_sh = f.scale;
                xmul_sh = f.xmul;
            }
        }
        __syncthreads();
        if (xmul_sh != 1.0) {
            // LAPACK's rescaling branch (denormal-range column), all of it by this one block: x is multiplied by 2^969
            // (the GEMV kernel that forms v = x * scale is the next launch on the stream) and s = V^T v is taken
            // directly from the rescaled x -- the per-block partials of z were summed in denormal arithmetic
            const double xmul = xmul_sh, scale = scale_sh;
            for (int r = j + 1 + tid; r < m; r += blockDim.x) a.pcol[r] *= xmul;
            __syncthreads();
            for (int t = w; t < j; t += nwarps) {
                double acc = 0.0;
                for (int r = j + 1 + lane; r < m; r += 32) acc = fma(a.V[(size_t)t * ld + r], a.pcol[r], acc);
                acc = warp_sum(acc);
                if (lane == 0) a.s[t] = fma(scale, acc, a.V[(size_t)t * ld + j]);
            }
            return;
        }
        const double scale = scale_sh;
        double *s = a.s;
        const double *Vrow = a.V + j;
        reduce_block_partials(a.colpart, a.ldt, j, gridDim.x,
                              [s, Vrow, scale, ld](int t, double sum) { s[t] = fma(scale, sum, Vrow[(size_t)t * ld]); });
    }
}

// ------------------------------------------------------------------------------------------------
// k_col_gemv: ypart[z][r] = sum_{local columns lc in chunk z} A_loc[i+1+r, lc] * v[l2g(lc) - (c+1)],
//             v[0] = 1, v[k] = p''[j+k]*scale      (c = i+j the column being reduced)
//
//   ncols = m - j: length of v;  the rank's columns of the global range [c+1, e) are the local range
//           [lc0, lc0+nloc) (ColMap; the whole range when P == 1)
//   A0    16-byte aligned pointer to A_loc[i+1-skip, lc0] (skip in {0,1}); padded row rp = r + skip
//   grid  = RB*S blocks of 128 threads: block b: row block b % RB (256 padded rows), column chunk b / RB
//   Every block also writes a slice of V(j+k, j) = v[k] and of the panel column (beta, exact zeros).
//   DIST: the block that finishes a row block last sums the S partials (fixed order) and pushes the 256 rows
//   to every rank's inbox over NVLink, then raises that rank's flag (see Xchg).
// Memory-bound: each thread streams one 16-byte load per column with 8 columns in flight.
// ------------------------------------------------------------------------------------------------
template <bool DIST>
__global__ void __launch_bounds__(GEMV_THREADS, 10) k_col_gemv(PanelArgs a, int j, int ncols, ColMap cm, int lc0, int nloc,
                                                                int gc0, const double *__restrict__ A0, int lda, int skip,
                                                                int kc, int RB, int S, double *__restrict__ acol, Xchg x)
{
    SB_DYNAMIC_SMEM(double, vs);        // kc
    const int tid = threadIdx.x;
    const int m = a.m;                  // ncols == m - j inside a reduction; free in the unit test (j == 0)
    const double scale = a.scal[j].scale;

    const int rb = blockIdx.x % RB, z = blockIdx.x / RB;
    const int k0 = z * kc;
    const int nk = max(0, min(kc, nloc - k0));
    // v for the chunk's columns: global column gc <-> v index gc - gc0  (gc0 = c+1)
    for (int k = tid; k < nk; k += GEMV_THREADS) {
        const int kk = cm.l2g(lc0 + k0 + k) - gc0;
        vs[k] = (kk == 0) ? 1.0 : a.pcol[j + kk] * scale;
    }
    {   // this block's slice of V(:, j) and of the reduced column
        const int per = (ncols + gridDim.x - 1) / gridDim.x;
        const int kb = blockIdx.x * per, ke = min(ncols, kb + per);
        for (int k = kb + tid; k < ke; k += GEMV_THREADS) {
            a.V[(size_t)j * a.ld + j + k] = (k == 0) ? 1.0 : a.pcol[j + k] * scale;
            acol[j + k] = (k == 0) ? a.scal[j].beta : 0.0;
        }
    }
    __syncthreads();

    const int mp = m + skip;
    const int rp = rb * 256 + tid * 2;           // padded row of .x ; rows rp, rp+1
    const int r = rp - skip;                     // logical row of .x
    double *yp = a.ypart + (size_t)z * a.ldp;
    if (rp < mp) {
        double2 acc = make_double2(0.0, 0.0);
        const double *Ap = A0 + (size_t)k0 * lda + rp;
        // software pipeline: U loads of the next column group are in flight while the current
        // group is accumulated (2*U 16-byte loads per thread outstanding)
        constexpr int U = 4;
        const size_t step = (size_t)lda;
        double2 cur[U], nxt[U];
        int k = 0;
        if (nk >= U) {
#pragma unroll
            for (int u = 0; u < U; u++) cur[u] = __ldcs((const double2 *)(Ap + u * step));
            const double *Pn = Ap + U * step;
            for (; k + 2 * U <= nk; k += U) {
#pragma unroll
                for (int u = 0; u < U; u++) nxt[u] = __ldcs((const double2 *)(Pn + u * step));
                Pn += U * step;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    double vk = vs[k + u];
                    acc.x = fma(cur[u].x, vk, acc.x);
                    acc.y = fma(cur[u].y, vk, acc.y);
                }
#pragma unroll
                for (int u = 0; u < U; u++) cur[u] = nxt[u];
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                double vk = vs[k + u];
                acc.x = fma(cur[u].x, vk, acc.x);
                acc.y = fma(cur[u].y, vk, acc.y);
            }
            k += U;
        }
        for (; k < nk; k++) {
            double vk = vs[k];
            double2 xx = __ldcs((const double2 *)(Ap + (size_t)k * step));
            acc.x = fma(xx.x, vk, acc.x);
            acc.y = fma(xx.y, vk, acc.y);
        }
        if (r >= 0) yp[r] = acc.x;
        if (r + 1 < m) yp[r + 1] = acc.y;
    }
    if (!DIST) return;

    // ---- multi-GPU tail: last chunk of this row block reduces and pushes
    __shared__ bool last_sh;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned ticket = atomicAdd(x.rbcount + rb, 1u);
        last_sh = (ticket == (unsigned)S - 1);
        if (last_sh) x.rbcount[rb] = 0;      // re-arm (stream ordered)
    }
    __syncthreads();
    if (!last_sh) return;
    __threadfence();
    const int par = x.epoch & 1;
    if (rp < mp) {
        // fixed summation order; loads are issued in independent batches of 8 (clamped addresses, no branches)
        const bool ok0 = r >= 0, ok1 = r + 1 < m;
        const double *q0 = a.ypart + (ok0 ? r : r + 1), *q1 = a.ypart + (ok1 ? r + 1 : r);
        double s0 = 0.0, s1 = 0.0;
        int zz = 0;
        for (; zz + 8 <= S; zz += 8) {
            double x0[8], x1[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { x0[u] = __ldcg(q0 + (size_t)(zz + u) * a.ldp); x1[u] = __ldcg(q1 + (size_t)(zz + u) * a.ldp); }
#pragma unroll
            for (int u = 0; u < 8; u++) { s0 += x0[u]; s1 += x1[u]; }
        }
        for (; zz < S; zz++) { s0 += __ldcg(q0 + (size_t)zz * a.ldp); s1 += __ldcg(q1 + (size_t)zz * a.ldp); }
        const size_t slot = ((size_t)par * x.P + x.g) * a.ldp;
#pragma unroll 1
        for (int d = 0; d < x.P; d++) {
            double *dst = x.inbox[(x.g + d) % x.P] + slot + r;      // start with the own inbox, then the neighbours
            if (ok0) dst[0] = s0;
            if (ok1) dst[1] = s1;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid < x.P) st_release_sys(x.yflag[tid] + ((size_t)par * x.P + x.g) * RB_MAX + rb, x.epoch);
}

// ------------------------------------------------------------------------------------------------
// multi-GPU helpers
// ------------------------------------------------------------------------------------------------
struct PeerPtrs { double *p[MAX_RANKS]; };

// Panel gather: every rank copies its columns of the global range [i, i+w) (rows i+1 .. i+m) into the panel
// buffer Pan (m x w, leading dimension ldv) of EVERY rank. grid = (row chunks of 1024, local panel columns)
__global__ void k_panel_push(ColMap cm, int i, int lc_first, int m, const double *__restrict__ Aloc, int lda, PeerPtrs pan, int ldv)
{
    const int lc = lc_first + blockIdx.y;
    const int col = cm.l2g(lc) - i;
    const double *src = Aloc + (size_t)lc * lda + i + 1;
    const int r0 = blockIdx.x * 1024;
    for (int d = 0; d < cm.P; d++) {
        double *dst = pan.p[(cm.g + d) % cm.P] + (size_t)col * ldv;
        for (int r = r0 + threadIdx.x; r < min(m, r0 + 1024); r += blockDim.x) dst[r] = src[r];
    }
}

// Panel write-back: the reduced panel columns (H entries, exact zeros) return to their owner's local storage
__global__ void k_panel_pull(ColMap cm, int i, int lc_first, int m, double *__restrict__ Aloc, int lda,
                             const double *__restrict__ pan, int ldv)
{
    const int lc = lc_first + blockIdx.y;
    const int col = cm.l2g(lc) - i;
    double *dst = Aloc + (size_t)lc * lda + i + 1;
    const double *src = pan + (size_t)col * ldv;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) dst[r] = src[r];
}

// Vg(x, t) = V(l2g(lc_first + x) - row0g, t): the rows of V / VT that belong to the rank's local columns
__global__ void k_gather_rows(ColMap cm, int lc_first, int ncl, int row0g, int w, const double *__restrict__ V,
                              const double *__restrict__ VT, int ld, double *__restrict__ Vg, double *__restrict__ VTg, int ldg)
{
    const int xx = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (xx >= ncl) return;
    const int row = cm.l2g(lc_first + xx) - row0g;
    Vg[(size_t)t * ldg + xx] = V[(size_t)t * ld + row];
    VTg[(size_t)t * ldg + xx] = VT[(size_t)t * ld + row];
}

// out(r, t) = sum over ranks s (fixed order) of part_s(r, t): the all-reduce of the top-row products,
// pulled from every rank's exchange buffer over NVLink
__global__ void k_sum_peers(int P, int rows, PeerPtrs part, int ldp, double *__restrict__ out, int ldo)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (r >= rows) return;
    double s = 0.0;
    for (int q = 0; q < P; q++) s += __ldcg(part.p[q] + (size_t)t * ldp + r);
    out[(size_t)t * ldo + r] = s;
}

// All-rank barrier on device flags: rank g writes `epoch` into slot g of every rank's bar[] and waits until
// all P slots of its own bar[] carry it. One block of MAX_RANKS threads. Writes of earlier kernels of this
// stream are ordered before the flag by the fence; readers acquire.
struct BarPtrs { unsigned *p[MAX_RANKS]; };
__global__ void k_barrier(int P, int g, unsigned epoch, BarPtrs bar, unsigned *status)
{
    const int t = threadIdx.x;
    __threadfence_system();
    if (t < P) st_release_sys(bar.p[t] + g, epoch);
    if (t < P) wait_flag(bar.p[g] + t, epoch, status);
    __syncthreads();
    __threadfence_system();
}

// flag |= 1 unless rows [q0, q0 + rows) of the n-column matrix Q (leading dimension ld; Q points at row q0) are exactly those
// rows of the identity (-0.0 counts as zero, a NaN as a mismatch): decides between forward and backward accumulation of Q
// (engine.cuh, Rank::reduce). One block per column, grid-stride.
__global__ void __launch_bounds__(256) k_is_identity(int rows, int q0, int n, const double *__restrict__ Q, int ld, unsigned *flag)
{
    bool bad = false;
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        const double *q = Q + (size_t)c * ld;
        for (int r = threadIdx.x; r < rows; r += blockDim.x) bad = bad || !(q[r] == (r + q0 == c ? 1.0 : 0.0));
    }
    if (bad) atomicExch(flag, 1u);
}

// the rank's contribution to the sum over ranks that decides the accumulation order of Q: 0.0 = "identity slab, room for the history"
__global__ void k_flag_to_double(const unsigned *flag, double *dst)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = *flag != 0u ? 1.0 : 0.0;
}

// the rank's columns of the n x n identity: Qc(r, lc) = (r == global index of local column lc).
// grid = (row chunks of 1024, local columns, grid-stride)
__global__ void __launch_bounds__(256) k_identity_cols(ColMap cm, int n, int nloc, double *__restrict__ Qc, int ldc)
{
    const int r_end = min(n, (int)(blockIdx.x + 1) * 1024);
    for (int lc = blockIdx.y; lc < nloc; lc += gridDim.y) {
        const int c = cm.l2g(lc);
        double *q = Qc + (size_t)lc * ldc;
        for (int r = blockIdx.x * 1024 + threadIdx.x; r < r_end; r += 256) q[r] = r == c ? 1.0 : 0.0;
    }
}

// Column blocks -> row slabs: Q(r - q0, c) = Qc_owner(c)(r, local index of c) for the rank's rows r in [q0, q0 + rows), pulled
// from the owner's exchange arena over NVLink (a column is contiguous there). grid = (row chunks of 256, columns, grid-stride)
__global__ void __launch_bounds__(256) k_qcols_to_rows(ColMap cm, int n, int q0, int rows, PeerPtrs qc, int ldc, double *__restrict__ Q, int ldq)
{
    const int rr = blockIdx.x * 256 + threadIdx.x;
    if (rr >= rows) return;
    for (int c = blockIdx.y; c < n; c += gridDim.y) {
        const int lc = (c / (cm.cb * cm.P)) * cm.cb + c % cm.cb;
        Q[(size_t)c * ldq + rr] = __ldcg(qc.p[cm.owner(c)] + (size_t)lc * ldc + q0 + rr);
    }
}

} // namespace sb200
