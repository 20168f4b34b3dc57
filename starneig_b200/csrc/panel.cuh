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
//   k_col_finish_update  (j >= 1) row-parallel: finishes column j-1 (Y, VT) and applies the right
//                        update p' = p - Y V(j-1,:)^T; reduces  w2 = VT^T p'   (last block sums)
//   k_col_reflector      row-parallel: p'' = p' - V w2; reduces ||p''(j+1:)||^2; the last block does
//                        the DLARFG scalar work (beta, tau, 1/(alpha-beta))
//   k_col_gemv           the HBM-bound GEMV over the trailing matrix with v formed on the fly from
//                        p'' and the scale; also writes V(:,j), the exact zeros and beta into A, and
//                        reduces s = V^T v on a few extra blocks
// Cross-block reductions write per-block partials; the block that finishes last adds them in a fixed
// order, so results are bitwise reproducible run to run.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int PR = 64;              // rows per block in the row-parallel panel kernels
constexpr int PG = 8;               // column groups (threads per row)
constexpr int PT = PR * PG;         // 512 threads
constexpr int GEMV_THREADS = 128;
constexpr int GEMV_SROWS = 256;     // rows per s-block in k_col_gemv

struct ColScal {        // DLARFG results for one column
    double tau, beta, scale, alpha;
};

struct PanelArgs {
    int m;              // rows of the panel (= end - i - 1); local row r <-> global row i+1+r
    int ld;             // leading dimension of V, Y, VT
    double *V, *Y, *VT;
    double *pcol;       // m   scratch: the column being reduced
    double *ypart;      // S x ldp  GEMV partial sums
    int ldp;
    double *s;          // nb  s = V^T v
    double *w2;         // nb
    double *w2part;     // blocks x ldw
    double *spart;      // blocks x ldw
    int ldw;
    double *sqpart;     // blocks
    ColScal *scal;      // nb
    unsigned *counter;  // zero-initialised
};

// ------------------------------------------------------------------------------------------------
// k_col_finish_update: finish column j-1, start column j.   grid = ceil(m / PR), block = PT
//   acol = &A[i+1, i+j] (unused when do_update == 0), S = number of GEMV partials of column j-1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) k_col_finish_update(PanelArgs a, int j, int S, double *__restrict__ acol, int do_update)
{
    extern __shared__ double sh[];
    const int jm1 = j - 1;
    double *s_sh = sh;                         // jm1
    double *vrow_sh = s_sh + jm1;              // j
    double *red = vrow_sh + j;                 // 4 * PG * PR
    double *pv = red + 4 * PG * PR;            // PR
    double *vtn = pv + PR;                     // PR

    const int tid = threadIdx.x, lr = tid % PR, g = tid / PR;
    const int r0 = blockIdx.x * PR, r = r0 + lr;
    const bool valid = r < a.m;
    const int ld = a.ld;

    for (int t = tid; t < jm1; t += PT) s_sh[t] = a.s[t];
    if (do_update)
        for (int t = tid; t < j; t += PT) vrow_sh[t] = a.V[(size_t)t * ld + jm1];
    __syncthreads();

    const double tau = a.scal[jm1].tau;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    if (valid) {
        const double *Yr = a.Y + r, *VTr = a.VT + r;
        if (do_update) {
#pragma unroll 4
            for (int t = g; t < jm1; t += PG) {
                double yv = Yr[(size_t)t * ld], vt = VTr[(size_t)t * ld], sv = s_sh[t];
                d0 = fma(yv, sv, d0);
                d1 = fma(yv, vrow_sh[t], d1);
                d2 = fma(vt, sv, d2);
            }
        } else {
#pragma unroll 4
            for (int t = g; t < jm1; t += PG) {
                double yv = Yr[(size_t)t * ld], vt = VTr[(size_t)t * ld], sv = s_sh[t];
                d0 = fma(yv, sv, d0);
                d2 = fma(vt, sv, d2);
            }
        }
        for (int z = g; z < S; z += PG) d3 += a.ypart[(size_t)z * a.ldp + r];
    }
    red[(0 * PG + g) * PR + lr] = d0;
    red[(1 * PG + g) * PR + lr] = d1;
    red[(2 * PG + g) * PR + lr] = d2;
    red[(3 * PG + g) * PR + lr] = d3;
    __syncthreads();
    if (g == 0) {
        double D0 = 0.0, D1 = 0.0, D2 = 0.0, D3 = 0.0;
#pragma unroll
        for (int q = 0; q < PG; q++) {
            D0 += red[(0 * PG + q) * PR + lr];
            D1 += red[(1 * PG + q) * PR + lr];
            D2 += red[(2 * PG + q) * PR + lr];
            D3 += red[(3 * PG + q) * PR + lr];
        }
        double pp = 0.0, vtnew = 0.0;
        if (valid) {
            double ynew = tau * (D3 - D0);                       // finish_column: Y(:,j-1)
            a.Y[(size_t)jm1 * ld + r] = ynew;
            double vr = a.V[(size_t)jm1 * ld + r];
            vtnew = tau * (vr - D2);                             // VT(:,j-1) = V * T(:,j-1)
            a.VT[(size_t)jm1 * ld + r] = vtnew;
            if (do_update) {
                pp = acol[r] - (D1 + ynew * vrow_sh[jm1]);       // prepare_column: p - Y V(j-1,:)^T
                a.pcol[r] = pp;
            }
        }
        pv[lr] = pp;
        vtn[lr] = vtnew;
    }
    if (!do_update) return;
    __syncthreads();

    // w2part[t] = sum over this block's rows of VT(r,t) * p'(r), one warp per column t
    const int warp = tid >> 5, lane = tid & 31;
    double *out = a.w2part + (size_t)blockIdx.x * a.ldw;
    for (int t = warp; t < j; t += PT / 32) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < PR / 32; k++) {
            int l2 = lane + 32 * k, rr = r0 + l2;
            if (rr < a.m) {
                double vt = (t == jm1) ? vtn[l2] : a.VT[(size_t)t * ld + rr];
                acc = fma(vt, pv[l2], acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) out[t] = acc;
    }
    if (last_block_done(a.counter, gridDim.x)) {
        for (int t = tid; t < j; t += PT) {
            double sum = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++) sum += __ldcg(a.w2part + (size_t)b * a.ldw + t);
            a.w2[t] = sum;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_col_reflector: p'' = p' - V(:, :j) w2 ; DLARFG scalars.     grid = ceil(m / PR), block = PT
//   acol = &A[i+1, i+j]; for j == 0 the column is read from acol directly.
// DLARFG (LAPACK, called at src/hessenberg/cpu.c:140): beta = -sign(alpha) * hypot(alpha, ||x||),
// tau = (beta - alpha) / beta, v = x / (alpha - beta); ||x|| == 0 (or an empty x) gives tau = 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) k_col_reflector(PanelArgs a, int j, double *__restrict__ acol)
{
    extern __shared__ double sh[];
    double *w2_sh = sh;                 // j
    double *red = w2_sh + j;            // PG * PR
    __shared__ double wsum[PR / 32];

    const int tid = threadIdx.x, lr = tid % PR, g = tid / PR;
    const int r = blockIdx.x * PR + lr;
    const bool valid = r < a.m;
    const int ld = a.ld;

    for (int t = tid; t < j; t += PT) w2_sh[t] = a.w2[t];
    __syncthreads();
    double d = 0.0;
    if (valid) {
        const double *Vr = a.V + r;
#pragma unroll 4
        for (int t = g; t < j; t += PG) d = fma(Vr[(size_t)t * ld], w2_sh[t], d);
    }
    red[g * PR + lr] = d;
    __syncthreads();
    if (g == 0) {       // threads 0..PR-1 = the first PR/32 warps
        double D = 0.0;
#pragma unroll
        for (int q = 0; q < PG; q++) D += red[q * PR + lr];
        double sq = 0.0;
        if (valid) {
            double pp = (j > 0 ? a.pcol[r] : acol[r]) - D;
            a.pcol[r] = pp;
            if (r < j) acol[r] = pp;            // final entries of H above the sub-diagonal
            if (r == j) a.scal[j].alpha = pp;
            if (r > j) sq = pp * pp;
        }
        sq = warp_sum(sq);
        if ((tid & 31) == 0) wsum[tid >> 5] = sq;
    }
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < PR / 32; q++) sum += wsum[q];
        a.sqpart[blockIdx.x] = sum;
    }
    if (last_block_done(a.counter, gridDim.x)) {
        if (tid == 0) {
            double ssq = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++) ssq += __ldcg(a.sqpart + b);
            double alpha = __ldcg(&a.scal[j].alpha);
            double xnorm = sqrt(ssq);
            double tau = 0.0, beta = alpha, scale = 0.0;
            if (a.m - j > 1 && xnorm != 0.0) {
                beta = -copysign(hypot(alpha, xnorm), alpha);
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            a.scal[j].tau = tau;
            a.scal[j].beta = beta;
            a.scal[j].scale = scale;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_col_gemv: ypart[z][r] = sum_{k in chunk z} A[i+1+r, c+1+k] * v[k],  v[0] = 1, v[k] = p''[j+k]*scale
//
//   A0    16-byte aligned pointer to A[i+1-skip, c+1] (skip in {0,1}); padded row rp = r + skip
//   ncols = number of columns (m - j)
//   grid  = nsb + RB*S blocks of 128 threads: the first nsb blocks reduce s = V(j:, :j)^T v over
//           GEMV_SROWS rows each; GEMV block b: row block b % RB (256 padded rows), chunk b / RB
//   The first row block of every chunk also stores V(j+k, j) = v[k], A[i+1+j+k, c] = (k==0 ? beta : 0).
// Memory-bound: each thread streams one 16-byte load per column with 8 columns in flight.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMV_THREADS) k_col_gemv(PanelArgs a, int j, int ncols, const double *__restrict__ A0,
                                                            int lda, int skip, int kc, int RB, int S, int nsb,
                                                            double *__restrict__ acol)
{
    extern __shared__ double vs[];      // max(kc, GEMV_SROWS)
    const int tid = threadIdx.x;
    const int m = a.m;                  // ncols == m - j inside a reduction; free in the unit test (j == 0)
    const double scale = a.scal[j].scale;
    const unsigned total_blocks = gridDim.x;

    if ((int)blockIdx.x < nsb) {
        // ---- s-block: rows j + b*GEMV_SROWS ...
        const int k0 = blockIdx.x * GEMV_SROWS;
        const int nk = min(GEMV_SROWS, ncols - k0);
        for (int k = tid; k < GEMV_SROWS; k += GEMV_THREADS)
            vs[k] = k < nk ? ((k0 + k == 0) ? 1.0 : a.pcol[j + k0 + k] * scale) : 0.0;
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        double *out = a.spart + (size_t)blockIdx.x * a.ldw;
        for (int t = warp; t < j; t += GEMV_THREADS / 32) {
            const double *Vt = a.V + (size_t)t * a.ld + j + k0;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < GEMV_SROWS / 32; q++) {
                int k = lane + 32 * q;
                if (k < nk) acc = fma(Vt[k], vs[k], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) out[t] = acc;
        }
    } else {
        // ---- GEMV block
        const int b = blockIdx.x - nsb;
        const int rb = b % RB, z = b / RB;
        const int k0 = z * kc;
        const int nk = min(kc, ncols - k0);
        for (int k = tid; k < nk; k += GEMV_THREADS) {
            double v = (k0 + k == 0) ? 1.0 : a.pcol[j + k0 + k] * scale;
            vs[k] = v;
            if (rb == 0) {
                a.V[(size_t)j * a.ld + j + k0 + k] = v;
                acol[j + k0 + k] = (k0 + k == 0) ? a.scal[j].beta : 0.0;
            }
        }
        __syncthreads();

        const int mp = m + skip;
        const int rp = rb * 256 + tid * 2;           // padded row of .x ; rows rp, rp+1
        double2 acc = make_double2(0.0, 0.0);
        if (rp < mp) {
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
                double2 x = __ldcs((const double2 *)(Ap + (size_t)k * step));
                acc.x = fma(x.x, vk, acc.x);
                acc.y = fma(x.y, vk, acc.y);
            }
            double *yp = a.ypart + (size_t)z * a.ldp;
            int r = rp - skip;                       // logical row of .x
            if (r >= 0) yp[r] = acc.x;
            if (r + 1 < m) yp[r + 1] = acc.y;
        }
    }

    if (nsb > 0 && last_block_done(a.counter, total_blocks)) {
        for (int t = tid; t < j; t += GEMV_THREADS) {
            double sum = 0.0;
            for (int b = 0; b < nsb; b++) sum += __ldcg(a.spart + (size_t)b * a.ldw + t);
            a.s[t] = sum;
        }
    }
}

} // namespace sb200
