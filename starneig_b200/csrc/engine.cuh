// engine.cuh -- per-rank engine of the B200-native blocked Hessenberg reduction.
//
// Replaces, for the path behind starneig_SEP_SM_Hessenberg (reference src/hessenberg/interface.c),
// the StarPU task graph of src/hessenberg/core.c:351-599 and the tile plumbing of src/common
// (matrix.c, vector.c, tiles.c, scratch.c): the matrix stays dense and column-major in HBM, the
// "task graph" is a fixed sequence of kernel launches on one CUDA stream per GPU, and all workspace
// comes from two arenas owned by the rank (a private one and a peer-visible exchange arena).
//
// One `Rank` drives one GPU. P ranks (threads of one process, or one process each) run the same code:
//   * A is 1-D block-cyclic by columns (ColMap, block width cb), every rank holds full-height columns;
//     Q is split by rows (its update is a right-multiplication, so row slabs need no communication).
//   * The panel (w columns) is gathered into a replicated buffer Pan on every rank; the level-2 panel
//     kernels run redundantly (bitwise identical) on all ranks, so V, Y, VT, tau never travel.
//   * The one per-column exchange is the sum of the GEMV partials: pushed over NVLink peer stores by the
//     GEMV kernel itself and consumed by the next k_col_finish_update (panel.cuh, struct Xchg).
//   * Per panel: trailing right/left updates touch local columns only; the rows above the panel need a
//     sum over ranks of (i+1) x w products, pulled from the peers' exchange arenas (k_sum_peers).
//
// Panel i (columns i .. i+w-1, m = end-i-1 rows below the diagonal), cf. SURVEY.md section 8a:
//   column loop          k_col_finish_update / k_col_reflector / k_col_gemv     (panel.cuh)
//   A(i+1:e, i+w:e) -= Y V(w-1:,:)^T                       core.c:523-540, cpu.c:315
//   A(i+1:e, i+w:e) -= V (A^T VT)^T                        core.c:546-547, cpu.c:373-435
//   A(0:i+1, i+1:e) -= (A VT) V^T                          core.c:320-327, cpu.c:492-554
//   A(i+1:e, e:n)   -= V (A^T VT)^T   (partial only)       core.c:329-336
//   Q(:, i+1:e)     -= (Q VT) V^T                          core.c:338-340
// with VT = V*T (see panel.cuh). The reference defers the last three to the end of the graph at lower
// priority; they only depend on this panel's V and VT and touch disjoint data, so issuing them right
// after the panel is the same computation.
#pragma once
#include <chrono>
#include "panel.cuh"
#include "panel_fused.cuh"
#include "dgemm.cuh"
#include "dgemm_tma.cuh"
#include <starneig_b200.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// GEMM dispatch
// ---------------------------------------------------------------------------------------------
// <A K-major, B K-major, warps along M, warps along N, 8-row blocks per warp, 8-col blocks per warp, stages, CTAs/SM, OPT>
// Defaults from the tile sweep against cuBLAS on the exact shapes (tools/gemm_sweep.cu, profiles/r2_v1_gemm_sweep_p2.txt;
// cuBLAS: NT 33.2, TN 34.8, NN 34.0 TFLOP/s): NT with the next stage's cp.async spread between the DMMAs (OPT bit 0:
// 28.2 -> 31.6 TFLOP/s); the skinny products with a 3-stage ring instead of 4 (TN 22.5 -> 30.7, NN 30.9 -> 32.2).
using GemmNT   = GemmConfig<false, false, 2, 2, 8, 4, 4, 2, 1>;  // 128 x  64, 128 threads: rank-nb updates
using GemmTN13 = GemmConfig<true,  true,  4, 1, 2, 13, 3, 2>;    //  64 x 104, W = A^T VT
using GemmTN12 = GemmConfig<true,  true,  4, 1, 2, 12, 3, 2>;    //  64 x  96
using GemmNN13 = GemmConfig<false, true,  4, 1, 2, 13, 3, 2>;    //  64 x 104, W = A VT
using GemmNN12 = GemmConfig<false, true,  4, 1, 2, 12, 3, 2>;    //  64 x  96
// Tiles fed by the TMA engine (dgemm_tma.cuh: one thread issues bulk tensor copies into 128-byte-swizzled shared memory,
// the warps run LDS + DMMA only). The defaults wherever a product can be framed for TMA (even leading dimensions; odd
// operand offsets are handled by moving the frame); STARNEIG_B200_GEMM_TMA=0 selects the cp.async kernels. Tiles from
// the sweep on the engine's shapes and operand parities (profiles/r2_v3_gemm_sweep_tma_p2.txt), TFLOP/s TMA / cuBLAS:
// NT 64 x 64, 3 stages, 4 CTAs/SM: 33.2 / 33.2; TN 64 x 104, 3 stages: 34.6 / 34.7; NN 64 x 104, 3 stages: 34.4 / 34.0.
using TmaNT   = GemmTmaConfig<false, false, 2, 2, 4, 4, 3, 4>;
using TmaTN13 = GemmTmaConfig<true,  true,  4, 1, 2, 13, 3, 2>;
using TmaTN12 = GemmTmaConfig<true,  true,  4, 1, 2, 12, 3, 2>;
using TmaNN13 = GemmTmaConfig<false, true,  4, 1, 2, 13, 3, 2>;
using TmaNN12 = GemmTmaConfig<false, true,  4, 1, 2, 12, 3, 2>;

static const size_t PANEL_SMEM_MAX = 200 * 1024;
// The persistent panel kernel keeps the CTA's rows of V in shared memory only while the launch stays below this: the GEMV
// needs the rest of the SM's 256 KB as L1 for its loads in flight (FusedSmem; profiles/r2_v15_sweep_l1_split.txt)
static const size_t FUSED_SLAB_SMEM_MAX = 131 * 1024;
static const size_t FUSED_SMEM_OPTIN = 226 * 1024;

// per-device function attributes (opt-in shared memory sizes)
static void prepare_device_functions()
{
    GemmNT::prepare(); GemmTN13::prepare(); GemmTN12::prepare(); GemmNN13::prepare(); GemmNN12::prepare();
    TmaNT::prepare(); TmaTN13::prepare(); TmaTN12::prepare(); TmaNN13::prepare(); TmaNN12::prepare();
    SB_CUDA(cudaFuncSetAttribute(k_col_finish_update<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    SB_CUDA(cudaFuncSetAttribute(k_col_finish_update<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    SB_CUDA(cudaFuncSetAttribute(k_col_reflector<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    SB_CUDA(cudaFuncSetAttribute(k_col_reflector<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    // Force-load every kernel of the multi-GPU path now. With lazy module loading the first launch of a kernel
    // may need a device-wide synchronisation, which would deadlock against a peer rank's kernel that is already
    // spinning on a flag this rank has not raised yet.
    cudaFuncAttributes fa;
    SB_CUDA(cudaFuncGetAttributes(&fa, k_col_gemv<false>));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_col_gemv<true>));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_panel_push));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_panel_pull));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_gather_rows));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_sum_peers));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_barrier));
    SB_CUDA(cudaFuncGetAttributes(&fa, splitk_reduce_kernel));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_is_identity));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_flag_to_double));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_identity_cols));
    SB_CUDA(cudaFuncGetAttributes(&fa, k_qcols_to_rows));
#define SB_FUSED_PREP(D, S) SB_CUDA(cudaFuncSetAttribute(k_panel_fused<D, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM_OPTIN))
    SB_FUSED_PREP(false, 0); SB_FUSED_PREP(false, 1); SB_FUSED_PREP(true, 0); SB_FUSED_PREP(true, 1);
#undef SB_FUSED_PREP
}

struct GemvPlan { int skip, RB, S, kc; const double *A0; };

// Largest matrix order the path is laid out for: the exchanged GEMV result has at most RB_MAX row blocks of 256 rows
// (panel.cuh) and the per-column kernels stage at most 2048 columns of v per block. Two FP64 n x n matrices of this order
// (137 GB each) exceed the memory of one B200 anyway; the entry points reject larger n instead of overrunning a buffer.
constexpr int SB_MAX_N = RB_MAX * 256 - 16;

// Capacity (doubles) of the GEMV partial-sum buffer for matrices up to order n:
//   * persistent panel kernel: every 128-thread group of the grid owns one slice per row block it touches:
//     at most (groups / RB + 2) slices of ldp <= 256 RB + 16 doubles, groups <= 148 * FUSED_VB = 592;
//   * per-column kernels (plan_gemv_for): S <= ceil(n / 2048) + 1 column chunks of ldp = roundup(m + 2, 16) doubles when
//     one wave of blocks cannot cover the matrix with shorter chunks (n > ~47000).
static inline size_t ypart_doubles(int n)
{
    const size_t ldp = (size_t)round_up(n + 2, 16);
    const size_t fused = (size_t)2 * 148 * 12 * 256 + 4 * ldp;
    const size_t unfused = ((size_t)ceil_div(n, 2048) + 2) * ldp;
    return std::max(fused, unfused);
}

// Decomposition of the per-column GEMV over rows [0, m) x ncols local columns: row blocks of 256 (padded) rows times S
// column chunks of kc <= 2048 columns, sized so that the grid is at most one full wave of `slots` resident blocks --
// or, for matrices too large for that, as few chunks as the 2048-column staging buffer allows. Pure host arithmetic
// (unit-tested through starneig_b200_plan_check). Returns S = 0 if the partial sums would not fit `ypart_cap` doubles.
static inline GemvPlan plan_gemv_for(int slots, int skip, int m, int ncols, int ldp, size_t ypart_cap)
{
    GemvPlan p;
    p.skip = skip; p.A0 = nullptr;
    const int mp = m + skip;
    p.RB = ceil_div(mp, 256);
    int S = std::max(1, slots / std::max(p.RB, 1));
    int kc = ceil_div(std::max(ncols, 1), S);
    kc = std::max(kc, 16);
    kc = std::min(round_up(kc, 4), 2048);
    S = std::max(1, ceil_div(ncols, kc));
    while ((size_t)S * ldp > ypart_cap && kc < 2048) { kc = std::min(2048, kc * 2); S = std::max(1, ceil_div(ncols, kc)); }
    p.kc = kc;
    p.S = (size_t)S * ldp > ypart_cap ? 0 : S;
    return p;
}

struct Stats : starneig_b200_stats {};

// private workspace of a rank
struct Workspace {
    int n_cap = 0, nb_cap = 0;
    int ldv = 0, nbp = 0;
    double *V = nullptr, *Y = nullptr, *VT = nullptr, *W = nullptr, *Wpart = nullptr;
    double *Vg = nullptr, *VTg = nullptr;      // rows of V, VT of the local columns (P > 1)
    size_t wpart_cap = 0;           // doubles
    double *pcol = nullptr, *ypart = nullptr;
    size_t ypart_cap = 0;           // doubles
    double *s = nullptr, *w2 = nullptr, *colpart = nullptr, *sqpart = nullptr;
    ColScal *scal = nullptr;
    unsigned *counter = nullptr;
    unsigned *gbar = nullptr;                   // grid barrier counter of the fused panel kernel
    unsigned long long *rbar = nullptr;         // its per-column arrival words (barrier after phase R + the vote on `linear`)
    unsigned long long *timers = nullptr;       // device-side phase timers of the fused panel kernel (ns)
    // Reflector history (backward accumulation of Q, Rank::reduce): V and VT = V T of EVERY panel, panel columns i .. i+w-1 at
    // columns i .. i+w-1 of two ldv x (n + 8) arrays (3.2 GB each at n = 20000, 20 GB each at n = 50000: HBM is 180 GB)
    double *Vh = nullptr, *VTh = nullptr;
    int hist_n = 0;
    std::vector<void *> allocs;

    template <typename T> T *alloc(size_t count)
    {
        void *p = nullptr;
        SB_CUDA(cudaMalloc(&p, count * sizeof(T) + 256));
        allocs.push_back(p);
        return (T *)p;
    }
    void release_history()
    {
        if (Vh) cudaFree(Vh);
        if (VTh) cudaFree(VTh);
        Vh = VTh = nullptr; hist_n = 0;
    }
    // after ensure(): room for the reflectors of a whole reduction of order n; false (nothing allocated) if the device
    // cannot spare the memory -- the caller then accumulates Q forward, panel by panel
    bool ensure_history(int n)
    {
        if (Vh && n <= hist_n) return true;
        release_history();
        const size_t bytes = ((size_t)ldv * (n_cap + 8) + 16) * sizeof(double);
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < 2 * bytes + ((size_t)4 << 30)) { cudaGetLastError(); return false; }
        if (cudaMalloc((void **)&Vh, bytes) != cudaSuccess) { cudaGetLastError(); Vh = nullptr; return false; }
        if (cudaMalloc((void **)&VTh, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(Vh); Vh = VTh = nullptr; return false; }
        hist_n = n_cap;
        return true;
    }
    void release()
    {
        release_history();
        for (void *p : allocs) cudaFree(p);
        allocs.clear();
        n_cap = nb_cap = 0;
    }
    void ensure(int n, int nb, bool dist)
    {
        if (n <= n_cap && nb <= nb_cap && (!dist || Vg != nullptr)) return;
        n = std::max(n, n_cap); nb = std::max(nb, nb_cap);
        release();
        n_cap = n; nb_cap = nb;
        ldv = round_up(n, 16);
        nbp = round_up(nb, 8);
        size_t panel = (size_t)ldv * nbp;
        V = alloc<double>(panel); Y = alloc<double>(panel); VT = alloc<double>(panel); W = alloc<double>(panel);
        if (dist) { Vg = alloc<double>(panel); VTg = alloc<double>(panel); }
        else Vg = VTg = nullptr;
        wpart_cap = 8 * (size_t)std::max(ldv, 4096) * nbp;
        Wpart = alloc<double>(wpart_cap);
        pcol = alloc<double>(ldv);
        ypart_cap = ypart_doubles(n);
        ypart = alloc<double>(ypart_cap);
        s = alloc<double>(nbp); w2 = alloc<double>(nbp);
        colpart = alloc<double>((size_t)nbp * PANEL_LDB);
        sqpart = alloc<double>(3 * PANEL_LDB);
        scal = alloc<ColScal>(nbp);
        counter = alloc<unsigned>(4);
        SB_CUDA(cudaMemset(counter, 0, 4 * sizeof(unsigned)));
        gbar = alloc<unsigned>(1024);
        rbar = alloc<unsigned long long>(FUSED_MAX_NB + 8);
        timers = alloc<unsigned long long>(8);
        SB_CUDA(cudaMemset(timers, 0, 8 * sizeof(unsigned long long)));
    }
};

// Peer-visible exchange arena of a rank: ONE cudaMalloc, so one pointer (same process) or one
// cudaIpcMemHandle (one process per GPU) is all that the ranks exchange. The layout is a pure function
// of (P, n_cap, nb_cap) and therefore identical on every rank.
struct ArenaLayout {
    int P = 0, n_cap = 0, nb_cap = 0, ldv = 0, nbp = 0, ldp = 0;
    size_t off_bar = 0, off_status = 0, off_rbcount = 0, off_yflag = 0, off_inbox = 0, off_pan = 0, off_wx = 0, off_qc = 0, bytes = 0;
    int qc_cols = 0;            // columns of the rank's column-distributed Q (backward accumulation, Rank::reduce); 0: no such region
    static size_t align(size_t x) { return (x + 255) / 256 * 256; }
    void set(int P_, int n, int nb, int cb = 0)
    {
        P = P_; n_cap = n; nb_cap = nb;
        ldv = round_up(n, 16); nbp = round_up(nb, 8); ldp = round_up(n + 2, 16);
        size_t o = 0;
        off_bar = o;      o = align(o + MAX_RANKS * sizeof(unsigned));
        off_status = o;   o = align(o + sizeof(unsigned));
        off_rbcount = o;  o = align(o + RB_MAX * sizeof(unsigned));
        off_yflag = o;    o = align(o + (size_t)2 * P * RB_MAX * sizeof(unsigned));
        off_inbox = o;    o = align(o + (size_t)2 * P * ldp * 16);      // 16-byte LL entries (fused kernel) or doubles
        off_pan = o;      o = align(o + (size_t)ldv * (nbp + 8) * sizeof(double));    // the panel and the column right of it
        off_wx = o;       o = align(o + (size_t)ldv * nbp * sizeof(double));
        // cb > 0: the rank's columns (block-cyclic like A: at most ceil(ceil(n / cb) / P) blocks) of Q = H_0 ... H_K-1, formed
        // backward after the last panel and pulled into the row slabs by the peers (k_qcols_to_rows)
        qc_cols = cb > 0 ? ceil_div(ceil_div(n, cb), P) * cb : 0;
        off_qc = qc_cols > 0 ? o : 0;
        o = align(o + (size_t)ldv * qc_cols * sizeof(double));
        bytes = o;
    }
};

// Widest panel <= nb whose row-block kernels fit their shared-memory layout for a panel of m rows on `ctas` CTAs: both
// panel paths keep 3-4 partial vectors per owned row and 32 panel columns in shared memory, so for very large matrices
// (n > ~70000 at the reference's default width) the panel is narrowed -- the same reduction with another blocking, as
// for panels wider than PANEL_MAX_NB. Returns 0 if not even 32 columns fit (cannot happen for m <= SB_MAX_N).
static inline int fit_panel_width(int m, int nb, int ctas, bool fused)
{
    for (int w = nb; w >= 8; w = (w > 32 ? std::max(32, (w - 8) / 8 * 8) : w - 8)) {
        const bool try_fused = fused && w <= FUSED_MAX_NB;
        if (try_fused) {
            const int nsub = std::max(1, ceil_div(m, 32 * std::max(ctas, 1)));
            if (fused_smem_bytes(std::min(w, std::max(m, 1)), nsub, FUSED_KC) <= PANEL_SMEM_MAX) return w;
        } else {
            const int nsub = std::max(1, ceil_div(m, 32 * PANEL_MAX_BLOCKS));
            const int NW = std::max(1, ceil_div(w, 32));
            const int RS = std::max(1, std::min(nsub, (NW <= 16 ? 16 : 32) / NW));
            const size_t smem_fu = (size_t)(2 * w + nsub * 4 * NW * 32 + 2 * nsub * 32 + (RS + 3) * NW * 32) * sizeof(double);
            if (smem_fu <= PANEL_SMEM_MAX) return w;
        }
        if (w <= 32) break;
    }
    return 0;
}

// Host staging hooks of a reduction (host-pointer API, hessenberg.cu): the upload of Q hides behind the first
// column loop and finished columns travel back while the next panels are factorised.
struct StageHook {
    // the stream is about to touch Q for the first time
    virtual void before_q(cudaStream_t s) = 0;
    // Everything enqueued on `s` so far completes panel [.., final_cols): global columns [0, final_cols) of A and -- with
    // `q_too` -- [0, final_cols] of Q will not change any more (backward accumulation: no column of Q is final before the end).
    virtual void panel_done(cudaStream_t s, int final_cols, bool q_too) = 0;
    virtual ~StageHook() {}
};
// the rank's row slab of Q: rows [q0, q1)
static inline void q_row_range(int P, int g, int n, int *q0, int *q1)
{
    const int per = round_up(ceil_div(n, P), 8);
    *q0 = std::min(n, g * per);
    *q1 = std::min(n, (g + 1) * per);
}

struct PanelGrid { int blocks; TileGeom tg; size_t smem_fu, smem_rf; };

struct Rank {
    int P = 1, g = 0, device = 0, cb = 64;
    bool ready = false;
    cudaStream_t stream = nullptr;          // the launch sequence of the reduction
    cudaStream_t copy = nullptr;            // host staging that overlaps the reduction (Q upload, write-back of finished columns)
    cudaEvent_t ev_q_up = nullptr, ev_cols_final = nullptr;
    double *host_word = nullptr;            // page-locked: results the host waits for in the middle of a reduction. (A device-to-host
                                            // copy into PAGEABLE memory blocks inside the driver while it waits for the stream; ranks
                                            // that share a device then cannot launch the kernels this stream is waiting for.)
    Workspace ws;
    ArenaLayout al;
    char *arena = nullptr;                  // own arena (device memory on `device`)
    char *peer[MAX_RANKS] = {};             // arena base of every rank as seen from this rank (peer[g] == arena)
    bool arena_is_ipc[MAX_RANKS] = {};
    unsigned bar_epoch = 0, y_epoch = 0;
    Stats stats{};
    int profile_level = 1;
    int gemv_slots = 0;                     // resident k_col_gemv blocks on the whole GPU (one wave)
    int fused = 1;                          // 1: one persistent kernel per panel (panel_fused.cuh); 0: three kernels per column
    int fused_ctas = 0;                     // grid of the fused kernel (number of SMs; fewer when ranks share a device)
    int gemv_kc = FUSED_KC;                 // fused kernel: columns of v staged per GEMV group at a time (2048 timed: no gain)
    int gemm_tma = 3;                       // DMMA kernels fed by the TMA engine (dgemm_tma.cuh): bit 0 = the rank-nb updates (NT),
                                            // bit 1 = the skinny products (TN, NN); 0: the cp.async kernels (dgemm.cuh)
    int gemv_linear = 1;                    // fused kernel: the GEMV streams against the unscaled x (FusedArgs::linear)
    int q_backward = 256;                   // an identity Q on entry is accumulated backward after the reduction (reduce()) for
                                            // matrices of at least this order; 0: never (STARNEIG_B200_Q_BACKWARD)
    int fused_slabs = 1;                    // fused kernel: the CTA's rows of V in shared memory when they fit (FusedSmem)
    std::vector<cudaEvent_t> events;        // phase events: 4 per panel
    std::vector<cudaEvent_t> gemv_events;   // 4 per timed column (profile level 2)
    size_t gemv_events_used = 0;

    void open(int P_, int g_, int device_)
    {
        if (ready) return;
        P = P_; g = g_; device = device_;
        SB_CUDA(cudaSetDevice(device));
        SB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        SB_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreateWithFlags(&ev_q_up, cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&ev_cols_final, cudaEventDisableTiming));
        prepare_device_functions();
        SB_CUDA(cudaMallocHost((void **)&host_word, 64));
        const char *e = getenv("STARNEIG_B200_COL_BLOCK");
        if (e && atoi(e) >= 8) cb = atoi(e) / 8 * 8;
        e = getenv("STARNEIG_B200_FUSED_PANEL");
        if (e) fused = atoi(e);
        SB_CUDA(cudaDeviceGetAttribute(&fused_ctas, cudaDevAttrMultiProcessorCount, device));
        e = getenv("STARNEIG_B200_FUSED_CTAS");
        if (e && atoi(e) >= 1) fused_ctas = std::min(fused_ctas, atoi(e));
        e = getenv("STARNEIG_B200_GEMV_KC");
        if (e && atoi(e) >= 64) gemv_kc = std::min(4096, atoi(e) / 8 * 8);
        e = getenv("STARNEIG_B200_GEMM_TMA");
        if (e) gemm_tma = atoi(e);
        e = getenv("STARNEIG_B200_GEMV_LINEAR");
        if (e) gemv_linear = atoi(e);
        e = getenv("STARNEIG_B200_Q_BACKWARD");
        if (e) q_backward = atoi(e);
        e = getenv("STARNEIG_B200_FUSED_SLABS");
        if (e) fused_slabs = std::max(0, std::min(1, atoi(e)));
        int coop = 0;
        SB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        if (!coop) fused = 0;
        ready = true;
    }
    void close()
    {
        if (!ready) return;
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        ws.release();
        for (auto e : events) cudaEventDestroy(e);
        for (auto e : gemv_events) cudaEventDestroy(e);
        events.clear(); gemv_events.clear();
        for (int s = 0; s < P; s++)
            if (s != g && arena_is_ipc[s] && peer[s]) { cudaIpcCloseMemHandle(peer[s]); peer[s] = nullptr; }
        if (arena) cudaFree(arena);
        arena = nullptr;
        al = ArenaLayout();
        if (host_word) cudaFreeHost(host_word);
        host_word = nullptr;
        cudaEventDestroy(ev_q_up); cudaEventDestroy(ev_cols_final);
        cudaStreamDestroy(stream);
        cudaStreamDestroy(copy);
        stream = copy = nullptr;
        ready = false;
    }
    // (re)allocates the own arena for (n, nb); returns true if a new allocation was made (peers must re-exchange)
    bool ensure_arena(int n, int nb)
    {
        if (arena && n <= al.n_cap && nb <= al.nb_cap) return false;
        n = std::max(n, al.n_cap); nb = std::max(nb, al.nb_cap);
        SB_CUDA(cudaSetDevice(device));
        if (arena) { SB_CUDA(cudaDeviceSynchronize()); SB_CUDA(cudaFree(arena)); }
        al.set(P, n, nb, q_backward > 0 ? cb : 0);
        SB_CUDA(cudaMalloc((void **)&arena, al.bytes));
        SB_CUDA(cudaMemset(arena, 0, al.off_pan));          // flags, counters, inbox
        SB_CUDA(cudaDeviceSynchronize());
        for (int s = 0; s < P; s++) peer[s] = nullptr;
        peer[g] = arena;
        bar_epoch = 0; y_epoch = 0;
        return true;
    }
    template <typename T> T *at(int s, size_t off) const { return (T *)(peer[s] + off); }

    cudaEvent_t phase_event(size_t idx)
    {
        while (events.size() <= idx) { cudaEvent_t e; SB_CUDA(cudaEventCreate(&e)); events.push_back(e); }
        return events[idx];
    }
    cudaEvent_t gemv_event(size_t idx)
    {
        while (gemv_events.size() <= idx) { cudaEvent_t e; SB_CUDA(cudaEventCreate(&e)); gemv_events.push_back(e); }
        return gemv_events[idx];
    }

    // -----------------------------------------------------------------------------------------
    enum GemmKind { GEMM_NT, GEMM_TN, GEMM_NN };

    // C = alpha*op(A)*op(B) + beta*C on the rank's stream; split-K through ws.Wpart for skinny outputs
    // k_guard: see GemmTmaConfig::launch (an operand whose k runs over the panel's rows may start at an odd element; the
    // engine's V / VT / Y buffers carry a zero row in front for that case)
    void gemm(GemmKind kind, int M, int N, int K, double alpha, const double *A, int lda,
              const double *B, int ldb, double beta, double *C, int ldc, bool k_guard = false)
    {
        if (M < 1 || N < 1) return;
        cudaStream_t st = stream;
        double *wpart = ws.Wpart;
        stats.gemm_flops += 2.0 * M * N * (double)K;
        if (kind == GEMM_NT) {
            if ((gemm_tma & 1) && TmaNT::launch(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 1, K, 0, 0, k_guard)) stats.gemm_tma_launches++;
            else { GemmNT::launch(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 1, K, 0); stats.gemm_cpasync_launches++; }
            stats.kernel_launches++;
            return;
        }
        // skinny output (N = panel width): pick the column tile with the least padding, split K if the
        // grid would not fill the GPU twice
        int bn = (ceil_div(N, 96) * 96 <= ceil_div(N, 104) * 104) ? 96 : 104;
        // 2 CTAs per SM are resident; split K so that the grid is >= ~8 waves (tail quantisation < ~6 %)
        int tiles = ceil_div(M, 64) * ceil_div(N, bn);
        int splits = 1;
        const int want = 8 * 2 * 148;
        if (beta == 0.0 && alpha == 1.0 && wpart != nullptr && tiles < want) {
            splits = std::min(32, ceil_div(want, tiles));
            splits = std::min(splits, std::max(1, K / 512));
            while (splits > 1 && (size_t)splits * ldc * N > ws.wpart_cap) splits--;
        }
        // (K + 1: a k frame moved by one element -- GemmTmaConfig::launch -- must still be covered by the slices)
        int klen = round_up(std::max(1, ceil_div(K + 1, splits)), GEMM_BK);
        splits = std::max(1, ceil_div(K + 1, klen));
        double *out = splits > 1 ? wpart : C;
        size_t stride = splits > 1 ? (size_t)ldc * N : 0;
        double b = splits > 1 ? 0.0 : beta;
#define SB_SKINNY(CFG) CFG::launch(st, M, N, K, alpha, A, lda, B, ldb, b, out, ldc, splits, klen, stride, 1)
#define SB_SKINNY_TMA(CFG) ((gemm_tma & 2) && CFG::launch(st, M, N, K, alpha, A, lda, B, ldb, b, out, ldc, splits, klen, stride, 1, k_guard))
        if (kind == GEMM_TN) {
            if (bn == 96) { if (SB_SKINNY_TMA(TmaTN12)) stats.gemm_tma_launches++; else { SB_SKINNY(GemmTN12); stats.gemm_cpasync_launches++; } }
            else          { if (SB_SKINNY_TMA(TmaTN13)) stats.gemm_tma_launches++; else { SB_SKINNY(GemmTN13); stats.gemm_cpasync_launches++; } }
        } else {
            if (bn == 96) { if (SB_SKINNY_TMA(TmaNN12)) stats.gemm_tma_launches++; else { SB_SKINNY(GemmNN12); stats.gemm_cpasync_launches++; } }
            else          { if (SB_SKINNY_TMA(TmaNN13)) stats.gemm_tma_launches++; else { SB_SKINNY(GemmNN13); stats.gemm_cpasync_launches++; } }
        }
#undef SB_SKINNY
#undef SB_SKINNY_TMA
        stats.kernel_launches++;
        if (splits > 1) {
            dim3 grid(ceil_div(M, 256), N);
            SB_LAUNCH(splitk_reduce_kernel, grid, 256, 0, st, M, N, splits, out, ldc, stride, C, ldc);
            stats.kernel_launches++;
        }
    }

    // -----------------------------------------------------------------------------------------
    // panel factorisation: columns i .. i+w-1
    // -----------------------------------------------------------------------------------------
    PanelArgs make_panel_args(int m, double *V, double *Y, double *VT, int ld)
    {
        PanelArgs pa;
        pa.m = m; pa.ld = ld; pa.V = V; pa.Y = Y; pa.VT = VT;
        pa.pcol = ws.pcol; pa.ypart = ws.ypart; pa.ldp = round_up(m + 2, 16);
        pa.s = ws.s; pa.w2 = ws.w2; pa.colpart = ws.colpart; pa.ldt = ws.nbp;
        pa.sqpart = ws.sqpart; pa.scal = ws.scal; pa.counter = ws.counter;
        return pa;
    }

    // decomposition of the GEMV over rows [0,m) x local columns [0,ncols) starting at `base`: row blocks of
    // 256 padded rows times S column chunks, sized so that the grid is (at most) one full wave
    GemvPlan plan_gemv(const double *base, int m, int ncols, int ldp)
    {
        if (gemv_slots == 0) {
            int occ = 0, sms = 0;
            SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_col_gemv<false>, GEMV_THREADS, 2048 * sizeof(double)));
            gemv_slots = std::max(1, occ) * sms;
        }
        const int skip = (int)(((uintptr_t)base / sizeof(double)) & 1);
        GemvPlan p = plan_gemv_for(gemv_slots, skip, m, ncols, ldp, ws.ypart_cap);
        if (p.S == 0) fatal("matrix too large for the GEMV partial-sum buffer", __FILE__, __LINE__);
        p.A0 = base - skip;
        return p;
    }

    // row blocks (<= PANEL_MAX_BLOCKS) and warp layout of the two row-block kernels for a panel of m rows
    // at column j (cols = number of columns the warps must cover)
    static PanelGrid panel_grid(int m, int cols, int j)
    {
        PanelGrid g;
        TileGeom &tg = g.tg;
        tg.nsub = std::max(1, ceil_div(m, 32 * PANEL_MAX_BLOCKS));
        g.blocks = ceil_div(m, 32 * tg.nsub);
        tg.NW = std::max(1, ceil_div(cols, 32));
        tg.RS = std::max(1, std::min(tg.nsub, (tg.NW <= 16 ? 16 : 32) / tg.NW));      // <= 512 threads unless the panel is wider than 512
        // a few warps at least: they share the sum over the GEMV partials and hide latency
        while (tg.NW * tg.RS < 4 && tg.NW * (tg.RS + 1) <= 32 && tg.RS < 4) tg.RS++;
        g.smem_fu = (size_t)(2 * j + tg.nsub * 4 * tg.NW * 32 + 2 * tg.nsub * 32 + tg.RS * tg.NW * 32) * sizeof(double);
        g.smem_rf = (size_t)(j + tg.nsub * tg.NW * 32 + tg.nsub * 32 + tg.RS * tg.NW * 32 + 96) * sizeof(double);
        if (g.smem_fu > PANEL_SMEM_MAX) fatal("matrix too large for the panel kernels' shared-memory layout", __FILE__, __LINE__);
        return g;
    }

    void launch_finish_update(const PanelArgs &pa, int j, int S, const double *yin, double *acol, int do_update, const YWait &yw)
    {
        PanelGrid pg = panel_grid(pa.m, j, j);
        const int threads = 32 * pg.tg.NW * pg.tg.RS;
        if (threads <= 512) SB_LAUNCH((k_col_finish_update<512>), pg.blocks, threads, pg.smem_fu, stream, pa, j, S, yin, acol, do_update, pg.tg, yw);
        else                SB_LAUNCH((k_col_finish_update<1024>), pg.blocks, threads, pg.smem_fu, stream, pa, j, S, yin, acol, do_update, pg.tg, yw);
        stats.kernel_launches++;
    }

    void launch_reflector(const PanelArgs &pa, int j, double *acol)
    {
        PanelGrid pg = panel_grid(pa.m, j, j);
        const int threads = 32 * pg.tg.NW * pg.tg.RS;
        if (threads <= 512) SB_LAUNCH((k_col_reflector<512>), pg.blocks, threads, pg.smem_rf, stream, pa, j, acol, pg.tg);
        else                SB_LAUNCH((k_col_reflector<1024>), pg.blocks, threads, pg.smem_rf, stream, pa, j, acol, pg.tg);
        stats.kernel_launches++;
    }

    Xchg make_xchg()
    {
        Xchg x;
        memset(&x, 0, sizeof(x));
        x.P = P; x.g = g;
        if (P > 1) {
            for (int s = 0; s < P; s++) { x.inbox[s] = at<double>(s, al.off_inbox); x.yflag[s] = at<unsigned>(s, al.off_yflag); }
            x.rbcount = at<unsigned>(g, al.off_rbcount);
            x.status = at<unsigned>(g, al.off_status);
        } else {
            x.status = ws.counter + 3;      // time-out word of the LL waits inside the fused kernel
        }
        return x;
    }

    // Columns i .. i+w-1 of the reduction of rows/cols < end. `pan` points at row i+1 of panel column 0 (leading
    // dimension ldpan): the matrix itself (P == 1) or the replicated panel buffer. A_loc is the rank's column
    // storage (leading dimension ldA), cm its column map.
    void panel_factor(const ColMap &cm, int i, int end, int w, const double *A_loc, int ldA, double *pan, int ldpan,
                      double *V, double *Y, double *VT, int ld, int ctas = 0)
    {
        if (ctas < 1 || ctas > fused_ctas) ctas = fused_ctas;
        cudaStream_t st = stream;
        const int m = end - i - 1;
        PanelArgs pa = make_panel_args(m, V, Y, VT, ld);
        Xchg x = make_xchg();
        const int lc_end = cm.lower(end);
        if (fused && w <= FUSED_MAX_NB) {
            FusedArgs f;
            memset(&f, 0, sizeof(f));
            f.a = pa; f.w = w; f.i = i; f.pan = pan; f.ldpan = ldpan; f.Aloc = A_loc; f.lda = ldA; f.cm = cm; f.lc_end = lc_end;
            f.nsub = std::max(1, ceil_div(m, 32 * ctas));
            f.rpc = 32 * f.nsub;
            f.gbar = ws.gbar; f.rbar = ws.rbar; f.timers = ws.timers;
            f.linear = gemv_linear;
            f.x = x;
            f.x.epoch = y_epoch + 1;
            f.kc = gemv_kc;
            size_t smem = fused_smem_bytes(w, f.nsub, f.kc);
            if (smem > PANEL_SMEM_MAX) { f.kc = FUSED_KC; smem = fused_smem_bytes(w, f.nsub, f.kc); }
            if (smem <= PANEL_SMEM_MAX) {
                y_epoch += w;
                SB_CUDA(cudaMemsetAsync(ws.gbar, 0, 1024 * sizeof(unsigned), st));
                SB_CUDA(cudaMemsetAsync(ws.rbar, 0, (FUSED_MAX_NB + 8) * sizeof(unsigned long long), st));
                // the CTA's rows of V in shared memory when they fit beside the fixed layout (m <= 32 * ctas rows at the
                // AED client's width 224, twice that at width 192)
                int slabs = fused_slabs;
                while (slabs > 0 && fused_smem_bytes(w, f.nsub, f.kc, slabs) > FUSED_SLAB_SMEM_MAX) slabs--;
                smem = fused_smem_bytes(w, f.nsub, f.kc, slabs);
#define SB_FUSED_GO(S) do { if (P > 1) SB_LAUNCH_COOP((k_panel_fused<true, S>), ctas, FUSED_THREADS, smem, st, f); \
                            else       SB_LAUNCH_COOP((k_panel_fused<false, S>), ctas, FUSED_THREADS, smem, st, f); } while (0)
                if (slabs) SB_FUSED_GO(1); else SB_FUSED_GO(0);
#undef SB_FUSED_GO
                stats.kernel_launches++;
                stats.fused_panels++;
                stats.fused_slab_panels[slabs]++;
                for (int j = 0; j < w; j++) {
                    const double bytes = 8.0 * (double)m * (lc_end - cm.lower(i + j + 1));
                    stats.gemv_bytes += bytes; stats.gemv_timed_bytes += bytes;
                }
                stats.gemv_launches += w; stats.gemv_timed_launches += w;
                return;
            }
        }
        SB_CUDA(cudaMemsetAsync(V, 0, (size_t)ld * w * sizeof(double), st));
        int S_prev = 0;
        const double *yin = ws.ypart;
        YWait yw; memset(&yw, 0, sizeof(yw));
        for (int j = 0; j <= w; j++) {
            const int c = i + j;
            double *acol = j < w ? pan + (size_t)j * ldpan : nullptr;
            const bool timed = j < w && (profile_level >= 3 || (profile_level == 2 && (j & 7) == 4));
            if (timed) SB_CUDA(cudaEventRecord(gemv_event(gemv_events_used++), st));
            // finish column j-1 (Y, VT) and start column j; the last pass (j == w) only finishes
            if (j > 0) launch_finish_update(pa, j, S_prev, yin, acol, j < w ? 1 : 0, yw);
            if (j == w) break;
            if (timed) SB_CUDA(cudaEventRecord(gemv_event(gemv_events_used++), st));
            launch_reflector(pa, j, acol);
            if (timed) SB_CUDA(cudaEventRecord(gemv_event(gemv_events_used++), st));

            const int ncols = m - j;                         // length of v
            const int lc0 = cm.lower(c + 1), nloc = lc_end - lc0;
            const double *base = A_loc + (size_t)lc0 * ldA + i + 1;
            GemvPlan gp = plan_gemv(base, m, nloc, pa.ldp);
            size_t sh = (size_t)gp.kc * sizeof(double);
            if (P == 1) {
                SB_LAUNCH((k_col_gemv<false>), gp.RB * gp.S, GEMV_THREADS, sh, st, pa, j, ncols, cm, lc0, nloc, c + 1, gp.A0, ldA, gp.skip,
                                                                           gp.kc, gp.RB, gp.S, acol, x);
                S_prev = gp.S;
            } else {
                x.epoch = ++y_epoch;
                SB_LAUNCH((k_col_gemv<true>), gp.RB * gp.S, GEMV_THREADS, sh, st, pa, j, ncols, cm, lc0, nloc, c + 1, gp.A0, ldA, gp.skip,
                                                                          gp.kc, gp.RB, gp.S, acol, x);
                const int par = x.epoch & 1;
                S_prev = P;
                yin = at<double>(g, al.off_inbox) + (size_t)par * P * pa.ldp;
                yw.flags = at<unsigned>(g, al.off_yflag) + (size_t)par * P * RB_MAX;
                yw.epoch = x.epoch; yw.P = P; yw.skip = gp.skip; yw.status = x.status;
            }
            if (timed) {
                SB_CUDA(cudaEventRecord(gemv_event(gemv_events_used++), st));
                stats.gemv_timed_launches++;
                stats.gemv_timed_bytes += 8.0 * (double)m * nloc;
            }
            stats.kernel_launches++;
            stats.gemv_launches++;
            stats.gemv_bytes += 8.0 * (double)m * nloc;
        }
    }

    void barrier()
    {
        if (P == 1) return;
        BarPtrs b;
        for (int s = 0; s < MAX_RANKS; s++) b.p[s] = s < P ? at<unsigned>(s, al.off_bar) : nullptr;
        SB_LAUNCH(k_barrier, 1, MAX_RANKS, 0, stream, P, g, ++bar_epoch, b, at<unsigned>(g, al.off_status));
        stats.kernel_launches++;
    }

    // room for the reflectors of a whole reduction when reduce() may accumulate Q backward (same conditions as there); callers
    // that drive several ranks allocate it before the ranks start
    void prepare_history(int n, int begin, int end, bool has_q)
    {
        if (q_backward > 0 && n >= q_backward && begin == 0 && end == n && has_q) ws.ensure_history(n);
    }

    // X(rows x m) <- X (I - V T V^T) = X - (X VT) V^T  (reference update_right_a/b, src/hessenberg/cpu.c:443-560)
    void deferred_right_update(int rows, int m, int w, double *X, int ldx, const double *V, const double *VT, int ld, double *W)
    {
        gemm(GEMM_NN, rows, w, m, 1.0, X, ldx, VT, ld, 0.0, W, ld, true);
        gemm(GEMM_NT, rows, m, w, -1.0, W, ld, V, ld, 1.0, X, ldx);
    }

    // -----------------------------------------------------------------------------------------
    // the whole reduction on this rank's shards: A_loc = local columns (full height n, leading dimension ldA),
    // Q_loc = rows [q0, q0+qrows) of Q (all n columns, leading dimension ldQ). P == 1: the matrices themselves.
    // -----------------------------------------------------------------------------------------
    void reduce(int n, int begin, int end, int nb, double *A, int ldA, double *Q, int ldQ, int qrows, StageHook *hook = nullptr)
    {
        SB_CUDA(cudaSetDevice(device));
        cudaStream_t st = stream;
        nb = std::min(nb, PANEL_MAX_NB);      // wider panels are split; the result only differs in rounding
        if (end - begin > 1) {
            const int fit = fit_panel_width(end - begin - 1, nb, fused_ctas, fused != 0);
            if (fit == 0) fatal("matrix too large for the panel kernels' shared-memory layout", __FILE__, __LINE__);
            nb = fit;
        }
        stats.panel_width_used = nb;
        ws.ensure(n, nb, P > 1);
        if (P > 1 && (arena == nullptr || n > al.n_cap || nb > al.nb_cap))
            fatal("internal error: exchange arena not prepared", __FILE__, __LINE__);
        const int ld = ws.ldv;
        const int lvl = profile_level;
        const ColMap cm{P, g, P == 1 ? std::max(n, 1) : cb};
        gemv_events_used = 0;
        SB_CUDA(cudaMemsetAsync(ws.timers, 0, 8 * sizeof(unsigned long long), st));
        int panel = 0;
        cudaEvent_t ev_first = phase_event(0), ev_last = phase_event(1);
        barrier();
        SB_CUDA(cudaEventRecord(ev_first, st));

        PeerPtrs panp, wxp;
        for (int s = 0; s < MAX_RANKS; s++) {
            panp.p[s] = (P > 1 && s < P) ? at<double>(s, al.off_pan) : nullptr;
            wxp.p[s] = (P > 1 && s < P) ? at<double>(s, al.off_wx) : nullptr;
        }
        const int lc_end = cm.lower(end);

        // Backward accumulation of Q (one GPU, full reduction, Q = I on entry -- the reference driver's and the usual
        // caller's case; LAPACK's dorghr does the same). Q U for a general Q costs 2 n^3 flops, panel by panel (forward: the
        // reference's order, core.c:338-340). U itself = H_0 H_1 ... H_K-1 applied to the identity from the LAST panel to the
        // first touches only the trailing (n - i - 1)^2 block at panel i: 4/3 n^3 flops, 2/3 n^3 (14 % of the level-3 work
        // of a reduction) less. It needs V and VT of every panel until the end: the panels write them into the history
        // arrays (Workspace::Vh / VTh) instead of one recycled buffer. Whether Q is the identity is decided on the device
        // when Q is touched for the first time (after the first panel; the upload of a host Q hides behind that panel).
        // Several GPUs: each rank forms ITS COLUMNS of the product (block-cyclic like A; the left-multiplications need no
        // communication at all), and the ranks then pull their row slabs out of the peers' column blocks over NVLink
        // (k_qcols_to_rows). Every rank must take the same branch (the barriers have to match): `bw_possible` depends on
        // the arguments and the environment only, and the ranks agree on "every slab is a slab of the identity and every
        // rank has room for the history" by a sum over ranks when Q is touched for the first time.
        int q0 = 0, q1 = n;
        if (P > 1) q_row_range(P, g, n, &q0, &q1);
        const bool bw_possible = q_backward > 0 && n >= q_backward && begin == 0 && end == n && Q != nullptr &&
                                 (P == 1 ? qrows == n : (al.off_qc != 0 && qrows == q1 - q0));
        const bool hist = bw_possible && ws.ensure_history(n);
        bool backward = false;
        stats.q_backward = 0;
        // STARNEIG_B200_TRACE: host-side progress of the rank on stderr (where is which rank when a cross-GPU wait times out)
        const bool tracing = getenv("STARNEIG_B200_TRACE") != nullptr;
        const auto trace_t0 = std::chrono::steady_clock::now();
        auto trace = [&](const char *msg, int panel_no) {
            if (!tracing) return;
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - trace_t0).count();
            fprintf(stderr, "[starneig][trace] rank %d (device %d) +%.1f ms: %s (panel %d, barriers %u)\n", g, device, ms, msg, panel_no, bar_epoch);
        };

        // Schedule: one stream. Per panel: column loop (persistent kernel), trailing right / left updates, then the updates
        // the reference defers (rows above the panel, columns right of the block, Q). Running the deferred updates next to
        // the column loops -- on a side stream with SMs set aside, or co-resident on the same SMs -- was measured on B200 at 1
        // and at 8 GPUs and lost both times: the GEMV needs every SM to saturate HBM and the part is power-limited
        // (profiles/r1_s5_overlap_sweep.txt, profiles/r2_v1_switch_sweep_n20000.txt,
        // profiles/r2_v5_visit8b_panelwidth_by_gpus_overlap.log); those variants were removed in round 2.
        for (int i = begin; i < end - 1; i += nb, panel++) {
            const int w = std::min(nb, end - i - 1);
            const int m = end - i - 1;
            // V and VT of the panel are stored with the row parity of the panel's first row in A (row i + 1; A is 16-byte
            // aligned with an even leading dimension), behind a zero guard row: the products over the panel's rows whose
            // operands are both K-major (W = A^T VT) then agree on where 16-byte units start, and moving the k frame by one
            // element for TMA adds a term that is zero (dgemm_tma.cuh)
            const int par = (i + 1) & 1;
            double *const Vbase = hist ? ws.Vh + (size_t)i * ld : ws.V, *const VTbase = hist ? ws.VTh + (size_t)i * ld : ws.VT;
            double *V = Vbase + par, *VT = VTbase + par;
            if (par) {
                SB_CUDA(cudaMemset2DAsync(Vbase, (size_t)ld * sizeof(double), 0, sizeof(double), (size_t)w, st));
                SB_CUDA(cudaMemset2DAsync(VTbase, (size_t)ld * sizeof(double), 0, sizeof(double), (size_t)w, st));
            }
            if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 0), st));
            const int pl0 = cm.lower(i), pl1 = cm.lower(i + w);
            const int ctas = fused_ctas;
            if (P == 1) {
                panel_factor(cm, i, end, w, A, ldA, A + (size_t)i * ldA + i + 1, ldA, V, ws.Y, VT, ld, ctas);
            } else {
                // gather the panel on every rank; the first barrier protects Pan and Wx of the previous panel
                barrier();
                // (the column right of the panel travels too: the row owners add it to the linear GEMV of the last column)
                const int pl1x = cm.lower(std::min(i + w + 1, end));
                if (pl1x > pl0) {
                    SB_LAUNCH(k_panel_push, dim3(ceil_div(m, 1024), pl1x - pl0), 256, 0, st, cm, i, pl0, m, A, ldA, panp, al.ldv);
                    stats.kernel_launches++;
                }
                barrier();
                double *pan = panp.p[g];
                panel_factor(cm, i, end, w, A, ldA, pan, al.ldv, V, ws.Y, VT, ld, ctas);
                if (pl1 > pl0) {
                    SB_LAUNCH(k_panel_pull, dim3(std::min(64, ceil_div(m, 256)), pl1 - pl0), 256, 0, st, cm, i, pl0, m, A, ldA, pan, al.ldv);
                    stats.kernel_launches++;
                }
            }
            if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 1), st));

            // rows of V, VT that belong to the local columns of the global range [i+1, end)
            const int cl0 = cm.lower(i + 1), ncl = lc_end - cl0;
            const double *Vg = V, *VTg = VT;
            int ldg = ld;
            if (P > 1) {
                Vg = ws.Vg; VTg = ws.VTg;
                if (ncl > 0) {
                    SB_LAUNCH(k_gather_rows, dim3(ceil_div(ncl, 128), w), 128, 0, st, cm, cl0, ncl, i + 1, w, V, VT, ld, ws.Vg, ws.VTg, ldg);
                    stats.kernel_launches++;
                }
            }

            const int tl0 = cm.lower(i + w), ntr = lc_end - tl0;
            if (ntr > 0) {
                double *Atr = A + (size_t)tl0 * ldA + i + 1;
                gemm(GEMM_NT, m, ntr, w, -1.0, ws.Y, ld, Vg + (tl0 - cl0), ldg, 1.0, Atr, ldA);
                gemm(GEMM_TN, ntr, w, m, 1.0, Atr, ldA, VT, ld, 0.0, ws.W, ld, true);
                gemm(GEMM_NT, m, ntr, w, -1.0, V, ld, ws.W, ld, 1.0, Atr, ldA);
            }
            if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 2), st));

            // ---- deferred updates
            if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 3), st));
            {   // rows above the panel
                double *X = A + (size_t)cl0 * ldA;
                if (P == 1) {
                    deferred_right_update(i + 1, m, w, X, ldA, V, VT, ld, ws.W);
                } else {
                    gemm(GEMM_NN, i + 1, w, ncl, 1.0, X, ldA, VTg, ldg, 0.0, wxp.p[g], al.ldv);
                    barrier();
                    SB_LAUNCH(k_sum_peers, dim3(ceil_div(i + 1, 256), w), 256, 0, st, P, i + 1, wxp, al.ldv, ws.W, ld);
                    stats.kernel_launches++;
                    gemm(GEMM_NT, i + 1, ncl, w, -1.0, ws.W, ld, Vg, ldg, 1.0, X, ldA);
                }
            }
            if (end < n) {   // columns right of the reduced block (partial reduction)
                const int xl0 = cm.lower(end), nx = cm.lower(n) - xl0;
                double *X = A + (size_t)xl0 * ldA + i + 1;
                gemm(GEMM_TN, nx, w, m, 1.0, X, ldA, VT, ld, 0.0, ws.W, ld, true);
                gemm(GEMM_NT, m, nx, w, -1.0, V, ld, ws.W, ld, 1.0, X, ldA);
            }
            if (hook && panel == 0) hook->before_q(st);
            if (bw_possible && panel == 0) {
                trace("order of the Q accumulation: asking the device", panel);
                // is Q the identity? (one pass over Q: ~0.5 ms at n = 20000; the host waits for the answer once per reduction)
                unsigned *flag = ws.counter + 2;
                SB_CUDA(cudaMemsetAsync(flag, hist ? 0 : 1, sizeof(unsigned), st));
                if (hist) {
                    SB_LAUNCH(k_is_identity, dim3(std::min(n, 148 * 8)), 256, 0, st, qrows, q0, n, Q, ldQ, flag);
                    stats.kernel_launches++;
                }
                if (P == 1) {
                    unsigned *h_flag = (unsigned *)host_word;
                    *h_flag = 1;
                    SB_CUDA(cudaMemcpyAsync(h_flag, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
                    SB_CUDA(cudaStreamSynchronize(st));
                    backward = *h_flag == 0;
                } else {
                    // sum over ranks of "my slab is not the identity / I have no room": through Wx, like the top-row products
                    double *h_sum = host_word;
                    *h_sum = 1.0;
                    barrier();                  // every rank has read Wx of the top-row update above
                    SB_LAUNCH(k_flag_to_double, 1, 32, 0, st, flag, wxp.p[g]);
                    barrier();
                    SB_LAUNCH(k_sum_peers, dim3(1, 1), 32, 0, st, P, 1, wxp, al.ldv, ws.W, ld);
                    stats.kernel_launches += 2;
                    SB_CUDA(cudaMemcpyAsync(h_sum, ws.W, sizeof(double), cudaMemcpyDeviceToHost, st));
                    SB_CUDA(cudaStreamSynchronize(st));
                    backward = *h_sum == 0.0;
                }
                stats.q_backward = backward ? 1 : 0;
                trace(backward ? "order of the Q accumulation: backward" : "order of the Q accumulation: forward", panel);
            }
            if (qrows > 0 && !backward)     // Q <- Q (I - V T V^T) on the rank's rows
                deferred_right_update(qrows, m, w, Q + (size_t)(i + 1) * ldQ, ldQ, V, VT, ld, ws.W);
            if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 4), st));
            if (hook) hook->panel_done(st, i + w, !backward);
        }
        if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 1), st));
        trace("all panels enqueued", panel);
        if (backward) {
            // Q = H_0 ( H_1 ( ... H_K-1 I)): panel k acts on rows >= i + 1 and, the product so far being the identity
            // outside its trailing block, on columns >= i + 1 only:  Qb <- (I - V T V^T) Qb = Qb - VT (Qb^T V)^T
            // One GPU: in place in Q. Several: on the rank's columns Qc (n x nloc in the exchange arena, identity at first).
            double *Qc = Q;
            int ldc = ldQ;
            const int nloc = cm.lower(n);
            if (P > 1) {
                Qc = at<double>(g, al.off_qc); ldc = al.ldv;
                if (nloc > 0) {
                    SB_LAUNCH(k_identity_cols, dim3(ceil_div(n, 1024), std::min(nloc, 16384)), 256, 0, st, cm, n, nloc, Qc, ldc);
                    stats.kernel_launches++;
                }
            }
            for (int k = panel - 1; k >= 0; k--) {
                const int i = begin + k * nb, w = std::min(nb, end - i - 1), m = end - i - 1, par = (i + 1) & 1;
                const double *V = ws.Vh + (size_t)i * ld + par, *VT = ws.VTh + (size_t)i * ld + par;
                const int cl0 = cm.lower(i + 1), ncl = nloc - cl0;          // local columns of the trailing block
                if (ncl <= 0) continue;
                double *Qb = Qc + (size_t)cl0 * ldc + i + 1;
                gemm(GEMM_TN, ncl, w, m, 1.0, Qb, ldc, V, ld, 0.0, ws.W, ld, true);
                gemm(GEMM_NT, m, ncl, w, -1.0, VT, ld, ws.W, ld, 1.0, Qb, ldc);
            }
            if (P > 1) {
                PeerPtrs qcp;
                for (int s = 0; s < MAX_RANKS; s++) qcp.p[s] = s < P ? at<double>(s, al.off_qc) : nullptr;
                barrier();                      // every rank's columns are complete
                if (qrows > 0) {
                    SB_LAUNCH(k_qcols_to_rows, dim3(ceil_div(qrows, 256), std::min(n, 16384)), 256, 0, st, cm, n, q0, qrows, qcp, al.ldv, Q, ldQ);
                    stats.kernel_launches++;
                }
                // (the barrier that ends the reduction keeps the columns alive until every peer has pulled its rows)
            }
        }
        if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel + 2), st));
        barrier();
        if (lvl >= 1) SB_CUDA(cudaEventRecord(phase_event(2 + 6 * panel), st));       // end of the critical path
        SB_CUDA(cudaEventRecord(ev_last, st));
        trace("everything enqueued", panel);
        SB_CUDA(cudaStreamSynchronize(st));
        trace("stream idle", panel);
        SB_CUDA(cudaGetLastError());
        if (P > 1) {
            unsigned status = 0;
            SB_CUDA(cudaMemcpy(&status, at<unsigned>(g, al.off_status), sizeof(status), cudaMemcpyDeviceToHost));
            if (status != 0)
                fatal(status == 2 ? "a cross-GPU wait for GEMV partial sums timed out (peer rank missing or stalled)"
                                  : "a cross-GPU barrier timed out (peer rank missing or stalled)", __FILE__, __LINE__);
        }

        stats.panels = panel;
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, ev_first, ev_last));
        stats.device_ms = ms;
        stats.overlap = 0;
        if (lvl >= 1) {
            for (int p = 0; p < panel; p++) {
                cudaEvent_t *e = &events[2 + 6 * p];
                SB_CUDA(cudaEventElapsedTime(&ms, e[0], e[1])); stats.panel_ms += ms;
                SB_CUDA(cudaEventElapsedTime(&ms, e[1], e[2])); stats.trail_ms += ms;
                // the updates the reference defers (rows above the panel, columns right of the block, Q)
                SB_CUDA(cudaEventElapsedTime(&ms, e[3], e[4])); stats.other_ms += ms;
            }
            // backward accumulation of Q after the last panel (part of the deferred work)
            SB_CUDA(cudaEventElapsedTime(&ms, events[2 + 6 * panel + 1], events[2 + 6 * panel + 2])); stats.q_backward_ms = ms; stats.other_ms += ms;
            // what the deferred updates add to the critical path: end of the last trailing update -> end of the call
            SB_CUDA(cudaEventElapsedTime(&ms, events[2 + 6 * panel], ev_last)); stats.side_tail_ms = ms;
        }
        if (stats.fused_panels > 0) {
            unsigned long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            SB_CUDA(cudaMemcpy(t, ws.timers, sizeof(t), cudaMemcpyDeviceToHost));
            stats.gemv_ms = 1e-6 * (double)t[0];            // %globaltimer around the GEMV phases (incl. their barrier)
            stats.fused_kernel_ms = 1e-6 * (double)t[1];
            for (int k = 0; k < 4; k++) stats.fused_phase_ms[k] = 1e-6 * (double)t[2 + k];
        } else if (lvl >= 2) {
            for (size_t k = 0; k + 3 < gemv_events_used; k += 4) {
                SB_CUDA(cudaEventElapsedTime(&ms, gemv_events[k], gemv_events[k + 1])); stats.finish_update_ms += ms;
                SB_CUDA(cudaEventElapsedTime(&ms, gemv_events[k + 1], gemv_events[k + 2])); stats.reflector_ms += ms;
                SB_CUDA(cudaEventElapsedTime(&ms, gemv_events[k + 2], gemv_events[k + 3])); stats.gemv_ms += ms;
            }
        }
    }
};

// The "automatic" panel width (conf->panel_width == STARNEIG_HESSENBERG_DEFAULT_PANEL_WIDTH). The reference fits a width to
// its CPU codelets (src/hessenberg/interface.c:74-78: 0.0019 n + 274, i.e. 312 at n = 20000). Here the trade is another one:
// the level-2 phases of a column stream V, Y, VT of the panel from L2 and grow with the width (and are replicated on every
// rank), the TMA-fed DMMA kernels lose little at a smaller K, and the skinny products W = A^T VT / X VT run on 96- or
// 104-column tiles, so that widths of 192, 208, 288, 312 waste no tile columns while 160 or 256 do. Measured on one B200
// (profiles/r2_v16_sweep_panel_width_1gpu.txt, n = 1000 ... 30000): 192 beats the reference formula by 1.4-2.8 % at every
// size (208 within 0.5 % of it; 96 loses at n >= 10000: +76 ms of level-3 time at n = 20000); on 4 and 8 GPUs 192 was the
// winner already (profiles/r2_v5_visit8b_panelwidth_by_gpus_overlap.log: -3.7 % / -6 %, flat between 96 and 192).
// An explicit conf->panel_width is always taken as given.
static inline int default_panel_width(int n, int P = 1)
{
    // tuning aid (tools/sweep.py): another "automatic" width without touching the caller's configuration
    const char *e = getenv("STARNEIG_B200_AUTO_PANEL_WIDTH");
    if (e && atoi(e) >= 8) return atoi(e);
    (void)n; (void)P;
    return 192;
}

} // namespace sb200
