// hessenberg.cu -- host driver of the B200-native blocked Hessenberg reduction and its C ABI.
//
// Replaces, for the path behind starneig_SEP_SM_Hessenberg (reference src/hessenberg/interface.c),
// the StarPU task graph of src/hessenberg/core.c:351-599 and the tile plumbing of src/common
// (matrix.c, vector.c, tiles.c, scratch.c): here the matrix stays dense and column-major in HBM,
// the "task graph" is a fixed sequence of kernel launches on CUDA streams, and all workspace comes
// from one arena owned by the node context.
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
#include "panel.cuh"
#include "dgemm.cuh"
#include <starneig_b200.h>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// GEMM dispatch
// ---------------------------------------------------------------------------------------------
// <A K-major, B K-major, warps along M, warps along N, 8-row blocks per warp, 8-col blocks per warp, stages, CTAs/SM>
using GemmNT   = GemmConfig<false, false, 2, 2, 8, 4, 4, 2>;     // 128 x  64, 128 threads: rank-nb updates
using GemmTN13 = GemmConfig<true,  true,  4, 1, 2, 13, 4, 2>;    //  64 x 104, W = A^T VT
using GemmTN12 = GemmConfig<true,  true,  4, 1, 2, 12, 4, 2>;    //  64 x  96
using GemmNN13 = GemmConfig<false, true,  4, 1, 2, 13, 4, 2>;    //  64 x 104, W = A VT
using GemmNN12 = GemmConfig<false, true,  4, 1, 2, 12, 4, 2>;    //  64 x  96

static void panel_prepare();
static bool g_gemm_prepared = false;
static void gemm_prepare()
{
    if (g_gemm_prepared) return;
    GemmNT::prepare(); GemmTN13::prepare(); GemmTN12::prepare(); GemmNN13::prepare(); GemmNN12::prepare();
    g_gemm_prepared = true;
}

struct Stats : starneig_b200_stats {};

struct Workspace {
    int n_cap = 0, nb_cap = 0;
    int ldv = 0, nbp = 0;
    double *V = nullptr, *Y = nullptr, *VT = nullptr, *W = nullptr, *Wpart = nullptr;
    size_t wpart_cap = 0;           // doubles
    double *pcol = nullptr, *ypart = nullptr;
    size_t ypart_cap = 0;           // doubles
    double *s = nullptr, *w2 = nullptr, *colpart = nullptr, *sqpart = nullptr;
    ColScal *scal = nullptr;
    unsigned *counter = nullptr;
    std::vector<void *> allocs;

    template <typename T> T *alloc(size_t count)
    {
        void *p = nullptr;
        SB_CUDA(cudaMalloc(&p, count * sizeof(T) + 256));
        allocs.push_back(p);
        return (T *)p;
    }
    void release()
    {
        for (void *p : allocs) cudaFree(p);
        allocs.clear();
        n_cap = nb_cap = 0;
    }
    void ensure(int n, int nb)
    {
        if (n <= n_cap && nb <= nb_cap) return;
        release();
        n_cap = n; nb_cap = nb;
        ldv = round_up(n, 16);
        nbp = round_up(nb, 8);
        size_t panel = (size_t)ldv * nbp;
        V = alloc<double>(panel); Y = alloc<double>(panel); VT = alloc<double>(panel); W = alloc<double>(panel);
        wpart_cap = 8 * (size_t)std::max(ldv, 4096) * nbp;
        Wpart = alloc<double>(wpart_cap);
        pcol = alloc<double>(ldv);
        ypart_cap = (size_t)2 * 148 * 12 * 256 + 4 * (size_t)ldv;
        ypart = alloc<double>(ypart_cap);
        s = alloc<double>(nbp); w2 = alloc<double>(nbp);
        colpart = alloc<double>((size_t)nbp * PANEL_LDB);
        sqpart = alloc<double>(PANEL_LDB);
        scal = alloc<ColScal>(nbp);
        counter = alloc<unsigned>(4);
        SB_CUDA(cudaMemset(counter, 0, 4 * sizeof(unsigned)));
    }
};

struct Context {
    bool ready = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    Workspace ws;
    Stats stats{};
    int profile_level = 1;
    std::vector<cudaEvent_t> events;        // phase events: 4 per panel
    std::vector<cudaEvent_t> gemv_events;   // 2 per column (profile level 2)
    size_t gemv_events_used = 0;

    void open()
    {
        if (ready) return;
        SB_CUDA(cudaGetDevice(&device));
        SB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        gemm_prepare();
        panel_prepare();
        ready = true;
    }
    void close()
    {
        if (!ready) return;
        cudaDeviceSynchronize();
        ws.release();
        for (auto e : events) cudaEventDestroy(e);
        for (auto e : gemv_events) cudaEventDestroy(e);
        events.clear(); gemv_events.clear();
        cudaStreamDestroy(stream);
        stream = nullptr;
        ready = false;
    }
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
};

static Context g_ctx;

enum GemmKind { GEMM_NT, GEMM_TN, GEMM_NN };

// C = alpha*op(A)*op(B) + beta*C on `st`. Wpart/wpart_cap: split-K scratch (may be null => no split).
static void gemm(Context &ctx, cudaStream_t st, GemmKind kind, int M, int N, int K, double alpha, const double *A, int lda,
                 const double *B, int ldb, double beta, double *C, int ldc)
{
    if (M < 1 || N < 1) return;
    ctx.stats.gemm_flops += 2.0 * M * N * (double)K;
    if (kind == GEMM_NT) {
        GemmNT::launch(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 1, K, 0);
        ctx.stats.kernel_launches++;
        return;
    }
    // skinny output (N = panel width): pick the column tile with the least padding, split K if the
    // grid would not fill the GPU twice
    int bn = (ceil_div(N, 96) * 96 <= ceil_div(N, 104) * 104) ? 96 : 104;
    // 2 CTAs per SM are resident; split K so that the grid is >= ~8 waves (tail quantisation < ~6 %)
    int tiles = ceil_div(M, 64) * ceil_div(N, bn);
    int splits = 1;
    const int want = 8 * 2 * 148;
    if (beta == 0.0 && alpha == 1.0 && ctx.ws.Wpart != nullptr && tiles < want) {
        splits = std::min(32, ceil_div(want, tiles));
        splits = std::min(splits, std::max(1, K / 512));
        while (splits > 1 && (size_t)splits * ldc * N > ctx.ws.wpart_cap) splits--;
    }
    int klen = round_up(ceil_div(K, splits), GEMM_BK);
    splits = ceil_div(K, klen);
    double *out = splits > 1 ? ctx.ws.Wpart : C;
    size_t stride = splits > 1 ? (size_t)ldc * N : 0;
    double b = splits > 1 ? 0.0 : beta;
    if (kind == GEMM_TN) {
        if (bn == 96) GemmTN12::launch(st, M, N, K, alpha, A, lda, B, ldb, b, out, ldc, splits, klen, stride);
        else          GemmTN13::launch(st, M, N, K, alpha, A, lda, B, ldb, b, out, ldc, splits, klen, stride);
    } else {
        if (bn == 96) GemmNN12::launch(st, M, N, K, alpha, A, lda, B, ldb, b, out, ldc, splits, klen, stride);
        else          GemmNN13::launch(st, M, N, K, alpha, A, lda, B, ldb, b, out, ldc, splits, klen, stride);
    }
    ctx.stats.kernel_launches++;
    if (splits > 1) {
        dim3 grid(ceil_div(M, 256), N);
        splitk_reduce_kernel<<<grid, 256, 0, st>>>(M, N, splits, out, ldc, stride, C, ldc);
        ctx.stats.kernel_launches++;
    }
}

// ---------------------------------------------------------------------------------------------
// panel factorisation: columns i .. i+w-1
// ---------------------------------------------------------------------------------------------
static PanelArgs make_panel_args(Workspace &ws, int m, double *V, double *Y, double *VT, int ld)
{
    PanelArgs pa;
    pa.m = m; pa.ld = ld; pa.V = V; pa.Y = Y; pa.VT = VT;
    pa.pcol = ws.pcol; pa.ypart = ws.ypart; pa.ldp = round_up(m + 2, 16);
    pa.s = ws.s; pa.w2 = ws.w2; pa.colpart = ws.colpart; pa.ldt = ws.nbp;
    pa.sqpart = ws.sqpart; pa.scal = ws.scal; pa.counter = ws.counter;
    return pa;
}

struct GemvPlan { int skip, RB, S, kc; const double *A0; };

static int g_gemv_slots = 0;    // resident k_col_gemv blocks on the whole GPU (one wave)

// decomposition of the GEMV over rows [0,m) x columns [0,ncols) starting at `base`: row blocks of 256
// padded rows times S column chunks, sized so that the grid is (at most) one full wave
static GemvPlan plan_gemv(const double *base, int m, int ncols, size_t ypart_cap, int ldp)
{
    if (g_gemv_slots == 0) {
        int occ = 0, dev = 0, sms = 0;
        SB_CUDA(cudaGetDevice(&dev));
        SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_col_gemv, GEMV_THREADS, 2048 * sizeof(double)));
        g_gemv_slots = std::max(1, occ) * sms;
    }
    GemvPlan p;
    p.skip = (int)(((uintptr_t)base / sizeof(double)) & 1);
    p.A0 = base - p.skip;
    int mp = m + p.skip;
    p.RB = ceil_div(mp, 256);
    int S = std::max(1, g_gemv_slots / p.RB);
    int kc = ceil_div(ncols, S);
    kc = std::max(kc, 16);
    kc = std::min(round_up(kc, 4), 2048);
    S = ceil_div(ncols, kc);
    while ((size_t)S * ldp > ypart_cap && kc < 2048) { kc *= 2; S = ceil_div(ncols, kc); }
    p.kc = kc; p.S = S;
    return p;
}

struct PanelGrid { int blocks; TileGeom tg; size_t smem_fu, smem_rf; };
static const size_t PANEL_SMEM_MAX = 200 * 1024;

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
    g.smem_rf = (size_t)(j + tg.nsub * tg.NW * 32 + tg.nsub * 32 + tg.RS * tg.NW * 32 + 32) * sizeof(double);
    if (g.smem_fu > PANEL_SMEM_MAX) fatal("matrix too large for the panel kernels' shared-memory layout", __FILE__, __LINE__);
    return g;
}

static void panel_prepare()
{
    static bool done = false;
    if (done) return;
    SB_CUDA(cudaFuncSetAttribute(k_col_finish_update<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    SB_CUDA(cudaFuncSetAttribute(k_col_finish_update<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    SB_CUDA(cudaFuncSetAttribute(k_col_reflector<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    SB_CUDA(cudaFuncSetAttribute(k_col_reflector<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_MAX));
    done = true;
}

static void launch_finish_update(cudaStream_t st, const PanelArgs &pa, int j, int S, double *acol, int do_update)
{
    PanelGrid pg = panel_grid(pa.m, j, j);
    const int threads = 32 * pg.tg.NW * pg.tg.RS;
    if (threads <= 512) k_col_finish_update<512><<<pg.blocks, threads, pg.smem_fu, st>>>(pa, j, S, acol, do_update, pg.tg);
    else                k_col_finish_update<1024><<<pg.blocks, threads, pg.smem_fu, st>>>(pa, j, S, acol, do_update, pg.tg);
}

static void launch_reflector(cudaStream_t st, const PanelArgs &pa, int j, double *acol)
{
    PanelGrid pg = panel_grid(pa.m, j, j);
    const int threads = 32 * pg.tg.NW * pg.tg.RS;
    if (threads <= 512) k_col_reflector<512><<<pg.blocks, threads, pg.smem_rf, st>>>(pa, j, acol, pg.tg);
    else                k_col_reflector<1024><<<pg.blocks, threads, pg.smem_rf, st>>>(pa, j, acol, pg.tg);
}

static void panel_factor(Context &ctx, cudaStream_t st, int i, int end, int w, double *A, int ldA,
                         double *V, double *Y, double *VT, int ld)
{
    Workspace &ws = ctx.ws;
    const int m = end - i - 1;
    PanelArgs pa = make_panel_args(ws, m, V, Y, VT, ld);
    SB_CUDA(cudaMemsetAsync(V, 0, (size_t)ld * w * sizeof(double), st));
    int S_prev = 0;
    for (int j = 0; j < w; j++) {
        const int c = i + j;
        double *acol = A + (size_t)c * ldA + i + 1;
        const bool timed = ctx.profile_level >= 3 || (ctx.profile_level == 2 && (j & 7) == 4);
        if (timed) SB_CUDA(cudaEventRecord(ctx.gemv_event(ctx.gemv_events_used++), st));
        if (j > 0) {
            launch_finish_update(st, pa, j, S_prev, acol, 1);
            ctx.stats.kernel_launches++;
        }
        if (timed) SB_CUDA(cudaEventRecord(ctx.gemv_event(ctx.gemv_events_used++), st));
        launch_reflector(st, pa, j, acol);
        ctx.stats.kernel_launches++;
        if (timed) SB_CUDA(cudaEventRecord(ctx.gemv_event(ctx.gemv_events_used++), st));
        {
            const int ncols = m - j;
            const double *base = A + (size_t)(c + 1) * ldA + i + 1;
            GemvPlan gp = plan_gemv(base, m, ncols, ws.ypart_cap, pa.ldp);
            size_t sh = (size_t)gp.kc * sizeof(double);
            k_col_gemv<<<gp.RB * gp.S, GEMV_THREADS, sh, st>>>(pa, j, ncols, gp.A0, ldA, gp.skip, gp.kc, gp.RB, acol);
            if (timed) {
                SB_CUDA(cudaEventRecord(ctx.gemv_event(ctx.gemv_events_used++), st));
                ctx.stats.gemv_timed_launches++;
                ctx.stats.gemv_timed_bytes += 8.0 * (double)m * ncols;
            }
            ctx.stats.kernel_launches++;
            ctx.stats.gemv_launches++;
            ctx.stats.gemv_bytes += 8.0 * (double)m * ncols;
            S_prev = gp.S;
        }
    }
    // finish the last column (Y, VT) without starting a new one
    launch_finish_update(st, pa, w, S_prev, nullptr, 0);
    ctx.stats.kernel_launches++;
}

// ---------------------------------------------------------------------------------------------
// the whole reduction on device-resident A, Q
// ---------------------------------------------------------------------------------------------
static void reduce_device(Context &ctx, int n, int begin, int end, int nb, double *A, int ldA, double *Q, int ldQ)
{
    Workspace &ws = ctx.ws;
    cudaStream_t st = ctx.stream;
    nb = std::min(nb, PANEL_MAX_NB);      // wider panels are split; the result only differs in rounding
    ws.ensure(n, nb);
    const int ld = ws.ldv;
    Stats &stt = ctx.stats;
    const int lvl = ctx.profile_level;
    ctx.gemv_events_used = 0;
    int panel = 0;
    cudaEvent_t ev_first = ctx.phase_event(0), ev_last = ctx.phase_event(1);
    SB_CUDA(cudaEventRecord(ev_first, st));

    for (int i = begin; i < end - 1; i += nb, panel++) {
        const int w = std::min(nb, end - i - 1);
        const int m = end - i - 1;
        if (lvl >= 1) SB_CUDA(cudaEventRecord(ctx.phase_event(2 + 4 * panel + 0), st));
        panel_factor(ctx, st, i, end, w, A, ldA, ws.V, ws.Y, ws.VT, ld);
        if (lvl >= 1) SB_CUDA(cudaEventRecord(ctx.phase_event(2 + 4 * panel + 1), st));

        const int ntr = end - (i + w);
        if (ntr > 0) {
            double *Atr = A + (size_t)(i + w) * ldA + i + 1;
            gemm(ctx, st, GEMM_NT, m, ntr, w, -1.0, ws.Y, ld, ws.V + (w - 1), ld, 1.0, Atr, ldA);
            gemm(ctx, st, GEMM_TN, ntr, w, m, 1.0, Atr, ldA, ws.VT, ld, 0.0, ws.W, ld);
            gemm(ctx, st, GEMM_NT, m, ntr, w, -1.0, ws.V, ld, ws.W, ld, 1.0, Atr, ldA);
        }
        if (lvl >= 1) SB_CUDA(cudaEventRecord(ctx.phase_event(2 + 4 * panel + 2), st));

        {   // rows above the panel
            double *X = A + (size_t)(i + 1) * ldA;
            gemm(ctx, st, GEMM_NN, i + 1, w, m, 1.0, X, ldA, ws.VT, ld, 0.0, ws.W, ld);
            gemm(ctx, st, GEMM_NT, i + 1, m, w, -1.0, ws.W, ld, ws.V, ld, 1.0, X, ldA);
        }
        if (end < n) {   // columns right of the reduced block (partial reduction)
            double *X = A + (size_t)end * ldA + i + 1;
            gemm(ctx, st, GEMM_TN, n - end, w, m, 1.0, X, ldA, ws.VT, ld, 0.0, ws.W, ld);
            gemm(ctx, st, GEMM_NT, m, n - end, w, -1.0, ws.V, ld, ws.W, ld, 1.0, X, ldA);
        }
        {   // Q <- Q (I - V T V^T)
            double *X = Q + (size_t)(i + 1) * ldQ;
            gemm(ctx, st, GEMM_NN, n, w, m, 1.0, X, ldQ, ws.VT, ld, 0.0, ws.W, ld);
            gemm(ctx, st, GEMM_NT, n, m, w, -1.0, ws.W, ld, ws.V, ld, 1.0, X, ldQ);
        }
        if (lvl >= 1) SB_CUDA(cudaEventRecord(ctx.phase_event(2 + 4 * panel + 3), st));
    }
    SB_CUDA(cudaEventRecord(ev_last, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CUDA(cudaGetLastError());

    stt.panels = panel;
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, ev_first, ev_last));
    stt.device_ms = ms;
    if (lvl >= 1) {
        for (int p = 0; p < panel; p++) {
            cudaEvent_t *e = &ctx.events[2 + 4 * p];
            SB_CUDA(cudaEventElapsedTime(&ms, e[0], e[1])); stt.panel_ms += ms;
            SB_CUDA(cudaEventElapsedTime(&ms, e[1], e[2])); stt.trail_ms += ms;
            SB_CUDA(cudaEventElapsedTime(&ms, e[2], e[3])); stt.other_ms += ms;
        }
    }
    if (lvl >= 2) {
        for (size_t k = 0; k + 3 < ctx.gemv_events_used; k += 4) {
            SB_CUDA(cudaEventElapsedTime(&ms, ctx.gemv_events[k], ctx.gemv_events[k + 1])); stt.finish_update_ms += ms;
            SB_CUDA(cudaEventElapsedTime(&ms, ctx.gemv_events[k + 1], ctx.gemv_events[k + 2])); stt.reflector_ms += ms;
            SB_CUDA(cudaEventElapsedTime(&ms, ctx.gemv_events[k + 2], ctx.gemv_events[k + 3])); stt.gemv_ms += ms;
        }
    }
}

static int default_panel_width(int n)
{
    // reference src/hessenberg/interface.c:74-78
    int w = (int)std::ceil((0.001875596476 * n + 273.5908216) / 8.0) * 8;
    return std::max(64, w);
}

static double wall_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

} // namespace sb200

// =================================================================================================
// node state (reference src/common/node.c) -- implemented in node.cpp
// =================================================================================================
extern "C" int starneig_b200_node_messages_enabled(void);
extern "C" int starneig_b200_node_pinning_enabled(void);

// The Hessenberg path is CUDA-only: without a selected GPU the call fails loudly instead of falling
// back to a CPU implementation.
static bool have_gpu()
{
    if (starneig_node_get_gpus() >= 1) return true;
    fprintf(stderr, "[starneig][error] No CUDA device selected/available: the Hessenberg path has no CPU "
                    "implementation. Exiting...\n");
    return false;
}

using namespace sb200;

extern "C" void starneig_b200_context_open(void) { g_ctx.open(); }
extern "C" void starneig_b200_context_close(void) { g_ctx.close(); }

extern "C" __attribute__((visibility("default")))
void starneig_b200_get_stats(struct starneig_b200_stats *stats) { *stats = g_ctx.stats; }

extern "C" __attribute__((visibility("default")))
void starneig_b200_set_profile_level(int level) { g_ctx.profile_level = level; }

extern "C" __attribute__((visibility("default")))
void starneig_hessenberg_init_conf(struct starneig_hessenberg_conf *conf)
{
    conf->tile_size = STARNEIG_HESSENBERG_DEFAULT_TILE_SIZE;
    conf->panel_width = STARNEIG_HESSENBERG_DEFAULT_PANEL_WIDTH;
}

// configuration checks of reference hessenberg(), src/hessenberg/interface.c:62-84
static starneig_error_t resolve_conf(struct starneig_hessenberg_conf const *conf, int n, int *panel_width)
{
    struct starneig_hessenberg_conf local;
    if (conf == NULL) starneig_hessenberg_init_conf(&local);
    else local = *conf;
    if (local.tile_size != STARNEIG_HESSENBERG_DEFAULT_TILE_SIZE && local.tile_size < 8) {
        fprintf(stderr, "[starneig][error] Invalid tile size. Exiting...\n");
        return STARNEIG_INVALID_CONFIGURATION;
    }
    if (local.panel_width == STARNEIG_HESSENBERG_DEFAULT_PANEL_WIDTH) {
        local.panel_width = default_panel_width(n);
        if (starneig_b200_node_messages_enabled())
            printf("[starneig][message] Setting panel width to %d.\n", local.panel_width);
    } else if (local.panel_width < 8) {
        fprintf(stderr, "[starneig][error] Invalid panel width. Exiting...\n");
        return STARNEIG_INVALID_CONFIGURATION;
    }
    *panel_width = local.panel_width;
    return STARNEIG_SUCCESS;
}

static void reset_stats(int n, int begin, int end, int nb)
{
    memset(&g_ctx.stats, 0, sizeof(g_ctx.stats));
    g_ctx.stats.n = n; g_ctx.stats.begin = begin; g_ctx.stats.end = end; g_ctx.stats.panel_width = nb;
}

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_b200_hessenberg_device(int n, int begin, int end, int panel_width,
                                                 double *dA, int ldA, double *dQ, int ldQ)
{
    if (n < 1) return -1;
    if (begin < 0) return -2;
    if (n < end) return -3;
    if (dA == NULL) return -5;
    if (ldA < n || (ldA & 1) || ((uintptr_t)dA & 15)) return -6;
    if (dQ == NULL) return -7;
    if (ldQ < n || (ldQ & 1) || ((uintptr_t)dQ & 15)) return -8;
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (panel_width < 0) panel_width = default_panel_width(n);
    if (panel_width < 8) return STARNEIG_INVALID_CONFIGURATION;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    g_ctx.open();
    reset_stats(n, begin, end, panel_width);
    double t0 = wall_ms();
    reduce_device(g_ctx, n, begin, end, panel_width, dA, ldA, dQ, ldQ);
    g_ctx.stats.wall_ms = wall_ms() - t0;
    return STARNEIG_SUCCESS;
}

// device staging buffers for the host API
namespace {
struct Staging {
    double *dA = nullptr, *dQ = nullptr;
    size_t cap = 0;     // doubles per matrix
    int ldd = 0;
    void ensure(int n)
    {
        ldd = round_up(n, 16);
        size_t need = (size_t)ldd * n + 64;
        if (need <= cap) return;
        release();
        SB_CUDA(cudaMalloc(&dA, need * sizeof(double)));
        SB_CUDA(cudaMalloc(&dQ, need * sizeof(double)));
        cap = need;
    }
    void release()
    {
        if (dA) cudaFree(dA);
        if (dQ) cudaFree(dQ);
        dA = dQ = nullptr; cap = 0;
    }
} g_staging;

// page-locks a caller buffer for the duration of a call (no-op if it already is pinned or on failure)
struct ScopedPin {
    void *p = nullptr;
    ScopedPin(void *ptr, size_t bytes, bool enable)
    {
        if (!enable) return;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) return;
        cudaGetLastError();
        if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess) p = ptr;
        else cudaGetLastError();
    }
    ~ScopedPin() { if (p) cudaHostUnregister(p); }
};
}

extern "C" void starneig_b200_staging_release(void) { g_staging.release(); }

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_SEP_SM_Hessenberg_expert(struct starneig_hessenberg_conf *conf, int n, int begin, int end,
                                                   double A[], int ldA, double Q[], int ldQ)
{
    // argument checks: reference src/hessenberg/interface.c:144-153
    if (n < 1) return -2;
    if (begin < 0) return -3;
    if (n < end) return -4;
    if (A == NULL) return -5;
    if (ldA < n) return -6;
    if (Q == NULL) return -7;
    if (ldQ < n) return -8;
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;

    int nb = 0;
    starneig_error_t ret = resolve_conf(conf, n, &nb);
    if (ret != STARNEIG_SUCCESS) return ret;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;

    g_ctx.open();
    reset_stats(n, begin, end, nb);
    double t0 = wall_ms();
    g_staging.ensure(n);
    const int ldd = g_staging.ldd;
    cudaStream_t st = g_ctx.stream;
    const size_t rowbytes = (size_t)n * sizeof(double);
    {
        ScopedPin pinA(A, ((size_t)ldA * (n - 1) + n) * sizeof(double), starneig_b200_node_pinning_enabled());
        ScopedPin pinQ(Q, ((size_t)ldQ * (n - 1) + n) * sizeof(double), starneig_b200_node_pinning_enabled());
        double t1 = wall_ms();
        SB_CUDA(cudaMemcpy2DAsync(g_staging.dA, (size_t)ldd * 8, A, (size_t)ldA * 8, rowbytes, n, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaMemcpy2DAsync(g_staging.dQ, (size_t)ldd * 8, Q, (size_t)ldQ * 8, rowbytes, n, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaStreamSynchronize(st));
        double t2 = wall_ms();
        reduce_device(g_ctx, n, begin, end, nb, g_staging.dA, ldd, g_staging.dQ, ldd);
        double t3 = wall_ms();
        SB_CUDA(cudaMemcpy2DAsync(A, (size_t)ldA * 8, g_staging.dA, (size_t)ldd * 8, rowbytes, n, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaMemcpy2DAsync(Q, (size_t)ldQ * 8, g_staging.dQ, (size_t)ldd * 8, rowbytes, n, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        double t4 = wall_ms();
        g_ctx.stats.h2d_ms = t2 - t1;
        g_ctx.stats.d2h_ms = t4 - t3;
        g_ctx.stats.h2d_bytes = 2 * (long long)rowbytes * n;
        g_ctx.stats.d2h_bytes = 2 * (long long)rowbytes * n;
    }
    g_ctx.stats.wall_ms = wall_ms() - t0;
    return STARNEIG_SUCCESS;
}

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_SEP_SM_Hessenberg(int n, double A[], int ldA, double Q[], int ldQ)
{
    // reference src/hessenberg/interface.c:175-184
    if (n < 1) return -1;
    if (A == NULL) return -2;
    if (ldA < n) return -3;
    if (Q == NULL) return -4;
    if (ldQ < n) return -5;
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    return starneig_SEP_SM_Hessenberg_expert(NULL, n, 0, n, A, ldA, Q, ldQ);
}

// ---------------------------------------------------------------------------------------------
// unit-level entry points
// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default")))
int starneig_b200_dgemm(char transa, char transb, int m, int n, int k, double alpha, const double *dA, int lda,
                        const double *dB, int ldb, double beta, double *dC, int ldc)
{
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    g_ctx.open();
    GemmKind kind;
    if (transa == 'N' && transb == 'T') kind = GEMM_NT;
    else if (transa == 'T' && transb == 'N') kind = GEMM_TN;
    else if (transa == 'N' && transb == 'N') kind = GEMM_NN;
    else return STARNEIG_INVALID_ARGUMENTS;
    if (m < 1 || n < 1 || k < 1) return STARNEIG_INVALID_ARGUMENTS;
    if (kind != GEMM_NT) g_ctx.ws.ensure(std::max(m, 16), std::max(n, 8));
    gemm(g_ctx, g_ctx.stream, kind, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc);
    SB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    SB_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_set_gemv_inputs(ColScal *scal, double *pcol, const double *v, int k)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { scal[0].tau = 0.0; scal[0].beta = 0.0; scal[0].scale = 1.0; scal[0].alpha = 0.0; }
    if (t < k) pcol[t] = v[t];
}

__global__ void k_sum_partials(int m, int S, const double *ypart, int ldp, double *y)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double s = 0.0;
    for (int z = 0; z < S; z++) s += ypart[(size_t)z * ldp + r];
    y[r] = s;
}

extern "C" __attribute__((visibility("default")))
int starneig_b200_gemv(int m, int k, const double *dA, int lda, const double *dv, double *dy, int reps, float *mean_ms)
{
    // Runs k_col_gemv as column j = 0 of a fictitious panel with pcol = v and scale = 1; the kernel
    // forms its vector as (1, scale * pcol[1:]), so v[0] must be 1 for an exact match.
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (m < 1 || k < 1 || (lda & 1)) return STARNEIG_INVALID_ARGUMENTS;
    g_ctx.open();
    Workspace &ws = g_ctx.ws;
    const int big = std::max(std::max(m, k), 16);
    ws.ensure(big, 8);
    cudaStream_t st = g_ctx.stream;
    double *scratch_col = nullptr;
    SB_CUDA(cudaMalloc(&scratch_col, (size_t)(big + 16) * sizeof(double)));
    PanelArgs pa = make_panel_args(ws, m, ws.V, ws.Y, ws.VT, ws.ldv);
    k_set_gemv_inputs<<<ceil_div(k, 256), 256, 0, st>>>(ws.scal, ws.pcol, dv, k);
    GemvPlan gp = plan_gemv(dA, m, k, ws.ypart_cap, pa.ldp);
    size_t sh = (size_t)gp.kc * sizeof(double);
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    if (reps < 1) reps = 1;
    for (int it = 0; it < reps + 1; it++) {
        if (it == 1) SB_CUDA(cudaEventRecord(e0, st));
        k_col_gemv<<<gp.RB * gp.S, GEMV_THREADS, sh, st>>>(pa, 0, k, gp.A0, lda, gp.skip, gp.kc, gp.RB, scratch_col);
    }
    SB_CUDA(cudaEventRecord(e1, st));
    k_sum_partials<<<ceil_div(m, 256), 256, 0, st>>>(m, gp.S, ws.ypart, pa.ldp, dy);
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CUDA(cudaGetLastError());
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (mean_ms) *mean_ms = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(scratch_col);
    return 0;
}

extern "C" __attribute__((visibility("default")))
int starneig_b200_panel(int n, int i, int end, int w, double *dA, int ldA, double *dV, double *dY, double *dVT, int ldw,
                        double *htau)
{
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (n < 1 || i < 0 || end > n || w < 1 || w > end - i - 1 || (ldA & 1) || ((uintptr_t)dA & 15)) return STARNEIG_INVALID_ARGUMENTS;
    g_ctx.open();
    g_ctx.ws.ensure(n, std::max(w, 8));
    panel_factor(g_ctx, g_ctx.stream, i, end, w, dA, ldA, dV, dY, dVT, ldw);
    SB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    SB_CUDA(cudaGetLastError());
    if (htau) {
        std::vector<ColScal> sc(w);
        SB_CUDA(cudaMemcpy(sc.data(), g_ctx.ws.scal, w * sizeof(ColScal), cudaMemcpyDeviceToHost));
        for (int j = 0; j < w; j++) htau[j] = sc[j].tau;
    }
    return 0;
}
