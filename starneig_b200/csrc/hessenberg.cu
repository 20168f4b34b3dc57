// hessenberg.cu -- C ABI of the B200-native blocked Hessenberg reduction and the host code that drives
// the per-GPU engines (engine.cuh).
//
// Boundary (reference file:line of what each entry point replaces is in include/starneig/*.h):
//   starneig_SEP_SM_Hessenberg[_expert]      host pointers in, host pointers out; `gpus` of starneig_node_init
//                                            selects the number of ranks (1, 2, 4, 8 GPUs of one box), each
//                                            driven by its own host thread of this process
//   starneig_b200_hessenberg_device          single GPU, matrices already in HBM
//   starneig_b200_dist_*                     one PROCESS per GPU (torchrun): the ranks exchange one
//                                            cudaIpcMemHandle each (through torch.distributed) and then run the
//                                            same engine; all cross-GPU traffic is NVLink peer stores/loads
//                                            issued by the kernels themselves
// There is no CPU implementation of this path: without a GPU every compute entry point fails loudly.
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include "engine.cuh"

namespace sb200 {

static double wall_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// device shards of one rank for the host-pointer API (owned by the library)
struct Shard {
    double *A = nullptr, *Q = nullptr;
    size_t capA = 0, capQ = 0;      // doubles
    int ldA = 0, ldQ = 0, ncols = 0, q0 = 0, qrows = 0;
    void ensure(const Rank &r, int n)
    {
        const ColMap cm{r.P, r.g, r.P == 1 ? std::max(n, 1) : r.cb};
        ncols = cm.lower(n);
        int q1;
        q_row_range(r.P, r.g, n, &q0, &q1);
        qrows = q1 - q0;
        ldA = round_up(n, 16);
        ldQ = round_up(std::max(qrows, 1), 16);
        size_t needA = (size_t)ldA * std::max(ncols, 1) + 64, needQ = (size_t)ldQ * n + 64;
        if (needA > capA) { if (A) cudaFree(A); SB_CUDA(cudaMalloc(&A, needA * sizeof(double))); capA = needA; }
        if (needQ > capQ) { if (Q) cudaFree(Q); SB_CUDA(cudaMalloc(&Q, needQ * sizeof(double))); capQ = needQ; }
    }
    void release()
    {
        if (A) cudaFree(A);
        if (Q) cudaFree(Q);
        A = Q = nullptr; capA = capQ = 0;
    }
};

// Host <-> device movement of one rank's shards (its block-cyclic columns of A, its row slab of Q), overlapped with
// the reduction where the data dependencies allow it:
//   * A goes up first on the compute stream (the first GEMV reads all of it); Q goes up on the copy stream and
//     the first Q update waits for it, so its transfer hides behind the first column loop;
//   * after panel [i, i+w) global columns [0, i+w) of A and [0, i+w] of Q are final (later panels only touch
//     columns to their right): they travel back on the copy stream while the next panels are factorised, and
//     only the columns of the last panel are still on the device when the reduction ends.
// The copies only overlap when the host buffers are page-locked (caller-pinned or registered for the call by
// ScopedPin); with pageable memory cudaMemcpy2DAsync blocks the calling thread, so everything stays in
// order on the compute stream: upload, reduce, download.
// columns of `width` bytes between pitched buffers; one flat copy when both sides are contiguous (the usual case:
// ld == n on the host and n a multiple of 16), so that the DMA engine sees one transfer instead of one per column
static void copy_columns(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t ncols,
                         cudaMemcpyKind kind, cudaStream_t st)
{
    if (ncols == 0 || width == 0) return;
    if (dpitch == width && spitch == width) SB_CUDA(cudaMemcpyAsync(dst, src, width * ncols, kind, st));
    else SB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, ncols, kind, st));
}

struct HostStage : StageHook {
    const Rank &r; Shard &sh;
    const int n; double *const A; const int ldA; double *const Q; const int ldQ;
    const bool overlap;
    const int cb;
    bool writeback = true;              // false: H and Q stay on the device (starneig_b200_SEP_SM_Hessenberg_stage)
    int a_done = 0, q_done = 0;         // local columns of A / columns of Q already on their way back

    HostStage(const Rank &r_, Shard &sh_, int n_, double *A_, int ldA_, double *Q_, int ldQ_, bool overlap_)
        : r(r_), sh(sh_), n(n_), A(A_), ldA(ldA_), Q(Q_), ldQ(ldQ_), overlap(overlap_), cb(r_.P == 1 ? std::max(n_, 1) : r_.cb) {}

    // local columns [l0, l1) of A: contiguous runs inside one column block
    void copy_a(int l0, int l1, bool to_device, cudaStream_t st)
    {
        const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        while (l0 < l1) {
            const int lb = l0 / cb, off = l0 - lb * cb;
            const int gc = (lb * r.P + r.g) * cb + off;
            const int nc = std::min(std::min(cb - off, l1 - l0), n - gc);
            if (nc <= 0) break;
            double *d = sh.A + (size_t)l0 * sh.ldA, *h = A + (size_t)gc * ldA;
            if (to_device) copy_columns(d, (size_t)sh.ldA * 8, h, (size_t)ldA * 8, (size_t)n * 8, nc, kind, st);
            else           copy_columns(h, (size_t)ldA * 8, d, (size_t)sh.ldA * 8, (size_t)n * 8, nc, kind, st);
            l0 += nc;
        }
    }
    // columns [c0, c1) of the rank's row slab of Q
    void copy_q(int c0, int c1, bool to_device, cudaStream_t st)
    {
        if (sh.qrows <= 0 || c1 <= c0) return;
        const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        double *d = sh.Q + (size_t)c0 * sh.ldQ, *h = Q + (size_t)c0 * ldQ + sh.q0;
        if (to_device) copy_columns(d, (size_t)sh.ldQ * 8, h, (size_t)ldQ * 8, (size_t)sh.qrows * 8, c1 - c0, kind, st);
        else           copy_columns(h, (size_t)ldQ * 8, d, (size_t)sh.ldQ * 8, (size_t)sh.qrows * 8, c1 - c0, kind, st);
    }
    void upload()
    {
        copy_a(0, sh.ncols, true, r.stream);
        if (overlap) {
            // Q follows A over the link (the copy stream waits for A's upload): started together the two uploads
            // would share the host-to-device bandwidth and delay A, which the first GEMV is waiting for, while Q is
            // not needed before the first panel's column loop has finished
            SB_CUDA(cudaEventRecord(r.ev_cols_final, r.stream));
            SB_CUDA(cudaStreamWaitEvent(r.copy, r.ev_cols_final, 0));
        }
        copy_q(0, n, true, overlap ? r.copy : r.stream);
        if (overlap) SB_CUDA(cudaEventRecord(r.ev_q_up, r.copy));
        SB_CUDA(cudaStreamSynchronize(r.stream));
    }
    void before_q(cudaStream_t s) override
    {
        if (overlap) SB_CUDA(cudaStreamWaitEvent(s, r.ev_q_up, 0));
    }
    void send_back(int final_cols, bool q_too, cudaStream_t st)
    {
        const ColMap cm{r.P, r.g, cb};
        const int a1 = std::min(sh.ncols, cm.lower(std::min(final_cols, n)));
        if (a1 > a_done) { copy_a(a_done, a1, false, st); a_done = a1; }
        const int q1 = q_too ? std::min(n, final_cols + 1) : 0;
        if (q1 > q_done) { copy_q(q_done, q1, false, st); q_done = q1; }
    }
    void panel_done(cudaStream_t s, int final_cols, bool q_too) override
    {
        if (!overlap || !writeback) return;
        SB_CUDA(cudaEventRecord(r.ev_cols_final, s));
        SB_CUDA(cudaStreamWaitEvent(r.copy, r.ev_cols_final, 0));
        send_back(final_cols, q_too, r.copy);
    }
    // after the reduction (its streams are idle): whatever is still on the device
    void finish()
    {
        if (!writeback) return;
        cudaStream_t st = overlap ? r.copy : r.stream;
        a_done = std::min(a_done, sh.ncols);
        copy_a(a_done, sh.ncols, false, st); a_done = sh.ncols;
        copy_q(q_done, n, false, st); q_done = n;
        SB_CUDA(cudaStreamSynchronize(st));
    }
};

// the whole buffer [p, p + bytes) is page-locked (caller-pinned or registered by ScopedPin): probed at both ends, because a
// registration that failed half-way (two buffers sharing a page of one malloc arena) leaves the first byte of the second
// buffer inside the first one's registered range
static bool page_locked(const void *p, size_t bytes)
{
    auto probe = [](const void *q) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, q) != cudaSuccess) { cudaGetLastError(); return false; }
        return attr.type == cudaMemoryTypeHost;
    };
    return probe(p) && probe((const char *)p + (bytes > 0 ? bytes - 1 : 0));
}

struct HostTimes { double h2d_ms = 0, d2h_ms = 0; long long h2d_bytes = 0, d2h_bytes = 0; int overlapped = 0; };

// upload, reduce, download on one rank (called on the rank's own host thread)
// `rendezvous` (several ranks): called between the upload and the first cross-GPU barrier kernel, so that a rank whose
// upload takes seconds longer (pageable memory, a loaded host) does not run into the time-out of the device-side waits
static void run_rank_host(Rank &r, Shard &sh, int n, int begin, int end, int nb, double *A, int ldA, double *Q, int ldQ, HostTimes *ht,
                          bool writeback = true, const std::function<void()> *rendezvous = nullptr)
{
    SB_CUDA(cudaSetDevice(r.device));
    const char *e = getenv("STARNEIG_B200_STAGE_OVERLAP");
    const size_t bytesA = ((size_t)ldA * (n - 1) + n) * sizeof(double), bytesQ = ((size_t)ldQ * (n - 1) + n) * sizeof(double);
    const bool overlap = (e ? atoi(e) != 0 : true) && page_locked(A, bytesA) && page_locked(Q, bytesQ);
    HostStage stage(r, sh, n, A, ldA, Q, ldQ, overlap);
    stage.writeback = writeback;
    ht->overlapped = overlap ? 1 : 0;
    double t1 = wall_ms();
    stage.upload();
    if (rendezvous) (*rendezvous)();
    double t2 = wall_ms();
    r.reduce(n, begin, end, nb, sh.A, sh.ldA, sh.Q, sh.ldQ, sh.qrows, &stage);
    double t3 = wall_ms();
    stage.finish();
    double t4 = wall_ms();
    // with overlap: h2d_ms is the upload of A alone (Q hides behind the first column loop) and d2h_ms is what is
    // left of the write-back once the reduction has ended
    ht->h2d_ms = t2 - t1; ht->d2h_ms = t4 - t3;
    ht->h2d_bytes = ht->d2h_bytes = ((long long)sh.ncols * n + (long long)sh.qrows * n) * 8;
}

// the ranks driven by this process: P host threads, one per GPU (P == 1 runs on the caller's thread)
struct Team {
    int P = 0;
    std::vector<Rank *> ranks;
    std::vector<Shard> shards;
    Stats stats{};
    int profile_level = 1;

    void open(int P_)
    {
        if (P == P_ && P_ == 1) {
            // the single-rank engine lives on the device that was current when it was created: a caller that has switched
            // devices since (torch.cuda.set_device) gets a new engine there instead of launches on foreign pointers
            int cur = 0;
            SB_CUDA(cudaGetDevice(&cur));
            if (ranks[0]->device == cur) return;
        } else if (P == P_) return;
        close();
        P = P_;
        int ndev = 0, cur = 0;
        SB_CUDA(cudaGetDeviceCount(&ndev));
        SB_CUDA(cudaGetDevice(&cur));
        for (int g = 0; g < P; g++) {
            Rank *r = new Rank();
            // P == 1: the caller's current device (one process per GPU under torchrun); P > 1: devices 0..P-1
            // (STARNEIG_B200_VIRTUAL_RANKS: several ranks may share a device -- development aid)
            r->open(P, g, P == 1 ? cur : g % ndev);
            ranks.push_back(r);
        }
        shards.resize(P);
        // ranks that share a device (development aid) must share its SMs: the persistent panel kernels of all of
        // them have to be co-resident
        bool shared = false;
        for (int g = 0; g < P; g++) {
            int sharing = 0;
            for (int s = 0; s < P; s++) sharing += ranks[s]->device == ranks[g]->device;
            ranks[g]->fused_ctas = std::max(1, ranks[g]->fused_ctas / sharing);
            shared = shared || sharing > 1;
        }
        // Ranks that share a device accumulate Q forward: the backward order makes every rank's host wait for its stream in the
        // middle of the reduction, and with several host threads on ONE device such waits can hold up the launches of the ranks
        // the stream is waiting for (seen with 8 ranks on 2 GPUs; one rank per GPU is not affected).
        if (shared)
            for (Rank *r : ranks) r->q_backward = 0;
        for (int g = 0; g < P; g++)
            for (int s = 0; s < P; s++)
                if (ranks[g]->device != ranks[s]->device) {
                    SB_CUDA(cudaSetDevice(ranks[g]->device));
                    cudaError_t e = cudaDeviceEnablePeerAccess(ranks[s]->device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                        fatal("cudaDeviceEnablePeerAccess failed: the GPUs of the node must be NVLink peers", __FILE__, __LINE__);
                    cudaGetLastError();
                }
        SB_CUDA(cudaSetDevice(cur));
    }
    void close()
    {
        if (P == 0) return;
        int cur = 0;
        cudaGetDevice(&cur);
        for (int g = 0; g < P; g++) {
            cudaSetDevice(ranks[g]->device);
            shards[g].release();
            ranks[g]->close();
            delete ranks[g];
        }
        cudaSetDevice(cur);
        ranks.clear(); shards.clear();
        P = 0;
    }
    void prepare_arenas(int n, int nb)
    {
        if (P == 1) return;
        bool grow = false;
        for (Rank *r : ranks) grow = grow || r->arena == nullptr || n > r->al.n_cap || nb > r->al.nb_cap;
        if (!grow) return;
        int cur = 0;
        SB_CUDA(cudaGetDevice(&cur));
        for (Rank *r : ranks) { r->al.n_cap = 0; r->al.nb_cap = 0; r->ensure_arena(n, nb); }    // all ranks re-allocate together
        for (Rank *r : ranks)
            for (int s = 0; s < P; s++) r->peer[s] = ranks[s]->arena;
        SB_CUDA(cudaSetDevice(cur));
    }
    template <typename F> void run(F body)
    {
        if (P == 1) { body(0); return; }
        int cur = 0;
        SB_CUDA(cudaGetDevice(&cur));
        std::vector<std::thread> th;
        for (int g = 0; g < P; g++) th.emplace_back([&body, g] { body(g); });
        for (auto &t : th) t.join();
        SB_CUDA(cudaSetDevice(cur));
    }
    void collect_stats(int n, int begin, int end, int nb)
    {
        stats = ranks[0]->stats;
        stats.n = n; stats.begin = begin; stats.end = end; stats.panel_width = nb; stats.ranks = P;
        for (int g = 1; g < P; g++) {
            stats.kernel_launches += ranks[g]->stats.kernel_launches;
            stats.gemm_flops += ranks[g]->stats.gemm_flops;
            stats.gemm_tma_launches += ranks[g]->stats.gemm_tma_launches;
            stats.gemm_cpasync_launches += ranks[g]->stats.gemm_cpasync_launches;
            stats.device_ms = std::max(stats.device_ms, ranks[g]->stats.device_ms);
        }
    }
    void reset_stats()
    {
        for (Rank *r : ranks) { memset(&r->stats, 0, sizeof(r->stats)); r->profile_level = profile_level; }
    }
};

static Team g_team;
static Rank *g_dist = nullptr;          // the rank of this process in one-process-per-GPU mode
static Shard g_dist_shard;

} // namespace sb200

// =================================================================================================
// node state (reference src/common/node.c) -- implemented in node.cpp
// =================================================================================================
extern "C" int starneig_b200_node_messages_enabled(void);
extern "C" int starneig_b200_node_pinning_enabled(void);

using namespace sb200;

// The Hessenberg path is CUDA-only: without a selected GPU the call fails loudly instead of falling
// back to a CPU implementation.
static bool have_gpu()
{
    if (starneig_node_get_gpus() >= 1) return true;
    fprintf(stderr, "[starneig][error] No CUDA device selected/available: the Hessenberg path has no CPU "
                    "implementation. Exiting...\n");
    return false;
}

extern "C" void starneig_b200_context_close(void)
{
    g_team.close();
    if (g_dist) { g_dist_shard.release(); g_dist->close(); delete g_dist; g_dist = nullptr; }
}

extern "C" __attribute__((visibility("default")))
void starneig_b200_get_stats(struct starneig_b200_stats *stats) { *stats = g_team.stats; }

extern "C" __attribute__((visibility("default")))
void starneig_b200_set_profile_level(int level) { g_team.profile_level = level; }

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
        local.panel_width = default_panel_width(n, std::min(std::max(starneig_node_get_gpus(), 1), MAX_RANKS));
        if (starneig_b200_node_messages_enabled())
            printf("[starneig][message] Setting panel width to %d.\n", local.panel_width);
    } else if (local.panel_width < 8) {
        fprintf(stderr, "[starneig][error] Invalid panel width. Exiting...\n");
        return STARNEIG_INVALID_CONFIGURATION;
    }
    *panel_width = local.panel_width;
    return STARNEIG_SUCCESS;
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
    if (n > SB_MAX_N) return STARNEIG_INVALID_ARGUMENTS;
    if (panel_width < 0) panel_width = default_panel_width(n);
    if (panel_width < 8) return STARNEIG_INVALID_CONFIGURATION;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    g_team.open(1);
    g_team.reset_stats();
    double t0 = wall_ms();
    g_team.ranks[0]->reduce(n, begin, end, panel_width, dA, ldA, dQ, ldQ, n);
    g_team.collect_stats(n, begin, end, panel_width);
    g_team.stats.wall_ms = wall_ms() - t0;
    return STARNEIG_SUCCESS;
}

namespace {
// page-locks a caller buffer for the duration of a call (no-op if it already is pinned or on failure)
struct ScopedPin {
    void *p = nullptr;
    ScopedPin(void *ptr, size_t bytes, bool enable)
    {
        if (!enable) return;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) return;
        cudaGetLastError();
        if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) == cudaSuccess) p = ptr;
        else cudaGetLastError();
    }
    ~ScopedPin() { if (p) cudaHostUnregister(p); }
};
}

namespace {
// all host threads of a team arrive before any of them goes on (one use per call)
struct ThreadRendezvous {
    std::mutex m; std::condition_variable cv; int waiting = 0; const int count;
    explicit ThreadRendezvous(int count_) : count(count_) {}
    void arrive()
    {
        std::unique_lock<std::mutex> lk(m);
        if (++waiting == count) cv.notify_all();
        else cv.wait(lk, [this] { return waiting >= count; });
    }
};
}

// the host-pointer call: upload, reduce, (writeback) download; `writeback == false` leaves H and Q in the library's device
// buffers of rank 0 (single GPU only)
static starneig_error_t hessenberg_host(struct starneig_hessenberg_conf *conf, int n, int begin, int end,
                                        double A[], int ldA, double Q[], int ldQ, bool writeback)
{
    int nb = 0;
    starneig_error_t ret = resolve_conf(conf, n, &nb);
    if (ret != STARNEIG_SUCCESS) return ret;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;

    const int P = std::min(starneig_node_get_gpus(), MAX_RANKS);
    if (!writeback && P != 1) {
        fprintf(stderr, "[starneig][error] The device-resident Hessenberg stage runs on one GPU (node initialised with %d). Exiting...\n", P);
        return STARNEIG_GENERIC_ERROR;
    }
    g_team.open(P);
    g_team.reset_stats();
    double t0 = wall_ms();
    {
        int cur = 0;
        SB_CUDA(cudaGetDevice(&cur));
        for (int g = 0; g < P; g++) {
            SB_CUDA(cudaSetDevice(g_team.ranks[g]->device));
            g_team.shards[g].ensure(*g_team.ranks[g], n);
            g_team.ranks[g]->ws.ensure(n, std::min(nb, PANEL_MAX_NB), P > 1);
            // (allocations happen here, before any rank's kernels wait for a peer: a cudaMalloc / cudaFree may wait for the
            // device, and ranks that share a device -- STARNEIG_B200_VIRTUAL_RANKS -- would then wait for each other)
            g_team.ranks[g]->prepare_history(n, begin, end, Q != nullptr);
        }
        SB_CUDA(cudaSetDevice(cur));
    }
    g_team.prepare_arenas(n, std::min(nb, PANEL_MAX_NB));
    std::vector<HostTimes> ht(P);
    {
        ScopedPin pinA(A, ((size_t)ldA * (n - 1) + n) * sizeof(double), starneig_b200_node_pinning_enabled());
        ScopedPin pinQ(Q, ((size_t)ldQ * (n - 1) + n) * sizeof(double), starneig_b200_node_pinning_enabled());
        ThreadRendezvous uploaded(P);
        const std::function<void()> rendezvous = [&uploaded] { uploaded.arrive(); };
        g_team.run([&](int g) {
            run_rank_host(*g_team.ranks[g], g_team.shards[g], n, begin, end, nb, A, ldA, Q, ldQ, &ht[g], writeback,
                          P > 1 ? &rendezvous : nullptr);
        });
    }
    g_team.collect_stats(n, begin, end, nb);
    g_team.stats.staging_overlapped = 1;
    for (int g = 0; g < P; g++) {
        g_team.stats.h2d_ms = std::max(g_team.stats.h2d_ms, ht[g].h2d_ms);
        g_team.stats.d2h_ms = std::max(g_team.stats.d2h_ms, ht[g].d2h_ms);
        g_team.stats.h2d_bytes += ht[g].h2d_bytes;
        g_team.stats.d2h_bytes += writeback ? ht[g].d2h_bytes : 0;
        g_team.stats.staging_overlapped = std::min(g_team.stats.staging_overlapped, ht[g].overlapped);
    }
    g_team.stats.wall_ms = wall_ms() - t0;
    return STARNEIG_SUCCESS;
}

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
    if (n > SB_MAX_N) {
        fprintf(stderr, "[starneig][error] Matrices larger than %d x %d are not supported by the CUDA Hessenberg path. Exiting...\n",
                SB_MAX_N, SB_MAX_N);
        return STARNEIG_INVALID_ARGUMENTS;
    }
    return hessenberg_host(conf, n, begin, end, A, ldA, Q, ldQ, true);
}

// ---------------------------------------------------------------------------------------------
// chain hand-off: the Hessenberg stage of starneig_SEP_SM_Reduce with H and Q left on the device
// ---------------------------------------------------------------------------------------------
static int g_stage_n = 0;       // order of the matrices the last stage call left on the device (0: none)

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_b200_SEP_SM_Hessenberg_stage(int n, double A[], int ldA, double Q[], int ldQ,
                                                       double **dH, int *lddH, double **dQ, int *lddQ)
{
    // same argument numbering as starneig_SEP_SM_Hessenberg (reference src/hessenberg/interface.c:175-184)
    if (n < 1) return -1;
    if (A == NULL) return -2;
    if (ldA < n) return -3;
    if (Q == NULL) return -4;
    if (ldQ < n) return -5;
    if (dH == NULL || lddH == NULL) return -6;
    if (dQ == NULL || lddQ == NULL) return -8;
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (n > SB_MAX_N) return STARNEIG_INVALID_ARGUMENTS;
    g_stage_n = 0;
    starneig_error_t ret = hessenberg_host(NULL, n, 0, n, A, ldA, Q, ldQ, false);
    if (ret != STARNEIG_SUCCESS) return ret;
    *dH = g_team.shards[0].A; *lddH = g_team.shards[0].ldA;
    *dQ = g_team.shards[0].Q; *lddQ = g_team.shards[0].ldQ;
    g_stage_n = n;
    return STARNEIG_SUCCESS;
}

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_b200_stage_fetch(int n, double A[], int ldA, double Q[], int ldQ)
{
    if (n < 1 || n != g_stage_n) return -1;
    if (A != NULL && ldA < n) return -3;
    if (Q != NULL && ldQ < n) return -5;
    if (!starneig_node_initialized() || g_team.P != 1) return STARNEIG_NOT_INITIALIZED;
    Rank &r = *g_team.ranks[0];
    Shard &sh = g_team.shards[0];
    SB_CUDA(cudaSetDevice(r.device));
    if (A) copy_columns(A, (size_t)ldA * 8, sh.A, (size_t)sh.ldA * 8, (size_t)n * 8, n, cudaMemcpyDeviceToHost, r.stream);
    if (Q) copy_columns(Q, (size_t)ldQ * 8, sh.Q, (size_t)sh.ldQ * 8, (size_t)n * 8, n, cudaMemcpyDeviceToHost, r.stream);
    SB_CUDA(cudaStreamSynchronize(r.stream));
    return STARNEIG_SUCCESS;
}

// The shape of starneig_SEP_SM_Reduce (reference src/common/combined.c:45-98) with this library's Hessenberg stage in
// front of caller-supplied next stages: the stages that take device pointers (`schur_device`) see H and Q where the
// Hessenberg stage left them, without a trip through host memory; host stages get them fetched first.
extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_b200_SEP_SM_Reduce(int n, double A[], int ldA, double Q[], int ldQ, double real[], double imag[],
                                             int (*predicate)(double real, double imag, void *arg), void *arg,
                                             int selected[], int *num_selected, const struct starneig_b200_chain *next)
{
    // argument checks: reference src/common/combined.c:56-63
    if (n < 1) return -1;
    if (A == NULL) return -2;
    if (ldA < n) return -3;
    if (Q == NULL) return -4;
    if (ldQ < n) return -5;
    if (next == NULL || (next->schur == NULL && next->schur_device == NULL)) return -12;
    if (predicate && (next->select == NULL || next->reorder_schur == NULL)) return -12;
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;

    starneig_error_t ret = STARNEIG_SUCCESS;
    int *own_selected = NULL;
    if (next->schur_device) {
        double *dH, *dQ;
        int lddH, lddQ;
        ret = starneig_b200_SEP_SM_Hessenberg_stage(n, A, ldA, Q, ldQ, &dH, &lddH, &dQ, &lddQ);
        if (ret != STARNEIG_SUCCESS) goto cleanup;
        ret = next->schur_device(n, dH, lddH, dQ, lddQ, real, imag);       // leaves S and Q on the device
        if (ret != STARNEIG_SUCCESS) goto cleanup;
        ret = starneig_b200_stage_fetch(n, A, ldA, Q, ldQ);
        if (ret != STARNEIG_SUCCESS) goto cleanup;
    } else {
        ret = starneig_SEP_SM_Hessenberg(n, A, ldA, Q, ldQ);
        if (ret != STARNEIG_SUCCESS) goto cleanup;
        ret = next->schur(n, A, ldA, Q, ldQ, real, imag);
        if (ret != STARNEIG_SUCCESS) goto cleanup;
    }
    if (predicate) {
        if (selected == NULL) selected = own_selected = (int *)malloc((size_t)n * sizeof(int));
        ret = next->select(n, A, ldA, predicate, arg, selected, num_selected);
        if (ret != STARNEIG_SUCCESS) goto cleanup;
        ret = next->reorder_schur(n, selected, A, ldA, Q, ldQ, real, imag);
    }
cleanup:
    free(own_selected);
    return ret;
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
// one process per GPU
// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default")))
int starneig_b200_dist_layout(int world, int rank, int n, int *col_block, int *local_cols, int *q_row0, int *q_rows)
{
    if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || n < 1) return STARNEIG_INVALID_ARGUMENTS;
    int cb = 64;
    const char *e = getenv("STARNEIG_B200_COL_BLOCK");
    if (e && atoi(e) >= 8) cb = atoi(e) / 8 * 8;
    if (world == 1) cb = n;
    const ColMap cm{world, rank, cb};
    int q0, q1;
    q_row_range(world, rank, n, &q0, &q1);
    if (col_block) *col_block = cb;
    if (local_cols) *local_cols = cm.lower(n);
    if (q_row0) *q_row0 = q0;
    if (q_rows) *q_rows = q1 - q0;
    return 0;
}

extern "C" __attribute__((visibility("default")))
int starneig_b200_dist_global_col(int world, int rank, int col_block, int local_col)
{
    const ColMap cm{world, rank, col_block};
    return cm.l2g(local_col);
}

extern "C" __attribute__((visibility("default")))
int starneig_b200_dist_init(int world, int rank, int n_max, int panel_width_max, void *handle_out)
{
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || n_max < 1 || n_max > SB_MAX_N) return STARNEIG_INVALID_ARGUMENTS;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    if (g_dist) { g_dist_shard.release(); g_dist->close(); delete g_dist; g_dist = nullptr; }
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    g_dist = new Rank();
    g_dist->open(world, rank, dev);
    // no limit given: room for the automatic width and for the reference's own default (src/hessenberg/interface.c:74-78)
    if (panel_width_max < 8) panel_width_max = std::max(default_panel_width(n_max, 1), (int)std::ceil((0.001875596476 * n_max + 273.5908216) / 8.0) * 8);
    panel_width_max = std::min(panel_width_max, PANEL_MAX_NB);
    memset(handle_out, 0, 64);
    if (world > 1) {
        g_dist->ensure_arena(n_max, panel_width_max);
        cudaIpcMemHandle_t h;
        SB_CUDA(cudaIpcGetMemHandle(&h, g_dist->arena));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handle_out, &h, 64);
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
int starneig_b200_dist_connect(const void *handles)
{
    if (!g_dist) return STARNEIG_NOT_INITIALIZED;
    Rank &r = *g_dist;
    for (int s = 0; s < r.P; s++) {
        if (s == r.g) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * s, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            fprintf(stderr, "[starneig][error] cudaIpcOpenMemHandle(rank %d) failed: %s\n", s, cudaGetErrorString(e));
            cudaGetLastError();
            return STARNEIG_GENERIC_ERROR;
        }
        r.peer[s] = (char *)p;
        r.arena_is_ipc[s] = true;
    }
    return 0;
}

static void dist_collect(Rank &r, int n, int begin, int end, int nb, double wall)
{
    g_team.stats = r.stats;
    g_team.stats.n = n; g_team.stats.begin = begin; g_team.stats.end = end; g_team.stats.panel_width = nb;
    g_team.stats.ranks = r.P;
    g_team.stats.wall_ms = wall;
}

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_b200_dist_hessenberg_device(int n, int begin, int end, int panel_width,
                                                      double *dA_loc, int ldA, double *dQ_loc, int ldQ)
{
    if (!g_dist) return STARNEIG_NOT_INITIALIZED;
    Rank &r = *g_dist;
    if (n < 1) return -1;
    if (begin < 0) return -2;
    if (n < end) return -3;
    if (dA_loc == NULL) return -5;
    if (ldA < n || (ldA & 1) || ((uintptr_t)dA_loc & 15)) return -6;
    if (dQ_loc == NULL) return -7;
    int q0, q1;
    q_row_range(r.P, r.g, n, &q0, &q1);
    if (ldQ < q1 - q0 || (ldQ & 1) || ((uintptr_t)dQ_loc & 15)) return -8;
    if (n > SB_MAX_N) return STARNEIG_INVALID_ARGUMENTS;
    if (panel_width < 0) panel_width = default_panel_width(n, r.P);
    if (panel_width < 8) return STARNEIG_INVALID_CONFIGURATION;
    if (r.P > 1 && (n > r.al.n_cap || std::min(panel_width, PANEL_MAX_NB) > r.al.nb_cap)) return STARNEIG_INVALID_ARGUMENTS;
    memset(&r.stats, 0, sizeof(r.stats));
    r.profile_level = g_team.profile_level;
    double t0 = wall_ms();
    r.reduce(n, begin, end, panel_width, dA_loc, ldA, dQ_loc, ldQ, q1 - q0);
    dist_collect(r, n, begin, end, panel_width, wall_ms() - t0);
    return STARNEIG_SUCCESS;
}

extern "C" __attribute__((visibility("default")))
starneig_error_t starneig_b200_dist_hessenberg_host(int n, int begin, int end, int panel_width,
                                                    double *A, int ldA, double *Q, int ldQ)
{
    if (!g_dist) return STARNEIG_NOT_INITIALIZED;
    Rank &r = *g_dist;
    if (n < 1) return -1;
    if (begin < 0) return -2;
    if (n < end) return -3;
    if (A == NULL) return -5;
    if (ldA < n) return -6;
    if (Q == NULL) return -7;
    if (ldQ < n) return -8;
    if (n > SB_MAX_N) return STARNEIG_INVALID_ARGUMENTS;
    if (panel_width < 0) panel_width = default_panel_width(n, r.P);
    if (panel_width < 8) return STARNEIG_INVALID_CONFIGURATION;
    if (r.P > 1 && (n > r.al.n_cap || std::min(panel_width, PANEL_MAX_NB) > r.al.nb_cap)) return STARNEIG_INVALID_ARGUMENTS;
    memset(&r.stats, 0, sizeof(r.stats));
    r.profile_level = g_team.profile_level;
    double t0 = wall_ms();
    g_dist_shard.ensure(r, n);
    HostTimes ht;
    run_rank_host(r, g_dist_shard, n, begin, end, panel_width, A, ldA, Q, ldQ, &ht);
    dist_collect(r, n, begin, end, panel_width, wall_ms() - t0);
    g_team.stats.h2d_ms = ht.h2d_ms; g_team.stats.d2h_ms = ht.d2h_ms;
    g_team.stats.h2d_bytes = ht.h2d_bytes; g_team.stats.d2h_bytes = ht.d2h_bytes;
    return STARNEIG_SUCCESS;
}

extern "C" __attribute__((visibility("default")))
void starneig_b200_dist_finalize(void)
{
    if (g_dist) { g_dist_shard.release(); g_dist->close(); delete g_dist; g_dist = nullptr; }
}

// host-only: the workspace plan of a reduction (see include/starneig_b200.h); lets the CPU test suite check the buffer
// bounds of both panel paths for every size up to STARNEIG_B200_MAX_N without a GPU
extern "C" __attribute__((visibility("default")))
int starneig_b200_plan_check(int n, int panel_width, int ranks, long long out[4])
{
    static_assert(STARNEIG_B200_MAX_N == SB_MAX_N, "header and engine disagree on the largest supported order");
    if (n < 1 || n > SB_MAX_N || ranks < 1 || ranks > MAX_RANKS) return STARNEIG_INVALID_ARGUMENTS;
    if (panel_width < 0) panel_width = default_panel_width(n, ranks);
    if (panel_width < 8) return STARNEIG_INVALID_CONFIGURATION;
    const int ctas = 148, slots = 148 * 8;      // a B200: one persistent CTA per SM; k_col_gemv: 8 resident blocks per SM
    int nb = std::min(panel_width, PANEL_MAX_NB);
    const int m0 = n - 1;
    if (m0 >= 1) nb = fit_panel_width(m0, nb, ctas, true);
    if (nb == 0) return STARNEIG_INVALID_ARGUMENTS;
    const int w = std::max(1, std::min(nb, m0));
    const int nsub = std::max(1, ceil_div(std::max(m0, 1), 32 * ctas));
    const size_t smem = nb <= FUSED_MAX_NB ? fused_smem_bytes(w, nsub, FUSED_KC) : 0;
    const size_t cap = ypart_doubles(n);
    // per-column path: the first panel has the longest columns; a rank's share of the trailing columns
    long long worst = 0;
    const int ldp = round_up(m0 + 2, 16);
    for (int j = 0; j < w; j++) {
        const int ncols = ceil_div(std::max(m0 - j, 0), ranks) + 64;
        for (int skip = 0; skip < 2; skip++) {
            const GemvPlan gp = plan_gemv_for(slots, skip, m0, std::min(ncols, n), ldp, cap);
            if (gp.S == 0) return STARNEIG_INVALID_ARGUMENTS;
            worst = std::max(worst, (long long)gp.S * ldp);
        }
    }
    // persistent kernel: slices of a group per row block (see ypart_doubles)
    if (smem > 0 && smem <= PANEL_SMEM_MAX) {
        const int RB = ceil_div(m0 + 1, 256);
        const long long groups = (long long)ctas * FUSED_VB;
        worst = std::max(worst, (groups / std::max(RB, 1) + 2) * (long long)ldp);
    }
    out[0] = nb; out[1] = (smem > 0 && smem <= PANEL_SMEM_MAX) ? (long long)smem : 0; out[2] = (long long)cap; out[3] = worst;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// unit-level entry points (single GPU)
// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default")))
int starneig_b200_dgemm(char transa, char transb, int m, int n, int k, double alpha, const double *dA, int lda,
                        const double *dB, int ldb, double beta, double *dC, int ldc)
{
    if (!starneig_node_initialized()) return STARNEIG_NOT_INITIALIZED;
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    g_team.open(1);
    Rank &r = *g_team.ranks[0];
    Rank::GemmKind kind;
    if (transa == 'N' && transb == 'T') kind = Rank::GEMM_NT;
    else if (transa == 'T' && transb == 'N') kind = Rank::GEMM_TN;
    else if (transa == 'N' && transb == 'N') kind = Rank::GEMM_NN;
    else return STARNEIG_INVALID_ARGUMENTS;
    if (m < 1 || n < 1 || k < 1) return STARNEIG_INVALID_ARGUMENTS;
    if (kind != Rank::GEMM_NT) r.ws.ensure(std::max(m, 16), std::max(n, 8), false);
    r.gemm(kind, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc);
    SB_CUDA(cudaStreamSynchronize(r.stream));
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
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    g_team.open(1);
    Rank &r = *g_team.ranks[0];
    Workspace &ws = r.ws;
    const int big = std::max(std::max(m, k), 16);
    ws.ensure(big, 8, false);
    cudaStream_t st = r.stream;
    double *scratch_col = nullptr;
    SB_CUDA(cudaMalloc(&scratch_col, (size_t)(big + 16) * sizeof(double)));
    PanelArgs pa = r.make_panel_args(m, ws.V, ws.Y, ws.VT, ws.ldv);
    SB_LAUNCH(k_set_gemv_inputs, ceil_div(k, 256), 256, 0, st, ws.scal, ws.pcol, dv, k);
    GemvPlan gp = r.plan_gemv(dA, m, k, pa.ldp);
    const ColMap cm{1, 0, std::max(k, 1)};
    Xchg x = r.make_xchg();
    size_t sh = (size_t)gp.kc * sizeof(double);
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    if (reps < 1) reps = 1;
    for (int it = 0; it < reps + 1; it++) {
        if (it == 1) SB_CUDA(cudaEventRecord(e0, st));
        SB_LAUNCH((k_col_gemv<false>), gp.RB * gp.S, GEMV_THREADS, sh, st, pa, 0, k, cm, 0, k, 0, gp.A0, lda, gp.skip, gp.kc, gp.RB, gp.S,
                                                                   scratch_col, x);
    }
    SB_CUDA(cudaEventRecord(e1, st));
    SB_LAUNCH(k_sum_partials, ceil_div(m, 256), 256, 0, st, m, gp.S, ws.ypart, pa.ldp, dy);
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
    if (!have_gpu()) return STARNEIG_GENERIC_ERROR;
    g_team.open(1);
    Rank &r = *g_team.ranks[0];
    r.ws.ensure(n, std::max(w, 8), false);
    const ColMap cm{1, 0, n};
    r.panel_factor(cm, i, end, w, dA, ldA, dA + (size_t)i * ldA + i + 1, ldA, dV, dY, dVT, ldw);
    SB_CUDA(cudaStreamSynchronize(r.stream));
    SB_CUDA(cudaGetLastError());
    if (htau) {
        std::vector<ColScal> sc(w);
        SB_CUDA(cudaMemcpy(sc.data(), r.ws.scal, w * sizeof(ColScal), cudaMemcpyDeviceToHost));
        for (int j = 0; j < w; j++) htau[j] = sc[j].tau;
    }
    return 0;
}
