// cusim.cpp -- scheduler and runtime of the kernel-logic emulator (see include/cuda_runtime.h, include/cusim_device.h).
// TEST INFRASTRUCTURE, not part of the product.
//
// Execution model. A launch runs on the calling host thread: every CUDA thread of a block is a fiber (own stack, a
// hand-written context switch), fibers are resumed round-robin and run until they block (block barrier, named barrier,
// warp collective), poll (acquire / volatile loads yield) or return. A normal launch executes its blocks one after the
// other; a cooperative launch keeps the fibers of ALL blocks alive at once, so grid-wide barriers built from global
// counters work as on the hardware. Ranks of a multi-GPU run are separate host threads, each with its own scheduler:
// their kernels really run concurrently and talk through "peer" memory (plain host memory) with the same
// acquire/release and LL protocols as on NVLink.
// Checks the hardware does not make: a pass in which every live fiber is blocked on a barrier that cannot complete is
// reported as a deadlock (mismatched barrier counts); CUSIM_SHUFFLE=seed resumes the fibers in a random order and
// CUSIM_SKEW=N lets some blocks of a cooperative grid lag far behind the others, to shake out ordering assumptions;
// fresh "device" and shared memory is filled with NaN patterns.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64"
#endif

// void cusim_switch(void **save_sp, void *load_sp): saves the callee-saved registers of the SysV ABI on the current
// stack, stores the stack pointer, continues on the other stack
asm(".text\n"
    ".globl cusim_switch\n"
    ".type cusim_switch,@function\n"
    "cusim_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size cusim_switch,.-cusim_switch\n");
extern "C" void cusim_switch(void **save_sp, void *load_sp);

namespace cusim {

thread_local ThreadCtx *g_thread = nullptr;
unsigned long long g_keep_loads = 0, g_prefetches = 0;
namespace {
struct StatsAtExit {
    ~StatsAtExit()
    {
        if (getenv("CUSIM_STATS"))
            fprintf(stderr, "cusim: %llu loads with the keep-in-L2 policy, %llu L2 prefetches\n", g_keep_loads, g_prefetches);
    }
} stats_at_exit;
}

namespace {
constexpr size_t STACK_BYTES = 96 * 1024;

struct Warp {
    alignas(16) unsigned char slot[2][32][16];
    int arrived[2] = {0, 0};
    unsigned gen = 0;
    int live = 0;
};
struct NamedBar { int arrived = 0; unsigned gen = 0; };
struct Cta {
    int live = 0, bar_arrived = 0;
    unsigned bar_gen = 0;
    NamedBar named[16];
    std::vector<Warp> warps;
    std::vector<char> smem;
};
enum State { RUN, BLOCKED, DONE };
struct Fiber {
    void *sp = nullptr;
    ThreadCtx ctx;
    Cta *cta = nullptr;
    Warp *warp = nullptr;
    State st = RUN;
    const unsigned *wait_ptr = nullptr;
    unsigned wait_val = 0;
};
struct Sched {
    void *sp = nullptr;
    std::vector<Fiber> fibers;
    Fiber *cur = nullptr;
    const std::function<void()> *body = nullptr;
    std::vector<char *> stacks;
    bool progress = false, polled = false;
    int remaining = 0;
};
thread_local Sched tls;

void yield_to_scheduler()
{
    Sched &s = tls;
    Fiber *f = s.cur;
    cusim_switch(&f->sp, s.sp);
}

void block_on(const unsigned *ptr, unsigned val)
{
    Fiber *f = tls.cur;
    f->st = BLOCKED; f->wait_ptr = ptr; f->wait_val = val;
    yield_to_scheduler();
}

void fiber_exit()
{
    Sched &s = tls;
    Fiber *f = s.cur;
    f->st = DONE;
    s.remaining--;
    s.progress = true;
    Cta *c = f->cta;
    c->live--;
    if (c->live > 0 && c->bar_arrived == c->live) { c->bar_arrived = 0; c->bar_gen++; }
    Warp *w = f->warp;
    w->live--;
    const int ph = w->gen & 1;
    if (w->live > 0 && w->arrived[ph] == w->live) { w->arrived[ph] = 0; w->gen++; }
    void *dummy;
    cusim_switch(&dummy, s.sp);
    abort();
}

void fiber_entry()
{
    (*tls.body)();
    fiber_exit();
}

char *get_stack(Sched &s, size_t idx)
{
    while (s.stacks.size() <= idx) {
        void *p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("cusim: mmap"); abort(); }
        s.stacks.push_back((char *)p);
    }
    return s.stacks[idx];
}

void prepare_fiber(Sched &s, Fiber &f, size_t idx)
{
    char *top = get_stack(s, idx) + STACK_BYTES;      // 16-byte aligned (page aligned)
    void **sp = (void **)(top - 16);
    *sp = (void *)&fiber_entry;                         // return address popped by cusim_switch's ret
    sp -= 6;                                            // r15 r14 r13 r12 rbx rbp
    for (int k = 0; k < 6; k++) sp[k] = nullptr;
    f.sp = sp;
}

double wall_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}
} // namespace

unsigned long long now_ns()
{
    using namespace std::chrono;
    return (unsigned long long)duration_cast<nanoseconds>(steady_clock::now().time_since_epoch()).count();
}

void poll_yield()
{
    tls.polled = true;
    yield_to_scheduler();
}

void cta_barrier()
{
    Cta *c = tls.cur->cta;
    const unsigned my = c->bar_gen;
    if (++c->bar_arrived == c->live) { c->bar_arrived = 0; c->bar_gen = my + 1; tls.progress = true; }
    else block_on(&c->bar_gen, my);
}

void named_barrier(int id, int nthreads)
{
    if (id < 0 || id > 15) { fprintf(stderr, "cusim: named barrier id %d out of range\n", id); abort(); }
    NamedBar &b = tls.cur->cta->named[id];
    const unsigned my = b.gen;
    if (++b.arrived == nthreads) { b.arrived = 0; b.gen = my + 1; tls.progress = true; }
    else block_on(&b.gen, my);
}

void warp_exchange(const void *mine, size_t bytes, void *all)
{
    Fiber *f = tls.cur;
    Warp *w = f->warp;
    const unsigned my = w->gen;
    const int ph = my & 1;
    memcpy(w->slot[ph][f->ctx.lane], mine, bytes);
    if (++w->arrived[ph] == w->live) { w->arrived[ph] = 0; w->gen = my + 1; tls.progress = true; }
    else block_on(&w->gen, my);
    memcpy(all, w->slot[ph], sizeof(w->slot[ph]));
}

void run_grid(dim3 grid, dim3 block, size_t smem_bytes, bool cooperative, const std::function<void()> &body)
{
    Sched &s = tls;
    if (s.cur != nullptr) { fprintf(stderr, "cusim: nested launch\n"); abort(); }
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if (nthreads == 0 || nblocks == 0 || nthreads > 1024) { fprintf(stderr, "cusim: invalid launch configuration\n"); abort(); }
    const size_t nwarps = (nthreads + 31) / 32;
    const size_t batch = cooperative ? nblocks : 1;
    static const char *shuffle_env = getenv("CUSIM_SHUFFLE");
    static const char *timeout_env = getenv("CUSIM_TIMEOUT");
    static const char *skew_env = getenv("CUSIM_SKEW");
    const unsigned skew = skew_env ? (unsigned)atoi(skew_env) : 0u;
    const double timeout_s = timeout_env ? atof(timeout_env) : 120.0;
    std::mt19937 rng(shuffle_env ? (unsigned)atoi(shuffle_env) : 0u);
    s.body = &body;
    std::vector<Cta> ctas(batch);
    std::vector<size_t> order;
    for (size_t b0 = 0; b0 < nblocks; b0 += batch) {
        const size_t nb = std::min(batch, nblocks - b0);
        s.fibers.assign(nb * nthreads, Fiber());
        for (size_t bb = 0; bb < nb; bb++) {
            Cta &c = ctas[bb];
            c = Cta();
            c.live = (int)nthreads;
            c.warps.assign(nwarps, Warp());
            c.smem.assign(smem_bytes + 64, (char)0xff);
            const size_t bid = b0 + bb;
            for (size_t t = 0; t < nthreads; t++) {
                Fiber &f = s.fibers[bb * nthreads + t];
                f.cta = &c;
                f.warp = &c.warps[t / 32];
                f.warp->live++;
                f.ctx.t_idx = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
                f.ctx.b_idx = uint3{(unsigned)(bid % grid.x), (unsigned)((bid / grid.x) % grid.y), (unsigned)(bid / ((size_t)grid.x * grid.y))};
                f.ctx.b_dim = block; f.ctx.g_dim = grid;
                f.ctx.smem = (char *)(((uintptr_t)c.smem.data() + 15) & ~(uintptr_t)15);
                f.ctx.lane = (unsigned)(t & 31); f.ctx.warp = (unsigned)(t / 32);
                prepare_fiber(s, f, bb * nthreads + t);
            }
        }
        s.remaining = (int)s.fibers.size();
        order.resize(s.fibers.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = k;
        double last_progress = wall_s();
        while (s.remaining > 0) {
            s.progress = false; s.polled = false;
            bool resumed = false;
            if (shuffle_env) std::shuffle(order.begin(), order.end(), rng);
            for (size_t k : order) {
                Fiber &f = s.fibers[k];
                if (f.st == DONE) continue;
                // CUSIM_SKEW=N: block c of a cooperative grid only runs in one pass out of 1 + c % N, so some blocks
                // lag far behind the others (exposes missing ordering between a fast and a slow block)
                if (skew > 1 && batch > 1 && (rng() % (1u + (unsigned)(k / nthreads) % skew)) != 0u) { resumed = true; continue; }
                if (f.st == BLOCKED) {
                    if (*f.wait_ptr == f.wait_val) continue;
                    f.st = RUN;
                }
                resumed = true;
                s.cur = &f;
                g_thread = &f.ctx;
                cusim_switch(&s.sp, f.sp);
            }
            s.cur = nullptr; g_thread = nullptr;
            if (s.remaining == 0) break;
            if (!resumed) {
                fprintf(stderr, "cusim: DEADLOCK: %d live threads, all blocked on block/warp barriers that cannot complete\n", s.remaining);
                abort();
            }
            if (s.progress) last_progress = wall_s();
            else {
                std::this_thread::yield();      // only pollers ran: let the other ranks' host threads make progress
                if (wall_s() - last_progress > timeout_s) {
                    fprintf(stderr, "cusim: no progress for %.0f s (%d live threads polling): giving up\n", timeout_s, s.remaining);
                    abort();
                }
            }
        }
    }
    s.body = nullptr;
    s.fibers.clear();
}

} // namespace cusim

// ---------------------------------------------------------------------------------------------------------------------
// runtime API
// ---------------------------------------------------------------------------------------------------------------------
namespace {
thread_local int cur_device = 0;
int env_int(const char *name, int dflt) { const char *e = getenv(name); return e && atoi(e) > 0 ? atoi(e) : dflt; }
double now_ms() { return 1e-6 * (double)cusim::now_ns(); }
}

const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cusim error"; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = env_int("CUSIM_DEVICES", 4); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { cur_device = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = cur_device; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int)
{
    if (a == cudaDevAttrMultiProcessorCount) *v = env_int("CUSIM_SMS", 4);
    else if (a == cudaDevAttrCooperativeLaunch) *v = 1;
    else return cudaErrorInvalidValue;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
namespace {
std::mutex alloc_mutex;
std::map<uintptr_t, size_t> alloc_map;      // base -> bytes as requested by the caller
}
namespace cusim {
bool inside_device_allocation(const void *p, size_t bytes)
{
    std::lock_guard<std::mutex> lock(alloc_mutex);
    auto it = alloc_map.upper_bound((uintptr_t)p);
    if (it == alloc_map.begin()) return false;
    --it;
    return (uintptr_t)p + bytes <= it->first + it->second;
}
}
cudaError_t cusimMalloc(void **p, size_t bytes)
{
    // CUSIM_EXACT_ALLOC=1 (runs under AddressSanitizer): exactly `bytes`, so that the first byte past a device allocation is a
    // red zone; otherwise the allocation granularity of the hardware is mimicked
    static const bool exact = getenv("CUSIM_EXACT_ALLOC") && atoi(getenv("CUSIM_EXACT_ALLOC")) != 0;
    const size_t rounded = exact ? std::max<size_t>(bytes, 1) : (bytes + 255) / 256 * 256 + 256;
    void *q = nullptr;
    if (posix_memalign(&q, 256, rounded) != 0 || !q) return cudaErrorMemoryAllocation;
    memset(q, 0xff, rounded);       // NaN doubles, 0xffffffff flags: nothing may rely on fresh memory being zero
    *p = q;
    std::lock_guard<std::mutex> lock(alloc_mutex);
    alloc_map[(uintptr_t)q] = bytes;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (p) { std::lock_guard<std::mutex> lock(alloc_mutex); alloc_map.erase((uintptr_t)p); }
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemset2DAsync(void *p, size_t pitch, int v, size_t width, size_t height, cudaStream_t)
{
    for (size_t r = 0; r < height; r++) memset((char *)p + r * pitch, v, width);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t bytes, cudaMemcpyKind) { memmove(d, s, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(d, s, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < height; r++) memmove((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int prio) { *s = new cusimStream{prio}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new cusimStream{0}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t e, unsigned)
{
    // streams execute immediately, in program order: an event must have been recorded before anybody waits for it
    if (e->t_ms < 0.0) { fprintf(stderr, "cusim: cudaStreamWaitEvent on an event that was never recorded\n"); abort(); }
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cusimEvent{-1.0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new cusimEvent{-1.0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    if (a->t_ms < 0.0 || b->t_ms < 0.0) return cudaErrorInvalidValue;
    *ms = (float)(b->t_ms - a->t_ms);
    return cudaSuccess;
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *attr, const void *p)
{
    attr->type = cudaMemoryTypeHost; attr->device = 0; attr->devicePointer = (void *)p; attr->hostPointer = (void *)p;
    return cudaSuccess;
}
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
