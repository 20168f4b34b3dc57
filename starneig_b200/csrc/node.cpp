// node.cpp -- node lifecycle (global singleton), the B200 counterpart of reference src/common/node.c.
//
// The reference configures StarPU workers, hwloc bindings and BLAS threading here (node.c:199-396,
// 434-584). None of that exists on this path: the "node" is the set of CUDA devices this process
// drives. Device objects (streams, events, workspace arena) are created by the Hessenberg context on
// first use and destroyed by starneig_node_finalize().
#include <starneig/starneig.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <thread>

extern "C" void starneig_b200_context_close(void);

namespace {
struct NodeState {
    bool is_init = false;
    starneig_flag_t flags = STARNEIG_DEFAULT;
    int avail_cores = 0, avail_gpus = 0;
    int used_cores = 0, used_gpus = 0;
    bool pinning = true;            // reference default: src/common/common.c:51
} state;

[[noreturn]] void fatal(const char *msg)
{
    // reference: starneig_fatal_error, src/common/common.c:143-152
    fprintf(stderr, "[starneig][fatal error] %s\n", msg);
    fflush(stderr);
    exit(EXIT_FAILURE);
}

void check_init()
{
    if (!state.is_init) fatal("The node is not initialized.");
}

int clip(int requested, int avail)
{
    if (requested == STARNEIG_USE_ALL || requested > avail) return avail;
    return requested < 0 ? 0 : requested;
}
}

extern "C" int starneig_b200_node_messages_enabled(void)
{
    return (state.flags & STARNEIG_NO_MESSAGES) != STARNEIG_NO_MESSAGES;
}

extern "C" int starneig_b200_node_pinning_enabled(void) { return state.pinning ? 1 : 0; }

extern "C" {

__attribute__((visibility("default")))
void starneig_node_init(int cores, int gpus, starneig_flag_t flags)
{
    if (state.is_init) fatal("The node is already initialized.");     // node.c:442-443
    state.flags = flags;

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); count = 0; }
    // development aid: STARNEIG_B200_VIRTUAL_RANKS=k lets up to k ranks share the available device(s), so that
    // the multi-GPU code path can be exercised on a single-GPU box (never set in production)
    if (count >= 1 && getenv("STARNEIG_B200_VIRTUAL_RANKS")) count = std::max(count, atoi(getenv("STARNEIG_B200_VIRTUAL_RANKS")));
    state.avail_gpus = count;
    unsigned hw = std::thread::hardware_concurrency();
    state.avail_cores = hw ? (int)hw : 1;
    state.used_cores = clip(cores, state.avail_cores);
    state.used_gpus = clip(gpus, state.avail_gpus);
    if (state.used_cores < 1) fatal("At least one CPU core must be selected.");   // node.c:252
    state.is_init = true;
    if ((flags & STARNEIG_NO_VERBOSE) == 0 && getenv("STARNEIG_B200_VERBOSE"))
        printf("[starneig][verbose] node: %d core(s) reported, %d of %d GPU(s) selected.\n",
               state.used_cores, state.used_gpus, state.avail_gpus);
}

__attribute__((visibility("default")))
int starneig_node_initialized(void) { return state.is_init ? 1 : 0; }

__attribute__((visibility("default")))
void starneig_node_finalize(void)
{
    check_init();
    starneig_b200_context_close();
    state.avail_cores = state.avail_gpus = 0;
    state.used_cores = state.used_gpus = 0;
    state.is_init = false;
}

__attribute__((visibility("default")))
int starneig_node_get_cores(void) { check_init(); return state.used_cores; }

__attribute__((visibility("default")))
void starneig_node_set_cores(int cores)
{
    check_init();
    state.used_cores = clip(cores, state.avail_cores);
    if (state.used_cores < 1) fatal("At least one CPU core must be selected.");
}

__attribute__((visibility("default")))
int starneig_node_get_gpus(void) { check_init(); return state.used_gpus; }

__attribute__((visibility("default")))
void starneig_node_set_gpus(int gpus) { check_init(); state.used_gpus = clip(gpus, state.avail_gpus); }

__attribute__((visibility("default")))
void starneig_node_enable_pinning(void) { state.pinning = true; }

__attribute__((visibility("default")))
void starneig_node_disable_pinning(void) { state.pinning = false; }

}
