/* Stand-in for the StarPU task runtime -- implementation. See starpu.h in this directory.
 * Test infrastructure only (oracle/_ref); NOT StarPU and no StarPU code.
 *
 * Two schedules of the same sequential task flow:
 *   inline   (executor threads = 0, the default): starpu_task_insert() runs the codelet's CPU body at its insertion point on
 *            the calling thread. The golden fixtures and the parity tests use this one.
 *   parallel (oracle_starpu_set_executors(W), W >= 1): W worker threads execute the tasks as soon as their data dependencies
 *            allow, three priority levels, highest first, first-in first-out inside a level (one central queue: the shape of
 *            StarPU's "prio" policy, which the reference selects for CPU-only runs, src/common/node.c:326-331). Dependencies are inferred per data
 *            handle from the access modes in insertion order -- StarPU's sequential-consistency rule: a reader waits for the
 *            last writer, a writer for the last writer and for every reader since. STARPU_COMMUTE is honoured conservatively
 *            (ordered like a plain RW access: a valid schedule with less freedom than StarPU's; it also keeps the order of
 *            the floating-point sums, so the result equals the inline schedule's bit for bit when BLAS is sequential).
 *            STARPU_SCRATCH buffers are private to the executing task.
 *            bench.py --impl reference times the reference's task graph under this schedule with sequential BLAS inside the
 *            codelets, which is how the reference runs on CPU cores.
 */
#include "starpu.h"
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>

enum handle_kind { H_MATRIX, H_VECTOR, H_VARIABLE };

union iface {
    struct starpu_matrix_interface m;
    struct starpu_vector_interface v;
    struct starpu_variable_interface s;
};

struct task;

struct oracle_starpu_handle {
    enum handle_kind kind;
    int owns;                 /* runtime-allocated (home_node == -1) */
    size_t bytes;
    union iface u;
    /* parallel schedule (guarded by `lock`) */
    struct task *last_writer; /* most recent writer in insertion order, NULL once it has finished */
    struct task **readers;    /* unfinished readers since that writer */
    int nreaders, cap_readers;
    int pending;              /* submitted, unfinished tasks that name the handle */
    int zombie;               /* starpu_data_unregister_submit() was called: freed when `pending` drops to 0 */
};

struct task {
    struct starpu_codelet *cl;
    char *blob;
    int nbuf;
    starpu_data_handle_t *handles;
    int *modes;
    int level;                /* 0 = highest priority */
    int ndeps;                /* unfinished predecessors (+1 while the task is being submitted) */
    struct task **succ;
    int nsucc, cap_succ;
    struct task *next;        /* ready-queue link */
};

static unsigned worker_count = 1;
static unsigned long tasks_executed = 0;

/* ---- parallel executor state ---- */
#define MAX_EXECUTORS 256
static int executors = 0;                 /* 0: inline schedule */
static pthread_t threads[MAX_EXECUTORS];
static int threads_running = 0, shutting_down = 0;
static pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t cv_work = PTHREAD_COND_INITIALIZER;     /* a task became ready / shutdown */
static pthread_cond_t cv_main = PTHREAD_COND_INITIALIZER;     /* a task finished */
static struct task *q_head[3], *q_tail[3];
static int unfinished = 0;                /* submitted, unfinished tasks */

void oracle_starpu_set_worker_count(unsigned workers) { worker_count = workers ? workers : 1; }
unsigned long oracle_starpu_tasks_executed(int reset)
{
    pthread_mutex_lock(&lock);
    unsigned long t = tasks_executed;
    if (reset) tasks_executed = 0;
    pthread_mutex_unlock(&lock);
    return t;
}

static uintptr_t *ptr_of(enum handle_kind kind, union iface *u)
{
    return kind == H_MATRIX ? &u->m.ptr : kind == H_VECTOR ? &u->v.ptr : &u->s.ptr;
}

static void *poisoned(size_t bytes)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) { fprintf(stderr, "mini_starpu: out of memory\n"); abort(); }
    /* StarPU hands out uninitialised buffers; poison them so that any read-before-write shows up */
    memset(p, 0xff, bytes);
    return p;
}

static void ensure_allocated(starpu_data_handle_t h)
{
    if (!h->owns) return;
    uintptr_t *pp = ptr_of(h->kind, &h->u);
    if (*pp == 0) *pp = (uintptr_t)poisoned(h->bytes);
}

void starpu_matrix_data_register(starpu_data_handle_t *handle, int home_node, uintptr_t ptr,
    uint32_t ld, uint32_t nx, uint32_t ny, size_t elemsize)
{
    starpu_data_handle_t h = calloc(1, sizeof(*h));
    h->kind = H_MATRIX;
    h->owns = home_node < 0;
    h->u.m.ptr = h->owns ? 0 : ptr;
    h->u.m.ld = h->owns ? nx : ld;
    h->u.m.nx = nx; h->u.m.ny = ny; h->u.m.elemsize = elemsize;
    h->bytes = (size_t)h->u.m.ld * ny * elemsize;
    *handle = h;
}

void starpu_vector_data_register(starpu_data_handle_t *handle, int home_node, uintptr_t ptr,
    uint32_t nx, size_t elemsize)
{
    starpu_data_handle_t h = calloc(1, sizeof(*h));
    h->kind = H_VECTOR;
    h->owns = home_node < 0;
    h->u.v.ptr = h->owns ? 0 : ptr;
    h->u.v.nx = nx; h->u.v.elemsize = elemsize;
    h->bytes = (size_t)nx * elemsize;
    *handle = h;
}

void starpu_variable_data_register(starpu_data_handle_t *handle, int home_node, uintptr_t ptr, size_t size)
{
    starpu_data_handle_t h = calloc(1, sizeof(*h));
    h->kind = H_VARIABLE;
    h->owns = home_node < 0;
    h->u.s.ptr = h->owns ? 0 : ptr;
    h->u.s.elemsize = size;
    h->bytes = size;
    *handle = h;
}

static void destroy_handle(starpu_data_handle_t h)
{
    if (h->owns) free((void *)*ptr_of(h->kind, &h->u));
    free(h->readers);
    free(h);
}

/* the calling thread waits until no submitted task names the handle any more */
static void wait_for_handle(starpu_data_handle_t h)
{
    pthread_mutex_lock(&lock);
    while (h->pending > 0) pthread_cond_wait(&cv_main, &lock);
    pthread_mutex_unlock(&lock);
}

void starpu_data_unregister(starpu_data_handle_t h)
{
    if (h == NULL) return;
    wait_for_handle(h);
    destroy_handle(h);
}

void starpu_data_unregister_submit(starpu_data_handle_t h)
{
    if (h == NULL) return;
    pthread_mutex_lock(&lock);
    const int busy = h->pending > 0;
    if (busy) h->zombie = 1;
    pthread_mutex_unlock(&lock);
    if (!busy) destroy_handle(h);
}

void starpu_data_invalidate(starpu_data_handle_t h) { (void)h; }
int starpu_data_acquire(starpu_data_handle_t h, enum starpu_data_access_mode mode)
{
    (void)mode;
    wait_for_handle(h);
    ensure_allocated(h);
    return 0;
}
void starpu_data_release(starpu_data_handle_t h) { (void)h; }
int starpu_data_prefetch_on_node(starpu_data_handle_t h, unsigned node, unsigned async) { (void)h; (void)node; (void)async; return 0; }
void starpu_data_set_reduction_methods(starpu_data_handle_t h, struct starpu_codelet *a, struct starpu_codelet *b) { (void)h; (void)a; (void)b; }
uint32_t starpu_matrix_get_nx(starpu_data_handle_t h) { return h->u.m.nx; }
uint32_t starpu_matrix_get_ny(starpu_data_handle_t h) { return h->u.m.ny; }
size_t starpu_matrix_get_elemsize(starpu_data_handle_t h) { return h->u.m.elemsize; }

unsigned starpu_worker_get_count(void) { return worker_count; }
int starpu_worker_get_ids_by_type(enum starpu_worker_archtype type, int *ids, int maxsize) { (void)type; (void)ids; (void)maxsize; return 0; }
unsigned starpu_worker_get_memory_node(unsigned w) { (void)w; return 0; }
ssize_t starpu_memory_get_total(unsigned node) { (void)node; return 0; }

int starpu_task_wait_for_all(void)
{
    pthread_mutex_lock(&lock);
    while (unfinished > 0) pthread_cond_wait(&cv_main, &lock);
    pthread_mutex_unlock(&lock);
    return 0;
}

int starpu_task_wait_for_n_submitted(unsigned n)
{
    pthread_mutex_lock(&lock);
    while (unfinished > (int)n) pthread_cond_wait(&cv_main, &lock);
    pthread_mutex_unlock(&lock);
    return 0;
}

int starpu_task_nsubmitted(void)
{
    pthread_mutex_lock(&lock);
    const int n = unfinished;
    pthread_mutex_unlock(&lock);
    return n;
}

/* ---- running one task (either schedule) ---- */

/* scratch buffers of the executing thread (parallel schedule): the k-th STARPU_SCRATCH buffer of a task lives in slot k */
#define SCRATCH_SLOTS 8
static __thread struct { void *p; size_t cap; } scratch_slot[SCRATCH_SLOTS];

static void run_task(struct task *t)
{
    void *stack_bufs[64];
    union iface scratch[SCRATCH_SLOTS];
    void **buffers = t->nbuf <= 64 ? stack_bufs : malloc(sizeof(void *) * (size_t)t->nbuf);
    int nscratch = 0;
    for (int i = 0; i < t->nbuf; i++) {
        starpu_data_handle_t h = t->handles[i];
        if (executors > 0 && (t->modes[i] & STARPU_SCRATCH)) {
            /* a buffer of the handle's shape that belongs to this execution alone */
            if (nscratch == SCRATCH_SLOTS) { fprintf(stderr, "mini_starpu: too many scratch buffers in one task\n"); abort(); }
            if (scratch_slot[nscratch].cap < h->bytes) {
                free(scratch_slot[nscratch].p);
                scratch_slot[nscratch].p = poisoned(h->bytes);
                scratch_slot[nscratch].cap = h->bytes;
            }
            scratch[nscratch] = h->u;
            *ptr_of(h->kind, &scratch[nscratch]) = (uintptr_t)scratch_slot[nscratch].p;
            buffers[i] = &scratch[nscratch++];
        } else {
            buffers[i] = &h->u;
        }
    }
    t->cl->cpu_funcs[0](buffers, t->blob);
    if (buffers != stack_bufs) free(buffers);
}

static void free_task(struct task *t)
{
    free(t->blob); free(t->handles); free(t->modes); free(t->succ); free(t);
}

/* ---- parallel schedule ---- */

static void push_ready(struct task *t)            /* lock held */
{
    t->next = NULL;
    if (q_tail[t->level]) q_tail[t->level]->next = t; else q_head[t->level] = t;
    q_tail[t->level] = t;
    pthread_cond_signal(&cv_work);
}

static struct task *pop_ready(void)               /* lock held */
{
    for (int l = 0; l < 3; l++)
        if (q_head[l]) {
            struct task *t = q_head[l];
            q_head[l] = t->next;
            if (!q_head[l]) q_tail[l] = NULL;
            return t;
        }
    return NULL;
}

static void add_dependency(struct task *before, struct task *after)       /* lock held; `before` is unfinished */
{
    if (before == after) return;
    if (before->nsucc == before->cap_succ) {
        before->cap_succ = before->cap_succ ? 2 * before->cap_succ : 4;
        before->succ = realloc(before->succ, sizeof(struct task *) * (size_t)before->cap_succ);
    }
    before->succ[before->nsucc++] = after;
    after->ndeps++;
}

static void finish_task(struct task *t)           /* lock held */
{
    for (int i = 0; i < t->nsucc; i++)
        if (--t->succ[i]->ndeps == 0) push_ready(t->succ[i]);
    for (int i = 0; i < t->nbuf; i++) {
        starpu_data_handle_t h = t->handles[i];
        if (h == NULL) continue;
        for (int k = i + 1; k < t->nbuf; k++)             /* a handle named twice is settled once */
            if (t->handles[k] == h) t->handles[k] = NULL;
        if (h->last_writer == t) h->last_writer = NULL;
        for (int r = 0; r < h->nreaders; r++)
            if (h->readers[r] == t) { h->readers[r] = h->readers[--h->nreaders]; r--; }
        if (--h->pending == 0 && h->zombie) destroy_handle(h);
    }
    tasks_executed++;
    unfinished--;
    pthread_cond_broadcast(&cv_main);
}

static void *executor_main(void *arg)
{
    (void)arg;
    pthread_mutex_lock(&lock);
    for (;;) {
        struct task *t = pop_ready();
        if (t == NULL) {
            if (shutting_down) break;
            pthread_cond_wait(&cv_work, &lock);
            continue;
        }
        pthread_mutex_unlock(&lock);
        run_task(t);
        pthread_mutex_lock(&lock);
        finish_task(t);
        free_task(t);
    }
    pthread_mutex_unlock(&lock);
    for (int k = 0; k < SCRATCH_SLOTS; k++) { free(scratch_slot[k].p); scratch_slot[k].p = NULL; scratch_slot[k].cap = 0; }
    return NULL;
}

static void stop_executors(void)
{
    starpu_task_wait_for_all();
    pthread_mutex_lock(&lock);
    shutting_down = 1;
    pthread_cond_broadcast(&cv_work);
    pthread_mutex_unlock(&lock);
    for (int i = 0; i < threads_running; i++) pthread_join(threads[i], NULL);
    threads_running = 0;
    shutting_down = 0;
}

void oracle_starpu_set_executors(int count)
{
    if (count < 0) count = 0;
    if (count > MAX_EXECUTORS) count = MAX_EXECUTORS;
    if (threads_running > 0) stop_executors();
    executors = count;
    for (int i = 0; i < count; i++) {
        if (pthread_create(&threads[i], NULL, executor_main, NULL) != 0) { fprintf(stderr, "mini_starpu: pthread_create failed\n"); abort(); }
        threads_running++;
    }
}

static void submit(struct task *t)
{
    pthread_mutex_lock(&lock);
    t->ndeps = 1;                                  /* guard: not ready before every dependency is known */
    for (int i = 0; i < t->nbuf; i++) {
        starpu_data_handle_t h = t->handles[i];
        const int mode = t->modes[i];
        int seen = 0;
        for (int k = 0; k < i; k++) seen |= t->handles[k] == h;
        if (!seen) h->pending++;
        if (mode & STARPU_SCRATCH) continue;       /* private buffer: no ordering */
        if (h->last_writer) add_dependency(h->last_writer, t);
        if (mode & (STARPU_W | STARPU_REDUX)) {
            for (int r = 0; r < h->nreaders; r++) add_dependency(h->readers[r], t);
            h->nreaders = 0;
            h->last_writer = t;
        } else {
            if (h->nreaders == h->cap_readers) {
                h->cap_readers = h->cap_readers ? 2 * h->cap_readers : 4;
                h->readers = realloc(h->readers, sizeof(struct task *) * (size_t)h->cap_readers);
            }
            h->readers[h->nreaders++] = t;
        }
    }
    unfinished++;
    if (--t->ndeps == 0) push_ready(t);
    pthread_mutex_unlock(&lock);
}

/* argument blob layout: size_t count; then per argument { size_t size; bytes (8-byte padded) } */

int starpu_task_insert(struct starpu_codelet *cl, ...)
{
    struct task *t = calloc(1, sizeof(*t));
    int cap_buf = 16;
    t->cl = cl;
    t->handles = malloc(sizeof(starpu_data_handle_t) * (size_t)cap_buf);
    t->modes = malloc(sizeof(int) * (size_t)cap_buf);
    t->level = 1;
    size_t cap = 1024, used = sizeof(size_t);
    char *blob = malloc(cap);
    int nargs = 0;

#define ADD_BUFFER(h_, m_) do { \
        if (t->nbuf == cap_buf) { cap_buf *= 2; t->handles = realloc(t->handles, sizeof(starpu_data_handle_t) * (size_t)cap_buf); \
                                  t->modes = realloc(t->modes, sizeof(int) * (size_t)cap_buf); } \
        if (!(executors > 0 && ((m_) & STARPU_SCRATCH))) ensure_allocated(h_); t->handles[t->nbuf] = (h_); t->modes[t->nbuf++] = (int)(m_); } while (0)

    va_list ap;
    va_start(ap, cl);
    for (;;) {
        int tag = va_arg(ap, int);
        if (tag == 0) break;
        if (tag == STARPU_VALUE) {
            void *p = va_arg(ap, void *);
            size_t sz = va_arg(ap, size_t);
            size_t padded = (sz + 7) & ~(size_t)7;
            if (used + sizeof(size_t) + padded > cap) { cap = 2 * (used + sizeof(size_t) + padded); blob = realloc(blob, cap); }
            memcpy(blob + used, &sz, sizeof(size_t)); used += sizeof(size_t);
            memcpy(blob + used, p, sz); used += padded;
            nargs++;
        } else if (tag == STARPU_PRIORITY) {
            const int prio = va_arg(ap, int);
            t->level = prio >= STARPU_MAX_PRIO ? 0 : prio <= STARPU_MIN_PRIO ? 2 : 1;
        } else if (tag == STARPU_EXECUTE_ON_NODE) {
            (void)va_arg(ap, int);
        } else if (tag == STARPU_EXECUTE_ON_DATA) {
            (void)va_arg(ap, starpu_data_handle_t);
        } else if (tag == STARPU_FLOPS) {
            (void)va_arg(ap, double);
        } else if (tag == STARPU_DATA_MODE_ARRAY) {
            struct starpu_data_descr *d = va_arg(ap, struct starpu_data_descr *);
            int count = va_arg(ap, int);
            for (int i = 0; i < count; i++) ADD_BUFFER(d[i].handle, d[i].mode);
        } else if ((tag & ~(STARPU_RW | STARPU_SCRATCH | STARPU_REDUX | STARPU_COMMUTE)) == 0) {
            starpu_data_handle_t h = va_arg(ap, starpu_data_handle_t);
            ADD_BUFFER(h, tag);
        } else {
            fprintf(stderr, "mini_starpu: unknown task_insert tag %d (codelet %s)\n", tag, cl->name ? cl->name : "?");
            abort();
        }
    }
    va_end(ap);
#undef ADD_BUFFER
    size_t n = (size_t)nargs;
    memcpy(blob, &n, sizeof(size_t));
    t->blob = blob;

    if (cl->cpu_funcs[0] == NULL) { fprintf(stderr, "mini_starpu: codelet %s has no CPU body\n", cl->name ? cl->name : "?"); abort(); }
    if (executors > 0) {
        submit(t);
    } else {
        run_task(t);
        pthread_mutex_lock(&lock);
        tasks_executed++;
        pthread_mutex_unlock(&lock);
        free_task(t);
    }
    return 0;
}

void starpu_codelet_unpack_args(void *cl_arg, ...)
{
    char *blob = cl_arg;
    size_t n; memcpy(&n, blob, sizeof(size_t));
    size_t off = sizeof(size_t);
    va_list ap;
    va_start(ap, cl_arg);
    for (size_t i = 0; i < n; i++) {
        size_t sz; memcpy(&sz, blob + off, sizeof(size_t)); off += sizeof(size_t);
        void *dst = va_arg(ap, void *);
        if (dst == NULL) break;       /* StarPU allows a NULL-terminated shorter list */
        memcpy(dst, blob + off, sz);
        off += (sz + 7) & ~(size_t)7;
    }
    va_end(ap);
}
