/* Sequential stand-in for StarPU -- implementation. See starpu.h in this directory.
 * Test infrastructure only (oracle/_ref). */
#include "starpu.h"
#include <stdarg.h>
#include <stdio.h>

enum handle_kind { H_MATRIX, H_VECTOR, H_VARIABLE };

struct oracle_starpu_handle {
    enum handle_kind kind;
    int owns;                 /* runtime-allocated (home_node == -1) */
    size_t bytes;
    union {
        struct starpu_matrix_interface m;
        struct starpu_vector_interface v;
        struct starpu_variable_interface s;
    } u;
};

static unsigned worker_count = 1;
static unsigned long tasks_executed = 0;

void oracle_starpu_set_worker_count(unsigned workers) { worker_count = workers ? workers : 1; }
unsigned long oracle_starpu_tasks_executed(int reset)
{
    unsigned long t = tasks_executed;
    if (reset) tasks_executed = 0;
    return t;
}

static void ensure_allocated(starpu_data_handle_t h)
{
    if (!h->owns) return;
    uintptr_t *pp = h->kind == H_MATRIX ? &h->u.m.ptr : h->kind == H_VECTOR ? &h->u.v.ptr : &h->u.s.ptr;
    if (*pp == 0) {
        void *p = NULL;
        if (posix_memalign(&p, 64, h->bytes ? h->bytes : 64) != 0) { fprintf(stderr, "seq_starpu: out of memory\n"); abort(); }
        /* StarPU hands out uninitialised buffers; poison them so that any read-before-write shows up */
        memset(p, 0xff, h->bytes);
        *pp = (uintptr_t)p;
    }
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

void starpu_data_unregister(starpu_data_handle_t h)
{
    if (h == NULL) return;
    if (h->owns) {
        uintptr_t p = h->kind == H_MATRIX ? h->u.m.ptr : h->kind == H_VECTOR ? h->u.v.ptr : h->u.s.ptr;
        free((void *)p);
    }
    free(h);
}

void starpu_data_unregister_submit(starpu_data_handle_t h) { starpu_data_unregister(h); }
void starpu_data_invalidate(starpu_data_handle_t h) { (void)h; }
int starpu_data_acquire(starpu_data_handle_t h, enum starpu_data_access_mode mode) { (void)mode; ensure_allocated(h); return 0; }
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
int starpu_task_wait_for_all(void) { return 0; }
int starpu_task_wait_for_n_submitted(unsigned n) { (void)n; return 0; }
int starpu_task_nsubmitted(void) { return 0; }

/* argument blob layout: int count; then per argument { size_t size; bytes (8-byte padded) } */
#define MAX_TASK_BUFFERS 4096

int starpu_task_insert(struct starpu_codelet *cl, ...)
{
    static void *buffers[MAX_TASK_BUFFERS];
    int nbuf = 0;
    size_t cap = 1024, used = sizeof(size_t);
    char *blob = malloc(cap);
    int nargs = 0;

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
        } else if (tag == STARPU_PRIORITY || tag == STARPU_EXECUTE_ON_NODE) {
            (void)va_arg(ap, int);
        } else if (tag == STARPU_EXECUTE_ON_DATA) {
            (void)va_arg(ap, starpu_data_handle_t);
        } else if (tag == STARPU_FLOPS) {
            (void)va_arg(ap, double);
        } else if (tag == STARPU_DATA_MODE_ARRAY) {
            struct starpu_data_descr *d = va_arg(ap, struct starpu_data_descr *);
            int count = va_arg(ap, int);
            for (int i = 0; i < count; i++) {
                if (nbuf >= MAX_TASK_BUFFERS) { fprintf(stderr, "seq_starpu: too many buffers\n"); abort(); }
                ensure_allocated(d[i].handle);
                buffers[nbuf++] = &d[i].handle->u;
            }
        } else if ((tag & ~(STARPU_RW | STARPU_SCRATCH | STARPU_REDUX | STARPU_COMMUTE)) == 0) {
            starpu_data_handle_t h = va_arg(ap, starpu_data_handle_t);
            if (nbuf >= MAX_TASK_BUFFERS) { fprintf(stderr, "seq_starpu: too many buffers\n"); abort(); }
            ensure_allocated(h);
            buffers[nbuf++] = &h->u;
        } else {
            fprintf(stderr, "seq_starpu: unknown task_insert tag %d (codelet %s)\n", tag, cl->name ? cl->name : "?");
            abort();
        }
    }
    va_end(ap);
    size_t n = (size_t)nargs;
    memcpy(blob, &n, sizeof(size_t));

    if (cl->cpu_funcs[0] == NULL) { fprintf(stderr, "seq_starpu: codelet %s has no CPU body\n", cl->name ? cl->name : "?"); abort(); }
    cl->cpu_funcs[0](buffers, blob);
    tasks_executed++;
    free(blob);
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
