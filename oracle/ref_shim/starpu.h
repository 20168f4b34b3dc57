/* Stand-in for the StarPU task runtime (NOT StarPU, no StarPU code).
 *
 * Purpose: let the reference's own Hessenberg path -- src/hessenberg/{interface,core,tasks,cpu}.c
 * and the src/common plumbing it uses -- compile from where it lies under /root/reference and run
 * on this image, which has no StarPU. StarPU's sequential-task-flow contract guarantees that
 * executing every task at its insertion point is a valid schedule, so by default starpu_task_insert()
 * here simply runs the codelet's CPU body immediately on the calling thread; with executor threads
 * (oracle_starpu_set_executors) the tasks run in parallel under the dependencies their access modes
 * imply (mini_starpu.c). Data handles are plain host buffers. Only the API surface the hot path
 * touches is declared.
 *
 * Test infrastructure only (oracle/_ref); never linked into the product library.
 */
#ifndef ORACLE_STARPU_SHIM_H
#define ORACLE_STARPU_SHIM_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#define STARPU_MAJOR_VERSION 1
#define STARPU_MINOR_VERSION 3

#define STARPU_NMAXWORKERS 64
#define STARPU_NMAXBUFS 8
#define STARPU_MAXIMPLEMENTATIONS 4
#define STARPU_MAIN_RAM 0
#define STARPU_MAX_PRIO 1
#define STARPU_DEFAULT_PRIO 0
#define STARPU_MIN_PRIO (-1)
#define STARPU_VARIABLE_NBUFFERS (-1)
#define STARPU_CUDA_ASYNC 1

typedef struct oracle_starpu_handle *starpu_data_handle_t;

enum starpu_data_access_mode {
    STARPU_NONE = 0, STARPU_R = 1, STARPU_W = 2, STARPU_RW = 3,
    STARPU_SCRATCH = 4, STARPU_REDUX = 8, STARPU_COMMUTE = 16
};

/* task_insert argument tags; kept clear of the access-mode bits */
#define STARPU_VALUE             (1 << 16)
#define STARPU_PRIORITY          (2 << 16)
#define STARPU_DATA_MODE_ARRAY   (3 << 16)
#define STARPU_EXECUTE_ON_NODE   (4 << 16)
#define STARPU_EXECUTE_ON_DATA   (5 << 16)
#define STARPU_FLOPS             (6 << 16)

struct starpu_data_descr {
    starpu_data_handle_t handle;
    enum starpu_data_access_mode mode;
};

struct starpu_matrix_interface {
    uintptr_t ptr;
    uint32_t nx;   /* rows (contiguous dimension) */
    uint32_t ny;   /* columns */
    uint32_t ld;
    size_t elemsize;
};

struct starpu_vector_interface {
    uintptr_t ptr;
    uint32_t nx;
    size_t elemsize;
};

struct starpu_variable_interface {
    uintptr_t ptr;
    size_t elemsize;
};

#define STARPU_MATRIX_GET_PTR(i)      (((struct starpu_matrix_interface *)(i))->ptr)
#define STARPU_MATRIX_GET_NX(i)       (((struct starpu_matrix_interface *)(i))->nx)
#define STARPU_MATRIX_GET_NY(i)       (((struct starpu_matrix_interface *)(i))->ny)
#define STARPU_MATRIX_GET_LD(i)       (((struct starpu_matrix_interface *)(i))->ld)
#define STARPU_MATRIX_GET_ELEMSIZE(i) (((struct starpu_matrix_interface *)(i))->elemsize)
#define STARPU_VECTOR_GET_PTR(i)      (((struct starpu_vector_interface *)(i))->ptr)
#define STARPU_VECTOR_GET_NX(i)       (((struct starpu_vector_interface *)(i))->nx)
#define STARPU_VECTOR_GET_ELEMSIZE(i) (((struct starpu_vector_interface *)(i))->elemsize)
#define STARPU_VARIABLE_GET_PTR(i)    (((struct starpu_variable_interface *)(i))->ptr)

struct starpu_codelet;

struct starpu_task {
    struct starpu_codelet *cl;
    void *cl_arg;
    size_t cl_arg_size;
};

enum starpu_perfmodel_type {
    STARPU_PERFMODEL_INVALID = 0, STARPU_PER_ARCH, STARPU_COMMON, STARPU_HISTORY_BASED,
    STARPU_REGRESSION_BASED, STARPU_NL_REGRESSION_BASED, STARPU_MULTIPLE_REGRESSION_BASED
};

struct starpu_perfmodel {
    enum starpu_perfmodel_type type;
    const char *symbol;
    size_t (*size_base)(struct starpu_task *, unsigned nimpl);
    void (*parameters)(struct starpu_task *task, double *parameters);
    const char **parameters_names;
    unsigned nparameters;
    unsigned **combinations;
    unsigned ncombinations;
};

typedef void (*starpu_cpu_func_t)(void **, void *);

struct starpu_codelet {
    const char *name;
    starpu_cpu_func_t cpu_funcs[STARPU_MAXIMPLEMENTATIONS];
    const char *cpu_funcs_name[STARPU_MAXIMPLEMENTATIONS];
    starpu_cpu_func_t cuda_funcs[STARPU_MAXIMPLEMENTATIONS];
    char cuda_flags[STARPU_MAXIMPLEMENTATIONS];
    int nbuffers;
    enum starpu_data_access_mode modes[STARPU_NMAXBUFS];
    struct starpu_perfmodel *model;
};

enum starpu_worker_archtype { STARPU_CPU_WORKER = 0, STARPU_CUDA_WORKER = 1 };

int starpu_task_insert(struct starpu_codelet *cl, ...);
void starpu_codelet_unpack_args(void *cl_arg, ...);

void starpu_matrix_data_register(starpu_data_handle_t *handle, int home_node, uintptr_t ptr,
    uint32_t ld, uint32_t nx, uint32_t ny, size_t elemsize);
void starpu_vector_data_register(starpu_data_handle_t *handle, int home_node, uintptr_t ptr,
    uint32_t nx, size_t elemsize);
void starpu_variable_data_register(starpu_data_handle_t *handle, int home_node, uintptr_t ptr, size_t size);
void starpu_data_unregister(starpu_data_handle_t handle);
void starpu_data_unregister_submit(starpu_data_handle_t handle);
void starpu_data_invalidate(starpu_data_handle_t handle);
int starpu_data_acquire(starpu_data_handle_t handle, enum starpu_data_access_mode mode);
void starpu_data_release(starpu_data_handle_t handle);
int starpu_data_prefetch_on_node(starpu_data_handle_t handle, unsigned node, unsigned async);
void starpu_data_set_reduction_methods(starpu_data_handle_t handle, struct starpu_codelet *redux_cl,
    struct starpu_codelet *init_cl);
uint32_t starpu_matrix_get_nx(starpu_data_handle_t handle);
uint32_t starpu_matrix_get_ny(starpu_data_handle_t handle);
size_t starpu_matrix_get_elemsize(starpu_data_handle_t handle);

unsigned starpu_worker_get_count(void);
int starpu_worker_get_ids_by_type(enum starpu_worker_archtype type, int *workerids, int maxsize);
unsigned starpu_worker_get_memory_node(unsigned workerid);
ssize_t starpu_memory_get_total(unsigned node);
int starpu_task_wait_for_all(void);
int starpu_task_wait_for_n_submitted(unsigned n);
int starpu_task_nsubmitted(void);

/* stand-in control: number of workers reported to the reference's default-tile-size formula */
void oracle_starpu_set_worker_count(unsigned workers);
/* stand-in control: number of threads that execute tasks in parallel (0: inline at the insertion point, the default) */
void oracle_starpu_set_executors(int count);
/* statistics: tasks executed since the last reset */
unsigned long oracle_starpu_tasks_executed(int reset);

#endif
