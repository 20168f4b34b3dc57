/* ref_glue.c -- glue around the reference's own Hessenberg sources (oracle/_ref). TEST INFRASTRUCTURE ONLY.
 *
 * The reference's node lifecycle (src/common/node.c) starts StarPU, hwloc and cuBLAS; none of which exist
 * here. The Hessenberg interface only needs the five internal hooks below (src/common/node_internal.h,
 * called from src/hessenberg/interface.c:152-164), so they are provided as no-ops: the StarPU stand-in's
 * executor threads (if any) are started by oracle_ref_set_executors, and BLAS threading is left to the caller
 * (oracle_ref_set_threads).
 */
#include "ref_shim/starpu.h"
#include "ref_shim/cblas.h"

int starneig_node_initialized(void) { return 1; }
void starneig_node_set_mode(int mode) { (void)mode; }
void starneig_node_set_blas_mode(int mode) { (void)mode; }
void starneig_node_resume_starpu(void) {}
void starneig_node_pause_starpu(void) {}
void starneig_node_resume_awake_starpu(void) {}

void oracle_ref_set_threads(int threads) { openblas_set_num_threads(threads); }
void oracle_ref_set_workers(int workers) { oracle_starpu_set_worker_count((unsigned)workers); }
void oracle_ref_set_executors(int count) { oracle_starpu_set_executors(count); }
unsigned long oracle_ref_tasks_executed(int reset) { return oracle_starpu_tasks_executed(reset); }
