/* <starneig/node.h> -- node lifecycle. Drop-in for reference src/include/starneig/node.h:72-241
 * (implementation there: src/common/node.c:434-650).
 *
 * The reference starts StarPU worker threads, hwloc and cuBLAS here. This library instead creates,
 * per selected GPU, the CUDA streams/events and the device workspace arena used by the Hessenberg
 * path; `cores` is recorded for reporting only (no CPU workers exist on this path).
 * Not re-entrant and not thread safe, like the reference (global singleton state, node.c:61-92). */
#ifndef STARNEIG_NODE_H
#define STARNEIG_NODE_H

#include <starneig/configuration.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned starneig_flag_t;                 /* node.h:72 */

#define STARNEIG_DEFAULT            0x0           /* node.h:78  */
#define STARNEIG_HINT_SM            0x0           /* node.h:86  */
#define STARNEIG_HINT_DM            0x1           /* node.h:95  */
#define STARNEIG_FXT_DISABLE        0x2           /* node.h:103 (accepted, no FxT here) */
#define STARNEIG_AWAKE_WORKERS      0x4           /* node.h:112 (accepted, streams are always live) */
#define STARNEIG_AWAKE_MPI_WORKER   0x8           /* node.h:122 (accepted, no MPI here) */
#define STARNEIG_FAST_DM            (STARNEIG_HINT_DM | STARNEIG_AWAKE_WORKERS | STARNEIG_AWAKE_MPI_WORKER)
#define STARNEIG_NO_VERBOSE         0x10          /* node.h:145 */
#define STARNEIG_NO_MESSAGES        (STARNEIG_NO_VERBOSE | 0x20)   /* node.h:152 */

#define STARNEIG_USE_ALL            -1            /* node.h:158 */

/* node.h:178 / node.c:434-584. gpus: number of GPUs to drive (STARNEIG_USE_ALL = every visible
 * device). Calling it twice without finalize is a fatal error (node.c:442-443). */
void starneig_node_init(int cores, int gpus, starneig_flag_t flags);

/* node.h:185 / node.c:586-590 */
int starneig_node_initialized(void);

/* node.h:192-215 / node.c:612-636 */
int starneig_node_get_cores(void);
void starneig_node_set_cores(int cores);
int starneig_node_get_gpus(void);
void starneig_node_set_gpus(int gpus);

/* node.h:220 / node.c:592-610 */
void starneig_node_finalize(void);

/* node.h:234,241 / src/common/common.c:53-67. When enabled (default), pageable caller buffers are
 * page-locked (cudaHostRegister) for the duration of a call so that host<->device copies run at
 * full PCIe rate and overlap with compute. */
void starneig_node_enable_pinning(void);
void starneig_node_disable_pinning(void);

#ifdef __cplusplus
}
#endif

#endif
