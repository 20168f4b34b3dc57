/* starneig_b200.h -- extension entry points of the B200-native Hessenberg library (C ABI).
 *
 * The drop-in boundary is <starneig/starneig.h> (same symbols as the reference). The functions below
 * are additions that the reference does not have: a device-resident variant (no host round trip, for a
 * downstream GPU stage and for kernel-only timing), per-call statistics for the roofline report, and
 * unit-level access to the individual kernels so that tests can check each one against the CPU oracle
 * through the C ABI. All pointers named d* are DEVICE pointers on the current CUDA device.
 */
#ifndef STARNEIG_B200_H
#define STARNEIG_B200_H

#include <starneig/starneig.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Device-resident reduction: same semantics as starneig_SEP_SM_Hessenberg_expert
 * (reference src/hessenberg/interface.c:138-167) but dA/dQ already live in HBM. Requirements:
 * 16-byte aligned pointers and even leading dimensions (returns -6 / -8 otherwise).
 * panel_width < 0 selects the automatic width (192 on this hardware; the reference's own default,
 * interface.c:74-78, is fitted to its CPU codelets). Blocking. */
starneig_error_t starneig_b200_hessenberg_device(
    int n, int begin, int end, int panel_width, double *dA, int ldA, double *dQ, int ldQ);

/* Statistics of the most recent reduction on this process (times in ms, device times from CUDA events). */
struct starneig_b200_stats {
    int n, begin, end, panel_width, panels;
    double wall_ms;          /* host wall clock of the whole call (host API: includes copies) */
    double h2d_ms, d2h_ms;   /* host<->device staging (host API only) */
    double device_ms;        /* first kernel to last kernel */
    double panel_ms;         /* sum over panels: column loops (panel kernels + GEMV) */
    double trail_ms;         /* sum over panels: trailing right + left updates (critical path) */
    double other_ms;         /* sum over panels: top rows, partial columns and Q updates, plus the backward pass over Q */
    double gemv_ms;          /* sum of the durations of the event-timed GEMV launches (profile level >= 2) */
    long long gemv_launches;
    double gemv_bytes;       /* algorithmic bytes read by all GEMV launches: 8 * sum rows*cols */
    long long gemv_timed_launches;  /* launches bracketed by CUDA events: every 8th column at level 2, all at level 3 */
    double gemv_timed_bytes; /* algorithmic bytes of the timed launches */
    double finish_update_ms, reflector_ms;   /* same sampling: the two row-block kernels of the timed columns */
    long long kernel_launches;
    double gemm_flops;       /* flops executed by the DMMA kernels */
    long long h2d_bytes, d2h_bytes;
    int ranks;               /* GPUs that took part; device/gemv figures are rank 0's, launches and flops are summed */
    int fused_panels;        /* panels factorised by the persistent kernel (panel_fused.cuh); then gemv_ms is the
                              * device-side %globaltimer time of its GEMV phases and every column counts as timed */
    double fused_kernel_ms;  /* total run time of the persistent panel kernels (device-side timer) */
    double fused_phase_ms[4];/* its level-2 phases, each including the grid barrier that ends it: finish+update (A),
                              * w2 reduction (A'), reflector (R), scalars + s (R') */
    int overlap;             /* always 0: the deferred updates run in line (every overlapped variant lost on hardware) */
    double side_tail_ms;     /* end of the last trailing update -> end of the call (what the deferred updates still add) */
    long long gemm_tma_launches, gemm_cpasync_launches;   /* DMMA kernel launches by kind of tile movement (dgemm_tma.cuh / dgemm.cuh) */
    int staging_overlapped;  /* host API: 1 if the host buffers were page-locked from end to end, so that Q's upload and the
                              * write-back of finished columns really overlapped the reduction (then h2d_ms is the upload of A
                              * alone and d2h_ms what was left of the write-back when the reduction ended) */
    int panel_width_used;    /* panel width of the reduction (the requested one unless it exceeds what the panel kernels'
                              * shared-memory layout holds: > 1024 columns, or a narrower limit for n > ~70000) */
    int q_backward;          /* 1: Q was the identity on entry and was accumulated backward after the last panel (full reduction,
                              * one or several GPUs: 4/3 n^3 instead of 2 n^3 flops for Q; engine.cuh, Rank::reduce) */
    double q_backward_ms;    /* duration of that backward pass (also counted in other_ms) */
    int fused_slab_panels[2];/* panels of the persistent kernel without [0] / with [1] the CTA's rows of V resident in shared
                              * memory (panel_fused.cuh, FusedSmem) */
};
void starneig_b200_get_stats(struct starneig_b200_stats *stats);

/* ---- chain hand-off: the Hessenberg stage in front of GPU-resident next stages ----
 *
 * starneig_b200_SEP_SM_Hessenberg_stage: as starneig_SEP_SM_Hessenberg (reference src/hessenberg/interface.c:170-185; same
 * argument numbering for errors -1 .. -5), but H and Q are not copied back: they stay in the library's device buffers
 * (*dH, *dQ: column-major, leading dimensions *lddH, *lddQ; valid until the next Hessenberg call on this library or
 * starneig_node_finalize) for a next stage that runs on the GPU. Host A, Q are left untouched. One GPU.
 * starneig_b200_stage_fetch copies them (either may be NULL) to host arrays when a later stage runs on the host after all.
 *
 * starneig_b200_SEP_SM_Reduce has the shape of starneig_SEP_SM_Reduce (reference src/common/combined.c:45-98: Hessenberg,
 * Schur, optional Select + ReorderSchur; same argument numbering, -12 for an incomplete `next`) with the stages this
 * library does not own supplied by the caller -- in a StarNEig build: starneig_SEP_SM_Schur, starneig_SEP_SM_Select,
 * starneig_SEP_SM_ReorderSchur (reference src/include/starneig/sep_sm.h). A `schur_device` stage takes the DEVICE pointers
 * of H and Q (no host round trip between the stages) and leaves the Schur form and the updated Q there. */
struct starneig_b200_chain {
    starneig_error_t (*schur)(int n, double H[], int ldH, double Q[], int ldQ, double real[], double imag[]);
    starneig_error_t (*schur_device)(int n, double *dH, int lddH, double *dQ, int lddQ, double real[], double imag[]);
    starneig_error_t (*select)(int n, double S[], int ldS, int (*predicate)(double real, double imag, void *arg), void *arg,
                               int selected[], int *num_selected);
    starneig_error_t (*reorder_schur)(int n, int selected[], double S[], int ldS, double Q[], int ldQ, double real[], double imag[]);
};
starneig_error_t starneig_b200_SEP_SM_Hessenberg_stage(int n, double A[], int ldA, double Q[], int ldQ,
                                                       double **dH, int *lddH, double **dQ, int *lddQ);
starneig_error_t starneig_b200_stage_fetch(int n, double A[], int ldA, double Q[], int ldQ);
starneig_error_t starneig_b200_SEP_SM_Reduce(int n, double A[], int ldA, double Q[], int ldQ, double real[], double imag[],
                                             int (*predicate)(double real, double imag, void *arg), void *arg,
                                             int selected[], int *num_selected, const struct starneig_b200_chain *next);

/* Largest supported matrix order (two n x n FP64 matrices of this order exceed one B200's memory anyway); the
 * entry points return STARNEIG_INVALID_ARGUMENTS beyond it instead of overrunning a workspace. */
#define STARNEIG_B200_MAX_N 131056

/* Host-only arithmetic (no GPU needed): the workspace plan for an n x n reduction with the given panel width
 * (< 0: the automatic width) on `ranks` GPUs. out[0] = panel width that would be used, out[1] = dynamic shared memory
 * (bytes) of the persistent panel kernel for the first panel (0: it does not fit, the per-column kernels run),
 * out[2] = capacity of the GEMV partial-sum buffer (doubles), out[3] = largest number of doubles any column of the
 * per-column path would write into it. Returns 0, or STARNEIG_INVALID_ARGUMENTS if n is not supported. */
int starneig_b200_plan_check(int n, int panel_width, int ranks, long long out[4]);

/* 0: no extra events; 1: per-panel phase events (default); 2: additionally time every 8th GEMV launch;
 * 3: time every GEMV launch */
void starneig_b200_set_profile_level(int level);

/* ---- one process per GPU (torchrun): 1-D block-cyclic column shards of A, row slabs of Q ----
 *
 * Every rank process selects its CUDA device, calls starneig_node_init, then
 *   starneig_b200_dist_init(world, rank, n_max, panel_width_max, handle)   -> 64-byte cudaIpcMemHandle of the
 *       rank's exchange arena (panel_width_max < 8: the automatic width for n_max),
 *   all-gathers the handles (e.g. torch.distributed.all_gather) and passes them, in rank order, to
 *   starneig_b200_dist_connect(handles).
 * After that the ranks call the reduction collectively. All data-path communication (the per-column sum of
 * the GEMV partials, the panel gather, the per-panel sum of the top-row products) is done by the kernels over
 * NVLink peer memory; no NCCL call sits on the path.
 *
 * Layout (starneig_b200_dist_layout): global column c belongs to rank (c / col_block) % world; a rank stores
 * its columns contiguously in ascending global order, full height n (local_cols of them); Q is split by rows:
 * rank r holds rows [q_row0, q_row0 + q_rows) of all n columns. Host-only arithmetic, no GPU needed. */
int starneig_b200_dist_layout(int world, int rank, int n, int *col_block, int *local_cols, int *q_row0, int *q_rows);
int starneig_b200_dist_global_col(int world, int rank, int col_block, int local_col);
int starneig_b200_dist_init(int world, int rank, int n_max, int panel_width_max, void *handle_out);
int starneig_b200_dist_connect(const void *handles);
/* shards already in HBM: dA_loc (n x local_cols, ldA >= n), dQ_loc (q_rows x n, ldQ >= q_rows); 16-byte
 * aligned, even leading dimensions. Same error numbering as starneig_b200_hessenberg_device. Collective. */
starneig_error_t starneig_b200_dist_hessenberg_device(
    int n, int begin, int end, int panel_width, double *dA_loc, int ldA, double *dQ_loc, int ldQ);
/* host arrays holding the WHOLE matrices (e.g. in memory shared by the rank processes): every rank moves
 * only its own shards to its GPU and back. Collective. */
starneig_error_t starneig_b200_dist_hessenberg_host(
    int n, int begin, int end, int panel_width, double *A, int ldA, double *Q, int ldQ);
void starneig_b200_dist_finalize(void);

/* ---- unit-level kernel access (tests / bench) ---- */

/* C = alpha*op(A)*op(B) + beta*C with the DMMA kernels; transa/transb in {'N','T'}; supported
 * combinations: NT, TN, NN. Returns 0 or STARNEIG_INVALID_ARGUMENTS. Synchronous. */
int starneig_b200_dgemm(char transa, char transb, int m, int n, int k, double alpha,
    const double *dA, int lda, const double *dB, int ldb, double beta, double *dC, int ldc);

/* y = A(m x k) * v with the panel GEMV kernel (v given explicitly). dA may have any 8-byte alignment;
 * lda must be even. reps > 1 repeats the launch and returns the mean kernel time in ms (CUDA events). */
int starneig_b200_gemv(int m, int k, const double *dA, int lda, const double *dv, double *dy,
    int reps, float *mean_ms);

/* one panel factorisation only (columns i .. i+w-1 of the reduction of rows/cols < end), leaving
 * V, Y, VT (m x w, leading dimension ldw) in the given device buffers; for tests. */
int starneig_b200_panel(int n, int i, int end, int w, double *dA, int ldA,
    double *dV, double *dY, double *dVT, int ldw, double *htau);

#ifdef __cplusplus
}
#endif

#endif
