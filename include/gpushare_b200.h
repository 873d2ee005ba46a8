/*
 * gpushare_b200.h -- C ABI of the B200-native clause sharer (libgpushare_b200.so).
 *
 * This is the drop-in boundary for GpuShareSat's one hot path: checking every clause of the
 * GPU clause database against up to 1024 bit-packed partial assignments and reporting
 * (clause, solver, assignment-mask) hits.  Every entry point below replaces exactly one
 * virtual method of the reference's abstract class `GpuShare::GpuClauseSharer`
 * (/root/reference/gpuShareLib/GpuClauseSharer.h); the cited line is the declaration it
 * stands for.  A ~100-line C++ shim (gpusharesat_b200/shim/GpuClauseSharerShim.cc, see
 * INTEGRATION.md) implements the reference's class and factory on top of these calls, so
 * glucose-syrup and rel-newtech link the library unchanged.
 *
 * Conventions: plain C types only; no exceptions cross the boundary; `int` 0/1 where the
 * C++ API returns bool; contract violations print to stderr and exit(1) exactly like the
 * reference (gpuShareLib/Assert.h:44-48).  There is NO CPU fallback: gss_create() fails
 * loudly (exit 1) when no CUDA device is usable.
 *
 * Threading is the reference's (GpuClauseSharer.h:79,114,124): the "GPU thread" calls
 * gss_gpu_run / gss_reduce_db / stats; each solver thread calls the per-solver functions
 * for its own solver id; gss_add_clause may be called from any thread.
 */
#ifndef GPUSHARE_B200_H
#define GPUSHARE_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gss_sharer gss_sharer; /* opaque handle */

/* Field for field GpuShare::GpuClauseSharerOptions (GpuClauseSharer.h:25-59); -1 = default.
 * gpuBlockCountGuideline/gpuThreadsPerBlockGuideline only steer the grid size (a guideline in
 * the reference too); initReportCountPerCategory * block guideline sizes the first hit buffer
 * (hits are never dropped here: an overflowing run is re-run with a larger buffer). */
typedef struct gss_options {
    int    gpuBlockCountGuideline;
    int    gpuThreadsPerBlockGuideline;
    int    minGpuLatencyMicros;
    int    verbosity;
    double clauseActivityDecay;
    int    quickProf;                  /* bool */
    int    initReportCountPerCategory;
    int    maxPageLockedMemory;
} gss_options;

typedef void (*gss_log_fn)(const char *msg, void *ctx);

/* GpuClauseSharerOptions() constructor defaults (GpuClauseSharer.h:49-58) */
void gss_options_default(gss_options *opts);

/* makeGpuClauseSharerPtr (GpuClauseSharer.h:165, GpuClauseSharerImpl.cu:15).  The device is
 * the calling thread's current CUDA device unless GPUSHARE_DEVICE is set. */
gss_sharer *gss_create(const gss_options *opts, gss_log_fn log, void *log_ctx);
/* virtual ~GpuClauseSharer (GpuClauseSharer.h:161) */
void gss_destroy(gss_sharer *h);

/* ---- GPU-thread methods ---- */
void    gss_gpu_run(gss_sharer *h);                              /* gpuRun            :83  */
void    gss_reduce_db(gss_sharer *h);                            /* reduceDb          :87  */
int64_t gss_get_added_clause_count(gss_sharer *h);               /* getAddedClauseCount :90 */
int64_t gss_get_added_clause_count_at_last_reduce_db(gss_sharer *h); /* :92 */
int     gss_has_run_out_of_gpu_memory_once(gss_sharer *h);       /* :95 */
void    gss_get_gpu_mem_info(gss_sharer *h, size_t *free_bytes, size_t *total_bytes); /* :97 */
int     gss_get_global_stat_count(gss_sharer *h);                /* :99  */
int64_t gss_get_global_stat(gss_sharer *h, int stat);            /* :101 (GlobalStats.h order) */
const char *gss_get_global_stat_name(gss_sharer *h, int stat);   /* :120 */
void    gss_write_clauses_in_cnf(gss_sharer *h, FILE *file);     /* :103 */
void    gss_set_var_count(gss_sharer *h, int new_count);         /* :105 */
void    gss_set_cpu_solver_count(gss_sharer *h, int count);      /* :112 (not thread safe) */

/* ---- any thread ---- */
/* addClause :109 -- returns the clause id, or -1 when the clause is longer than the maximum
 * clause length (reference: MAX_CL_SIZE 100, Clauses.cu:350).  lits are borrowed. */
int64_t gss_add_clause(gss_sharer *h, int solver_id, const int *lits, int count);

/* ---- per-solver-thread methods ---- */
int     gss_try_set_solver_values(gss_sharer *h, int solver_id, const int *lits, int count); /* :129 */
void    gss_unset_solver_values(gss_sharer *h, int solver_id, const int *lits, int count);   /* :134 */
int64_t gss_try_send_assignment(gss_sharer *h, int solver_id);                                /* :139 */
/* popReportedClause :148 -- *lits points into library memory private to this solver, writable
 * (callers permute it in place), valid until the next pop for the same solver. */
int     gss_pop_reported_clause(gss_sharer *h, int solver_id, int **lits, int *count, int64_t *gpu_clause_id);
int64_t gss_get_last_assig_all_reported(gss_sharer *h, int solver_id);                        /* :151 */
void    gss_get_current_assignment(gss_sharer *h, int solver_id, uint8_t *assig);             /* :155 */
int     gss_get_one_solver_stat_count(gss_sharer *h);                                         /* :122 */
int64_t gss_get_one_solver_stat(gss_sharer *h, int solver_id, int stat);                      /* :159 */
const char *gss_get_one_solver_stat_name(gss_sharer *h, int stat);                            /* :118 */

/* ------------------------------------------------------------------------------------------
 * Additions that are NOT part of GpuClauseSharer.h: parity / bench hooks and the knobs the
 * reference hard-codes.  The C++ shim does not use them.
 * ---------------------------------------------------------------------------------------- */

/* one hit exactly as the kernels produce it (cf. ReportedClause, BaseTypes.cuh:146-151) */
typedef struct gss_hit {
    int64_t  clause_id; /* id returned by gss_add_clause */
    int32_t  solver_id;
    uint32_t mask;      /* bit p <=> fires on assignment slot p (= assignment id % 32) */
} gss_hit;

/* Hits of the most recently gathered run, sorted by (clause_id, solver_id).  Returns the
 * number of hits of that run; at most cap are written.  GPU thread only. */
int64_t gss_debug_last_hits(gss_sharer *h, gss_hit *out, int64_t cap);

/* Same as n calls of gss_add_clause(h, -1, ...) on CSR input, under one lock.  Returns the id
 * of the first clause added (ids are consecutive), or -1 if any clause is too long (none added). */
int64_t gss_add_clauses_bulk(gss_sharer *h, const int64_t *offsets, const int *lits, int64_t nclauses);

/* Maximum clause length accepted by gss_add_clause (default 100 = the reference's
 * MAX_CL_SIZE, BaseTypes.cuh:28; at most 65535).  Must be called before the first clause is
 * added.  The environment variable GPUSHARE_MAX_CLAUSE_LEN sets it at gss_create, so that the
 * limit can be lifted through the unmodified GpuClauseSharer.h (BASELINE config 5). */
void gss_set_max_clause_len(gss_sharer *h, int max_len);

/* kernel mode for the NEXT runs: 0 production (two-level filter + early exit, default),
 * 1 dense (bench only: no filter, no early exit, every (literal, 32-slot word) pair is
 * evaluated).  Hit sets are identical in both modes. */
void gss_debug_set_dense(gss_sharer *h, int dense);

/* Re-launch the check kernels `iters` times on the tables of the last started run (the
 * assignment tables stay intact until the next run starts) and return the average device
 * time of one sweep in microseconds (CUDA events on the library's stream).  The hit buffer
 * of the last iteration replaces the run's hits.  GPU thread only; waits for the run.
 * dense: 0 = production kernels (k_filter + k_exact), 1 = dense kernel, 2 = k_filter alone,
 * 3 = k_exact alone, 4 = k_apply_updates alone, 5 = k_collapse alone, 6 = k_emit alone (-1 when
 * the last run did not go through the direct pipeline). */
double gss_debug_time_check(gss_sharer *h, int iters, int dense);
/* The level-1 kernel is built in several variants (csrc/kernels.cu: kFilterVariants) so that the
 * choices can be timed against each other on the device; GSS_FILTER_VARIANT picks one at start-up. */
int gss_debug_filter_variants(void);
/* Cumulative host wall time in microseconds of the phases of gss_gpu_run: [0] finishing the previous
 * run (wait for the GPU, device-side sort / resolve of a large hit list, D2H), [1] starting the next
 * run (drain clauses, upload, collect the solvers' deltas, enqueue), [2] hand-over to the solver
 * queues, [3] collecting the deltas alone (part of 1), [4] waiting for the GPU alone (part of 0),
 * [5] sort / resolve / D2H of a large hit list alone (part of 0). */
void gss_debug_host_phases(gss_sharer *h, double out_us[6]);
const char *gss_debug_filter_variant_name(int v);
void gss_debug_set_filter_variant(int v);
/* same for the level-2 kernel (kExactVariants, GSS_EXACT_VARIANT) */
int gss_debug_exact_variants(void);
const char *gss_debug_exact_variant_name(int v);
void gss_debug_set_exact_variant(int v);

/* Device time in microseconds of the phases of the last gathered run (CUDA events on the
 * library's stream): [0] host->device copies (new clause tiles, run header, assignment deltas),
 * [1] table kernels (collapse of the previous batch + delta apply), [2] check kernels,
 * [3] total from the first copy to the end of the device->host copy of the hits.
 * Returns 0 if no run has been gathered. */
int gss_debug_last_run_times(gss_sharer *h, double out_us[4]);

/* Register-only LOP3 micro-benchmark on the sharer's device: returns the measured peak in
 * thread-level LOP3 operations per second (the integer-pipe roofline denominator). */
double gss_debug_lop3_peak(gss_sharer *h);

/* host->device bytes of the last STARTED run and device->host bytes of the last FINISHED run */
void gss_debug_last_run_bytes(gss_sharer *h, int64_t *h2d, int64_t *d2h);

/* number of kernel launches issued by the library so far */
int64_t gss_debug_kernel_launches(gss_sharer *h);

/* Sum of clause lengths / clause count currently in the database */
void gss_debug_db_size(gss_sharer *h, int64_t *nclauses, int64_t *nlits);

/* Clauses streamed in behind the first-literal-sorted part of their arena, and how often the arenas have
 * been put back in order on the device so far (csrc/reduce.cu) */
void gss_debug_db_order(gss_sharer *h, int64_t *unsorted_clauses, int64_t *resorts);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (new functionality; the reference drives device 0 only, GpuClauseSharerImpl.cu:52).
 * One process per GPU.  Every rank creates a sharer, calls gss_set_shard(rank, world) and then
 * the same gss_set_var_count / gss_set_cpu_solver_count / gss_add_clause sequence: the host
 * mirror and the device arenas are complete everywhere; rank r CHECKS its contiguous share of the
 * clause tiles of every length ([tiles*r/world, tiles*(r+1)/world)).
 * Rank 0 is the front-end (solver threads, assignment slots, hand-over).  Per batch:
 *   rank 0      gss_mgpu_collect   -> run parameters + assignment deltas (pinned host memory)
 *   all ranks   [broadcast them, e.g. NCCL over NVLink]  gss_mgpu_run(payload: host OR device ptr)
 *   all ranks   gss_mgpu_wait      -> this rank's hits (global clause references)
 *   rank 0      [gather the hits]  gss_mgpu_import(union of all ranks' hits)
 * ---------------------------------------------------------------------------------------- */
typedef struct gss_raw_hit {
    uint32_t mask;
    int32_t  solver;
    int32_t  len;  /* clause length */
    int32_t  idx;  /* index of the clause among the clauses of that length (global) */
} gss_raw_hit;

void    gss_set_shard(gss_sharer *h, int rank, int world);
/* returns 1 if this batch rebuilds the device tables (payload lists every variable), 0 if not,
 * -1 if there is no clause yet (nothing to run) */
int     gss_mgpu_collect(gss_sharer *h, const void **params, int64_t *params_bytes, const void **updates, int64_t *n_updates);
void    gss_mgpu_run(gss_sharer *h, const void *params, int64_t params_bytes, const void *updates, int64_t n_updates, int rebuild);
/* hits == NULL: only wait and return the count (the hits are fetched with gss_mgpu_hits_to_device) */
int64_t gss_mgpu_wait(gss_sharer *h, const gss_raw_hit **hits);
/* Fast path used by bench.py: the batch as ONE packed payload [64 B header][params][deltas].
 * gss_mgpu_collect_to (rank 0) collects and copies it host->device into dev_dst on the library's
 * stream and returns its size in bytes (the header says "nothing to run" when there is no clause);
 * after the broadcast every rank calls gss_mgpu_run_payload on its copy (returns -1 nothing ran,
 * 0 batch, 1 batch that rebuilt the tables).  gss_mgpu_hits_to_device copies this rank's hits of
 * the finished run device->device into dev_dst (the NCCL gather source) and returns their count. */
int64_t gss_mgpu_collect_to(gss_sharer *h, void *dev_dst, int64_t cap_bytes);
int     gss_mgpu_run_payload(gss_sharer *h, const void *dev_payload, int64_t payload_bytes);
int64_t gss_mgpu_hits_to_device(gss_sharer *h, void *dev_dst, int64_t cap_records);
/* Fully asynchronous receiver path (no host synchronisation between the broadcast and the gather):
 * gss_mgpu_enqueue_payload (ranks != 0) enqueues the whole batch from the first valid_bytes of the
 * broadcast payload without reading it on the host (-1: no clause yet); gss_mgpu_enqueue_result
 * (all ranks) enqueues [64 B header {int64 nHits, int64 overflow}][hits x cap_records] into dev_dst
 * (the all-gather source) and returns its size; after the caller has synchronised,
 * gss_mgpu_finish (all ranks) closes the run and returns 1 if this rank had to run again with
 * larger buffers (its header said overflow).  gss_mgpu_redo_payload (ranks != 0) runs the same batch
 * again on the complete payload when the broadcast had been truncated (header.totalBytes, the 4th
 * int64 of the payload, exceeded the predicted broadcast size). */
int     gss_mgpu_enqueue_payload(gss_sharer *h, const void *dev_payload, int64_t valid_bytes);
void    gss_mgpu_redo_payload(gss_sharer *h, const void *dev_payload, int64_t total_bytes);
int64_t gss_mgpu_enqueue_result(gss_sharer *h, void *dev_dst, int64_t cap_records);
int     gss_mgpu_finish(gss_sharer *h);
/* rank 0: hand over the union of all ranks' hits straight from the all-gathered DEVICE buffer
 * (world blocks of slot_bytes, each [64 B header][hits]; counts[r] = hits of rank r).  Large
 * unions are sorted and resolved on the device like the hits of a local run. */
void    gss_mgpu_import_gathered(gss_sharer *h, const void *dev_gathered, int world, int64_t slot_bytes, const int64_t *counts);
/* ---- multi-GPU, one process per GPU: no collective on the data path (csrc/peer.cu) ----
 * Every rank exports a window of its device memory with CUDA IPC.  Rank 0 pushes every batch into
 * the workers' windows with peer stores over NVLink (straight from the solver threads' page-locked
 * delta buffers) and signals a mailbox; the workers' streams wait on it with stream memory
 * operations.  Every rank sorts and emits its OWN hits into a ring of result buffers in a POSIX
 * shared-memory segment it owns (over its own PCIe link); rank 0 maps the rings and hands every
 * solver one slice per rank, merged into the single-device order.  Setup (once, after gss_set_shard):
 *   gss_peer_init     allocates this rank's window (and, on workers, its result ring: 4 buffers of
 *                     about slot_hits x 52 bytes in /dev/shm) and writes an opaque blob (<= 128
 *                     bytes) to blob_out; returns the blob size.  payload_cap: bytes reserved for one
 *                     batch (64 B + 288 B per solver + 12 B per variable delta); slot_hits: hit
 *                     records each rank may return per batch.  Same values on every rank.
 *   [exchange the blobs, e.g. torch.distributed.all_gather_object]
 *   gss_peer_connect  blobs of all ranks in rank order, blob_bytes apart.
 * Per batch, on every rank (SPMD, like the clause stream):
 *   gss_peer_enqueue  rank 0: collect the batch (buffer swap per solver), push it, check its own
 *                     tiles.  Workers: enqueue wait + kernels, never read the batch on the host.
 *                     Returns -1 no clause yet, 1 tables rebuilt, else 0.
 *   gss_peer_finish   waits for this rank's result (a rank whose buffers overflowed runs again by
 *                     itself before it publishes); rank 0 then waits for every rank's publication
 *                     and hands the results over.  Returns the number of hit records (rank 0: of
 *                     all ranks). */
int64_t gss_peer_init(gss_sharer *h, int rank, int world, int64_t payload_cap, int64_t slot_hits, void *blob_out, int64_t blob_cap);
void    gss_peer_connect(gss_sharer *h, const void *blobs, int64_t blob_bytes);
int     gss_peer_enqueue(gss_sharer *h);
int64_t gss_peer_finish(gss_sharer *h);
/* Workers also write their sorted record keys and masks (12 B per hit) next to their results: rank 0 then bumps the clause
 * activities of every rank's hits on its own device and gss_debug_last_hits covers the hits of every rank (the parity
 * hook of the tests and of bench.py).  Off (default; GPUSHARE_PEER_RECORDS=1 turns it on at start-up): every rank bumps
 * the activities of its own hits on its own device.  Same setting on every rank, between two batches. */
void    gss_debug_set_peer_records(gss_sharer *h, int on);
/* Make the library enqueue everything on the caller's CUDA stream (e.g. torch's current stream,
 * so that its work is ordered with the NCCL collectives without host synchronisation). */
void    gss_set_stream(gss_sharer *h, void *cuda_stream);
void    gss_mgpu_import(gss_sharer *h, const gss_raw_hit *hits, int64_t n);

/* library build info: "gpushare_b200 <version> sm_100a" */
const char *gss_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUSHARE_B200_H */
