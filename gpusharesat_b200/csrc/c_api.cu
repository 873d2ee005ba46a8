// c_api.cu -- the extern "C" boundary declared in include/gpushare_b200.h.  Thin forwarding
// only; every function names the GpuClauseSharer.h method it stands for in the header.
#include "../../include/gpushare_b200.h"
#include "sharer.h"

using gss::Sharer;

struct gss_sharer {
    Sharer impl;
    gss_sharer(const gss_options &o, gss_log_fn log, void *ctx) : impl(o, log, ctx) {}
};

extern "C" {

void gss_options_default(gss_options *o) {
    // GpuClauseSharerOptions(), GpuClauseSharer.h:49-58
    o->gpuBlockCountGuideline = -1;
    o->gpuThreadsPerBlockGuideline = -1;
    o->minGpuLatencyMicros = -1;
    o->verbosity = 1;
    o->clauseActivityDecay = -1;
    o->quickProf = 1;
    o->initReportCountPerCategory = -1;
    o->maxPageLockedMemory = -1;
}

gss_sharer *gss_create(const gss_options *opts, gss_log_fn log, void *log_ctx) {
    gss_options o;
    if (opts) o = *opts; else gss_options_default(&o);
    return new gss_sharer(o, log, log_ctx);
}

void gss_destroy(gss_sharer *h) { delete h; }

void gss_gpu_run(gss_sharer *h) { h->impl.gpuRun(); }
void gss_reduce_db(gss_sharer *h) { h->impl.reduceDb(); }
int64_t gss_get_added_clause_count(gss_sharer *h) { return h->impl.addedClauseCount(); }
int64_t gss_get_added_clause_count_at_last_reduce_db(gss_sharer *h) { return h->impl.addedClauseCountAtLastReduceDb(); }
int gss_has_run_out_of_gpu_memory_once(gss_sharer *h) { return h->impl.hasRunOutOfGpuMemoryOnce() ? 1 : 0; }
void gss_get_gpu_mem_info(gss_sharer *h, size_t *free_bytes, size_t *total_bytes) { h->impl.gpuMemInfo(free_bytes, total_bytes); }
int gss_get_global_stat_count(gss_sharer *) { return gss::G_COUNT; }
int64_t gss_get_global_stat(gss_sharer *h, int stat) {
    if (stat < 0 || stat >= gss::G_COUNT) return 0;
    return h->impl.globalStat(stat);
}
const char *gss_get_global_stat_name(gss_sharer *, int stat) {
    return (stat < 0 || stat >= gss::G_COUNT) ? nullptr : gss::kGlobalStatNames[stat];
}
void gss_write_clauses_in_cnf(gss_sharer *h, FILE *file) { h->impl.writeClausesInCnf(file); }
void gss_set_var_count(gss_sharer *h, int n) { h->impl.setVarCount(n); }
void gss_set_cpu_solver_count(gss_sharer *h, int n) { h->impl.setCpuSolverCount(n); }

int64_t gss_add_clause(gss_sharer *h, int solver_id, const int *lits, int count) { return h->impl.addClause(solver_id, lits, count); }

int gss_try_set_solver_values(gss_sharer *h, int s, const int *lits, int count) { return h->impl.trySetSolverValues(s, lits, count) ? 1 : 0; }
void gss_unset_solver_values(gss_sharer *h, int s, const int *lits, int count) { h->impl.unsetSolverValues(s, lits, count); }
int64_t gss_try_send_assignment(gss_sharer *h, int s) { return h->impl.trySendAssignment(s); }
int gss_pop_reported_clause(gss_sharer *h, int s, int **lits, int *count, int64_t *id) {
    int *l = nullptr;
    int c = 0;
    int64_t i = 0;
    if (!h->impl.popReportedClause(s, l, c, i)) return 0;
    *lits = l;
    *count = c;
    *id = i;
    return 1;
}
int64_t gss_get_last_assig_all_reported(gss_sharer *h, int s) { return h->impl.lastAssigAllReported(s); }
void gss_get_current_assignment(gss_sharer *h, int s, uint8_t *assig) { h->impl.currentAssignment(s, assig); }
int gss_get_one_solver_stat_count(gss_sharer *) { return gss::S_COUNT; }
int64_t gss_get_one_solver_stat(gss_sharer *h, int s, int stat) {
    if (stat < 0 || stat >= gss::S_COUNT) return 0;
    return h->impl.oneSolverStat(s, stat);
}
const char *gss_get_one_solver_stat_name(gss_sharer *, int stat) {
    return (stat < 0 || stat >= gss::S_COUNT) ? nullptr : gss::kOneSolverStatNames[stat];
}

int64_t gss_debug_last_hits(gss_sharer *h, gss_hit *out, int64_t cap) { return h->impl.lastHits(out, cap); }
int64_t gss_add_clauses_bulk(gss_sharer *h, const int64_t *offsets, const int *lits, int64_t n) { return h->impl.addClausesBulk(offsets, lits, n); }
void gss_set_max_clause_len(gss_sharer *h, int max_len) { h->impl.setMaxClauseLen(max_len); }
void gss_debug_set_dense(gss_sharer *h, int dense) { h->impl.setDense(dense != 0); }
double gss_debug_time_check(gss_sharer *h, int iters, int dense) {
    // 0 production (k_filter + k_exact), 1 dense kernel, 2 k_filter, 3 k_exact, 4 k_apply_updates, 5 k_collapse
    return h->impl.timeCheck(iters, dense);
}
void gss_debug_host_phases(gss_sharer *h, double out_us[6]) { h->impl.hostPhases(out_us); }
int gss_debug_filter_variants(void) { return gss::filterVariantCount(); }
const char *gss_debug_filter_variant_name(int v) { return gss::filterVariantName(v); }
void gss_debug_set_filter_variant(int v) { gss::setFilterVariant(v); }
int gss_debug_exact_variants(void) { return gss::exactVariantCount(); }
const char *gss_debug_exact_variant_name(int v) { return gss::exactVariantName(v); }
void gss_debug_set_exact_variant(int v) { gss::setExactVariant(v); }
int gss_debug_last_run_times(gss_sharer *h, double out_us[4]) { return h->impl.lastRunTimes(out_us); }
double gss_debug_lop3_peak(gss_sharer *h) { return h->impl.lop3Peak(); }
void gss_debug_last_run_bytes(gss_sharer *h, int64_t *h2d, int64_t *d2h) { h->impl.lastRunBytes(h2d, d2h); }
int64_t gss_debug_kernel_launches(gss_sharer *h) { return h->impl.kernelLaunches(); }
void gss_debug_db_size(gss_sharer *h, int64_t *nclauses, int64_t *nlits) { h->impl.dbSize(nclauses, nlits); }
void gss_debug_db_order(gss_sharer *h, int64_t *unsorted_clauses, int64_t *resorts) { h->impl.dbOrder(unsorted_clauses, resorts); }
void gss_set_shard(gss_sharer *h, int rank, int world) { h->impl.setShard(rank, world); }
int gss_mgpu_collect(gss_sharer *h, const void **params, int64_t *params_bytes, const void **updates, int64_t *n_updates) {
    return h->impl.mgpuCollect(params, params_bytes, updates, n_updates);
}
void gss_mgpu_run(gss_sharer *h, const void *params, int64_t params_bytes, const void *updates, int64_t n_updates, int rebuild) {
    h->impl.mgpuRun(params, params_bytes, updates, n_updates, rebuild);
}
int64_t gss_mgpu_collect_to(gss_sharer *h, void *dev_dst, int64_t cap_bytes) { return h->impl.mgpuCollectTo(dev_dst, cap_bytes); }
int gss_mgpu_run_payload(gss_sharer *h, const void *dev_payload, int64_t payload_bytes) { return h->impl.mgpuRunPayload(dev_payload, payload_bytes); }
int64_t gss_mgpu_hits_to_device(gss_sharer *h, void *dev_dst, int64_t cap_records) { return h->impl.mgpuHitsToDevice(dev_dst, cap_records); }
int gss_mgpu_enqueue_payload(gss_sharer *h, const void *dev_payload, int64_t valid_bytes) { return h->impl.mgpuEnqueuePayload(dev_payload, valid_bytes); }
void gss_mgpu_redo_payload(gss_sharer *h, const void *dev_payload, int64_t total_bytes) { h->impl.mgpuRedoPayload(dev_payload, total_bytes); }
int64_t gss_mgpu_enqueue_result(gss_sharer *h, void *dev_dst, int64_t cap_records) { return h->impl.mgpuEnqueueResult(dev_dst, cap_records); }
int gss_mgpu_finish(gss_sharer *h) { return h->impl.mgpuFinish(); }
void gss_mgpu_import_gathered(gss_sharer *h, const void *dev_gathered, int world, int64_t slot_bytes, const int64_t *counts) {
    h->impl.mgpuImportGathered(dev_gathered, world, slot_bytes, counts);
}
int64_t gss_peer_init(gss_sharer *h, int rank, int world, int64_t payload_cap, int64_t slot_hits, void *blob_out, int64_t blob_cap) {
    return h->impl.peerInit(rank, world, payload_cap, slot_hits, blob_out, blob_cap);
}
void gss_peer_connect(gss_sharer *h, const void *blobs, int64_t blob_bytes) { h->impl.peerConnect(blobs, blob_bytes); }
int gss_peer_enqueue(gss_sharer *h) { return h->impl.peerEnqueue(); }
int64_t gss_peer_finish(gss_sharer *h) { return h->impl.peerFinish(); }
void gss_debug_set_peer_records(gss_sharer *h, int on) { h->impl.setPeerRecords(on != 0); }
void gss_set_stream(gss_sharer *h, void *cuda_stream) { h->impl.setStream(cuda_stream); }
int64_t gss_mgpu_wait(gss_sharer *h, const gss_raw_hit **hits) {
    if (!hits) return h->impl.mgpuWait(nullptr);
    const gss::HitRecord *p = nullptr;
    int64_t n = h->impl.mgpuWait(&p);
    *hits = reinterpret_cast<const gss_raw_hit *>(p);
    return n;
}
void gss_mgpu_import(gss_sharer *h, const gss_raw_hit *hits, int64_t n) {
    h->impl.mgpuImport(reinterpret_cast<const gss::HitRecord *>(hits), n);
}
const char *gss_version(void) { return "gpushare_b200 0.1 sm_100a"; }

} // extern "C"
