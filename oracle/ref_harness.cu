/*
 * ref_harness.cu -- TEST INFRASTRUCTURE.  A small C ABI around the UNMODIFIED reference
 * library (/root/reference/gpuShareLib, compiled in place by oracle/Makefile into
 * oracle/_ref/libgpushare_ref.so) so that tests and bench.py can drive the reference's own GPU
 * checker with the same calls as libgpushare_b200.so and read back its exact
 * (clause, solver, mask) triples.  Nothing here is product code and nothing of the reference
 * is copied: this file only includes its headers.
 *
 * The triples come from GpuRunner::reportedCls (gpuShareLib/GpuRunner.cuh:56, private, hence
 * the access hack below) and are mapped from GpuCref to the GpuClauseId with
 * HostClauses::getClause (gpuShareLib/Clauses.cu:365-367), as SURVEY.md 8(c) prescribes.
 */
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <algorithm>

#define private public
#define protected public
#include "gpuShareLib/GpuRunner.cuh"
#include "gpuShareLib/GpuClauseSharerImpl.cuh"
#include "gpuShareLib/Clauses.cuh"
#include "gpuShareLib/Assigs.cuh"
#include "gpuShareLib/Reported.cuh"
#undef private
#undef protected

using namespace GpuShare;

struct ref_hit {
    int64_t clause_id;
    int32_t solver_id;
    uint32_t mask;
};

struct ref_options {
    int gpuBlockCountGuideline, gpuThreadsPerBlockGuideline, minGpuLatencyMicros, verbosity;
    double clauseActivityDecay;
    int quickProf, initReportCountPerCategory, maxPageLockedMemory;
};

struct ref_sharer {
    GpuClauseSharerImpl impl;
    ref_sharer(GpuClauseSharerOptions o) : impl(o, [](const std::string &) {}) {}
};

extern "C" {

ref_sharer *ref_create(const ref_options *o) {
    GpuClauseSharerOptions opts;
    opts.gpuBlockCountGuideline = o->gpuBlockCountGuideline;
    opts.gpuThreadsPerBlockGuideline = o->gpuThreadsPerBlockGuideline;
    opts.minGpuLatencyMicros = o->minGpuLatencyMicros;
    opts.verbosity = o->verbosity;
    opts.clauseActivityDecay = o->clauseActivityDecay;
    opts.quickProf = o->quickProf != 0;
    opts.initReportCountPerCategory = o->initReportCountPerCategory;
    opts.maxPageLockedMemory = o->maxPageLockedMemory;
    return new ref_sharer(opts);
}
void ref_destroy(ref_sharer *h) { delete h; }
void ref_gpu_run(ref_sharer *h) { h->impl.gpuRun(); }
void ref_reduce_db(ref_sharer *h) { h->impl.reduceDb(); }
void ref_set_var_count(ref_sharer *h, int n) { h->impl.setVarCount(n); }
void ref_set_cpu_solver_count(ref_sharer *h, int n) { h->impl.setCpuSolverCount(n); }
int64_t ref_add_clause(ref_sharer *h, int solver, const int *lits, int n) { return h->impl.addClause(solver, (int *)lits, n); }
int64_t ref_add_clauses_bulk(ref_sharer *h, const int64_t *offsets, const int *lits, int64_t n) {
    int64_t first = -1;
    for (int64_t c = 0; c < n; c++) {
        int64_t id = h->impl.addClause(-1, (int *)(lits + offsets[c]), (int)(offsets[c + 1] - offsets[c]));
        if (c == 0) first = id;
        // "HClauses is designed to copy clauses in small chunks" (perfTest.cu:99-103)
        if (c % 5000 == 4999) h->impl.gpuRun();
    }
    return first;
}
int ref_try_set_solver_values(ref_sharer *h, int s, const int *lits, int n) { return h->impl.trySetSolverValues(s, (int *)lits, n) ? 1 : 0; }
void ref_unset_solver_values(ref_sharer *h, int s, const int *lits, int n) { h->impl.unsetSolverValues(s, (int *)lits, n); }
int64_t ref_try_send_assignment(ref_sharer *h, int s) { return h->impl.trySendAssignment(s); }
int ref_pop_reported_clause(ref_sharer *h, int s, int **lits, int *count, int64_t *id) {
    int *l;
    int c;
    long i;
    if (!h->impl.popReportedClause(s, l, c, i)) return 0;
    *lits = l;
    *count = c;
    *id = i;
    return 1;
}
int64_t ref_get_global_stat(ref_sharer *h, int stat) { return h->impl.getGlobalStat((GlobalStats)stat); }
int64_t ref_get_one_solver_stat(ref_sharer *h, int s, int stat) { return h->impl.getOneSolverStat(s, (OneSolverStats)stat); }
int64_t ref_get_last_assig_all_reported(ref_sharer *h, int s) { return h->impl.getLastAssigAllReported(s); }
void ref_get_current_assignment(ref_sharer *h, int s, uint8_t *assig) { h->impl.getCurrentAssignment(s, assig); }

/* hits of the most recently gathered run, sorted by (clause id, solver) */
int64_t ref_last_hits(ref_sharer *h, ref_hit *out, int64_t cap) {
    std::vector<ReportedClause> &rep = h->impl.gpuRunner->reportedCls;
    std::vector<ref_hit> hits(rep.size());
    std::vector<Lit> lits;
    for (size_t i = 0; i < rep.size(); i++) {
        int id;
        h->impl.clauses->getClause(lits, id, rep[i].gpuCref);
        hits[i] = ref_hit{(int64_t)id, rep[i].solverId, rep[i].reportedAssignments};
    }
    std::sort(hits.begin(), hits.end(), [](const ref_hit &a, const ref_hit &b) {
        return a.clause_id != b.clause_id ? a.clause_id < b.clause_id : a.solver_id < b.solver_id;
    });
    int64_t n = (int64_t)hits.size();
    for (int64_t i = 0; i < n && i < cap; i++) out[i] = hits[i];
    return n;
}

} /* extern "C" */
