// GpuClauseSharerShim.cc -- the reference-side binding: implements the reference's abstract
// class GpuShare::GpuClauseSharer and its factory makeGpuClauseSharerPtr on top of the C ABI of
// libgpushare_b200.so.  It is compiled against the UNMODIFIED reference header
// (-I<GpuShareSat checkout>, gpuShareLib/GpuClauseSharer.h), so the vtable layout is whatever
// glucose-syrup and rel-newtech were compiled against.  Linking this object (or
// libgpushare_b200_shim.so) instead of gpuShareLib's .cu objects is the whole integration; see
// INTEGRATION.md.
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>

#include "gpuShareLib/GpuClauseSharer.h" // from the GpuShareSat checkout, not from this repo

#include "gpushare_b200.h"

namespace GpuShare {

namespace {

class B200ClauseSharer : public GpuClauseSharer {
    gss_sharer *h_;
    std::function<void(const std::string &)> log_;

    static void logThunk(const char *msg, void *ctx) {
        auto *self = static_cast<B200ClauseSharer *>(ctx);
        if (self->log_) self->log_(std::string(msg));
    }

public:
    B200ClauseSharer(GpuClauseSharerOptions o, std::function<void(const std::string &)> logFunc) : log_(logFunc) {
        gss_options c;
        c.gpuBlockCountGuideline = o.gpuBlockCountGuideline;
        c.gpuThreadsPerBlockGuideline = o.gpuThreadsPerBlockGuideline;
        c.minGpuLatencyMicros = o.minGpuLatencyMicros;
        c.verbosity = o.verbosity;
        c.clauseActivityDecay = o.clauseActivityDecay;
        c.quickProf = o.quickProf ? 1 : 0;
        c.initReportCountPerCategory = o.initReportCountPerCategory;
        c.maxPageLockedMemory = o.maxPageLockedMemory;
        h_ = gss_create(&c, &B200ClauseSharer::logThunk, this);
    }
    ~B200ClauseSharer() override { gss_destroy(h_); }

    void gpuRun() override { gss_gpu_run(h_); }
    void reduceDb() override { gss_reduce_db(h_); }
    long getAddedClauseCount() override { return gss_get_added_clause_count(h_); }
    long getAddedClauseCountAtLastReduceDb() override { return gss_get_added_clause_count_at_last_reduce_db(h_); }
    bool hasRunOutOfGpuMemoryOnce() override { return gss_has_run_out_of_gpu_memory_once(h_) != 0; }
    void getGpuMemInfo(size_t &free, size_t &total) override { gss_get_gpu_mem_info(h_, &free, &total); }
    int getGlobalStatCount() override { return gss_get_global_stat_count(h_); }
    long getGlobalStat(GlobalStats stat) override { return gss_get_global_stat(h_, (int)stat); }
    void writeClausesInCnf(FILE *file) override { gss_write_clauses_in_cnf(h_, file); }
    void setVarCount(int newCount) override { gss_set_var_count(h_, newCount); }
    long addClause(int solverId, int *lits, int count) override { return gss_add_clause(h_, solverId, lits, count); }
    void setCpuSolverCount(int count) override { gss_set_cpu_solver_count(h_, count); }
    const char *getOneSolverStatName(OneSolverStats stat) override { return gss_get_one_solver_stat_name(h_, (int)stat); }
    const char *getGlobalStatName(GlobalStats stat) override { return gss_get_global_stat_name(h_, (int)stat); }
    int getOneSolverStatCount() override { return gss_get_one_solver_stat_count(h_); }
    bool trySetSolverValues(int cpuSolverId, int *lits, int count) override {
        return gss_try_set_solver_values(h_, cpuSolverId, lits, count) != 0;
    }
    void unsetSolverValues(int cpuSolverId, int *lits, int count) override { gss_unset_solver_values(h_, cpuSolverId, lits, count); }
    long trySendAssignment(int cpuSolverId) override { return gss_try_send_assignment(h_, cpuSolverId); }
    bool popReportedClause(int cpuSolverId, int *&lits, int &count, long &gpuClauseId) override {
        int64_t id = 0;
        if (!gss_pop_reported_clause(h_, cpuSolverId, &lits, &count, &id)) return false;
        gpuClauseId = (long)id;
        return true;
    }
    long getLastAssigAllReported(int cpuSolverId) override { return gss_get_last_assig_all_reported(h_, cpuSolverId); }
    void getCurrentAssignment(int cpuSolverId, uint8_t *assig) override { gss_get_current_assignment(h_, cpuSolverId, assig); }
    long getOneSolverStat(int cpuSolverId, OneSolverStats stat) override { return gss_get_one_solver_stat(h_, cpuSolverId, (int)stat); }
};

} // namespace

// gpuShareLib/GpuClauseSharer.h:165
GpuClauseSharer *makeGpuClauseSharerPtr(GpuClauseSharerOptions opts, std::function<void(const std::string &str)> logFunc) {
    return new B200ClauseSharer(opts, logFunc);
}

} // namespace GpuShare
