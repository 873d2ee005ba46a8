// assigs.h -- host side of the assignment tables: the per-solver slot state machine
// (reference OneSolverAssigs / HostAssigs, gpuShareLib/Assigs.{cuh,cu}; rules restated in
// SURVEY.md Appendix A).  A solver owns 32 assignment slots (slot = assignment id % 32); a
// frozen slot holds the solver's whole partial assignment at trySendAssignment time.  Per run
// the GPU thread ships one VarUpdate (all 32 slots of one variable) per variable touched since
// the previous run, plus the masks the kernels need.
#pragma once
#include "common.h"
#include "mem.h"
#include <memory>
#include <mutex>
#include <vector>

namespace gss {

struct AssigIds {
    int64_t start = 0;
    int32_t count = 0;
};

class SolverAssigs {
public:
    explicit SolverAssigs(int varCount);
    void setVarCount(int varCount);

    void lock() { lock_.lock(); }
    bool tryLock() { return lock_.try_lock(); }
    void unlock() { lock_.unlock(); }

    // all *Locked methods need the lock (reference Assigs.cu:163-201)
    void setVarLocked(int var, uint8_t val);
    bool isAssignmentAvailableLocked() const { return currentId_ != firstIdUsed_ + kSlots; }
    int64_t assignmentDoneLocked();
    void getCurrentAssignment(uint8_t *assig);

    // Collect this solver's share of a run (reference copyUpdatesLocked, Assigs.cu:243-300).
    // Appends the updates to `updates` and fills `p` and `ids`.  When fullRebuild is set the
    // update list covers EVERY variable (the device tables are being re-created).
    void collectLocked(HostBuf<VarUpdate> &updates, SolverRunParams &p, AssigIds &ids, bool fullRebuild);
    // the same in two steps, so that the solvers' deltas can be copied concurrently: the number of
    // pending delta records, then the copy into `dst` (room for exactly that many) + parameters
    size_t pendingUpdatesLocked() const { return updates().size(); }
    void collectIntoLocked(VarUpdate *dst, int32_t updStart, SolverRunParams &p, AssigIds &ids);
    void finishCollectLocked(int32_t updStart, int32_t updCount, SolverRunParams &p);
    // Zero-copy collect: hands out the delta buffer itself (page-locked when a device is present, so
    // the kernel that applies the deltas reads it in place over PCIe) and continues in the other one.
    // The caller guarantees that the run that read the other buffer has finished (runs of a sharer
    // finish in order and a run is finished before the next one is collected).  *pinned = false: the
    // buffer is ordinary memory (budget exhausted / no device) and must be copied by the caller.
    void takeUpdatesLocked(const VarUpdate *&ptr, int32_t updStart, SolverRunParams &p, AssigIds &ids, bool *pinned);

    void setAggBits(int start, int end) { startAggBit_ = start; endAggBit_ = end; }
    // device whose context page-locks a delta buffer that has to grow on a solver thread
    void setAllocDevice(int device) { allocDevice_ = device; }
    int64_t updatesSent() const { return updatesSent_; }
    int varCount() const { return (int)lastVarVal_.size(); }

private:
    static uint32_t maskFromTo(int64_t fromId, int64_t toId);

    std::mutex lock_;
    uint32_t notCompletedMask_ = ~0u; // slots that are free or still being written
    int64_t updatesSent_ = 0;
    std::vector<uint8_t> lastVarVal_; // the solver's current value of every variable
    HostBuf<VarUpdate> upd_[2];       // one per variable touched in the batch being built (two in rotation)
    int curUpd_ = 0;
    HostBuf<VarUpdate> &updates() { return upd_[curUpd_]; }
    const HostBuf<VarUpdate> &updates() const { return upd_[curUpd_]; }
    std::vector<int32_t> varToUpdatePos_;
    int64_t firstIdUsed_ = 0; // first assignment id not yet shipped to the GPU
    int64_t currentId_ = 0;   // id of the assignment being written
    int startAggBit_ = 0, endAggBit_ = 0;
    int allocDevice_ = -1;
};

class HostAssigs {
public:
    HostAssigs();
    void setVarCount(int varCount);
    void growSolvers(int count); // reference growSolverAssigs, Assigs.cu:402-426
    int solverCount() const { return (int)solvers_.size(); }
    int varCount() const { return varCount_; }
    SolverAssigs &solver(int s) { return *solvers_[s]; }

private:
    int varCount_ = 0;
    std::vector<std::unique_ptr<SolverAssigs>> solvers_;
};

} // namespace gss
