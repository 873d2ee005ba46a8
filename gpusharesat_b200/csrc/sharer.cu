// sharer.cu -- see sharer.h.  The run pipeline keeps the reference's shape (start run k+1,
// then hand run k's hits to the solvers while the GPU works: GpuRunner.cu:211-261) with three
// differences: (1) the post-run "collapse to last slot" is deferred to the start of the next run,
// so the tables of a finished run stay intact and a run whose survivor/hit buffer overflowed is
// simply re-launched with larger buffers -- hits are never dropped (reference: Reporter.cuh:46-48
// drops them); (2) new clauses go up as plain copies of the touched tiles; (3) counters are 64-bit.
#include "sharer.h"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <atomic>
#include <thread>

namespace gss {

std::shared_ptr<Sharer::RunBufPool> makeRunBufPool(); // pipeline.cu

Sharer::Sharer(const gss_options &o, gss_log_fn log, void *logCtx, int workerOfDevice) : opts_(o) {
    isWorker_ = workerOfDevice >= 0;
    logger_.verbosity = o.verbosity;
    if (log) logger_.fn = [log, logCtx](const std::string &s) { log(s.c_str(), logCtx); };

    // no CPU fallback: without a usable device the library refuses to work
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0)
        GSS_DIE(std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); gpushare_b200 has no CPU fallback");
    const char *envDev = getenv("GPUSHARE_DEVICE");
    if (isWorker_) device_ = workerOfDevice;
    else if (envDev) device_ = atoi(envDev);
    else GSS_CUDA(cudaGetDevice(&device_));
    GSS_CUDA(cudaSetDevice(device_));
    cudaDeviceProp props;
    GSS_CUDA(cudaGetDeviceProperties(&props, device_));
    numSMs_ = props.multiProcessorCount;

    // defaults: GpuClauseSharerImpl.cu:41-63
    if (opts_.minGpuLatencyMicros < 0) opts_.minGpuLatencyMicros = 50;
    if (opts_.clauseActivityDecay < 0) opts_.clauseActivityDecay = 0.99999;
    size_t pinnedLimit;
    if (opts_.maxPageLockedMemory < 0) {
        size_t freeB, totalB;
        GSS_CUDA(cudaMemGetInfo(&freeB, &totalB));
        pinnedLimit = totalB / 3;
    } else {
        pinnedLimit = (size_t)opts_.maxPageLockedMemory;
    }
    if (opts_.clauseActivityDecay >= 1) GSS_DIE("Clause activity decay must be strictly smaller than 1");
    if (opts_.initReportCountPerCategory < 0) opts_.initReportCountPerCategory = 10;
    if (opts_.initReportCountPerCategory == 0) GSS_DIE("initReportCountPerCategory must not be 0");
    if (opts_.gpuThreadsPerBlockGuideline == 0) GSS_DIE("gpuThreadsPerBlockGuideline must not be 0");
    if (opts_.gpuBlockCountGuideline == 0) GSS_DIE("gpuBlockCountGuideline must not be 0");

    // The guidelines only steer the grid (GpuClauseSharer.h:26-29); -1 lets occupancy decide.
    dims_.blocks = opts_.gpuBlockCountGuideline > 0 ? opts_.gpuBlockCountGuideline : 0;
    int thr = opts_.gpuThreadsPerBlockGuideline > 0 ? opts_.gpuThreadsPerBlockGuideline : 256;
    thr = std::max(32, std::min(256, (thr / 32) * 32));
    dims_.threads = thr;
    int categories = opts_.gpuBlockCountGuideline > 0 ? opts_.gpuBlockCountGuideline : 2 * numSMs_;
    hitCap_ = (size_t)opts_.initReportCountPerCategory * (size_t)categories;
    survCap_ = hitCap_;

    GSS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto &s : slots_) {
        GSS_CUDA(cudaEventCreate(&s.evStart));
        GSS_CUDA(cudaEventCreate(&s.evH2DDone));
        GSS_CUDA(cudaEventCreate(&s.evBeforeCheck));
        GSS_CUDA(cudaEventCreate(&s.evAfterCheck));
        GSS_CUDA(cudaEventCreate(&s.evEnd));
        s.headHost.setPinnedLimit(pinnedLimit);
        s.updHost.setPinnedLimit(pinnedLimit);
        s.resHost.setPinnedLimit(pinnedLimit);
    }
    db_ = std::make_unique<ClauseDb>(opts_.clauseActivityDecay, logger_, pinnedLimit);
    // the reference hard-codes MAX_CL_SIZE 100 (BaseTypes.cuh:28); the knob is an environment variable
    // so that it is reachable through the unmodified GpuClauseSharer.h
    if (const char *e = getenv("GPUSHARE_MAX_CLAUSE_LEN")) db_->setMaxLen(atoi(e));
    assigs_ = std::make_unique<HostAssigs>();
    reported_ = std::make_unique<Reported>(*db_, oneSolverStats_);
    reported_->setPool(&pool_);
    db_->setDeviceActivities(true);
    denseSliced_ = getenv("GSS_DENSE_FLAT") == nullptr; // (bench: the round-1 dense kernel, 256-byte rows from HBM)
    fuseHeader_ = getenv("GPUSHARE_NO_FUSED_HEADER") == nullptr;
    if (getenv("GPUSHARE_HOST_REDUCE")) db_->setDeviceReduce(false); // round-1 reduceDb (host compaction + re-upload), for comparison
    reported_->setHostBumps(false);
    setCpuSolverCount(1);
    runBufs_ = makeRunBufPool();
    directEnabled_ = getenv("GPUSHARE_LEGACY_PIPELINE") == nullptr;
    if (const char *e = getenv("GPUSHARE_EAGER_RESULTS")) eagerResults_ = atoi(e) != 0;
    logger_.log(1, std::string("c gpushare_b200 on ") + props.name + ", " + std::to_string(numSMs_) + " SMs\n");
    // several devices behind this one sharer (the factory keeps the reference's signature: the count
    // comes from the environment)
    if (!isWorker_)
        if (const char *e = getenv("GPUSHARE_DEVICES"))
            if (atoi(e) > 1) multiInit(atoi(e));
}

Sharer::~Sharer() {
    multiShutdown();
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_); // nothing may still be using the buffers (or a peer's window)
    peerClose();
    if (bumpFlagEv_) cudaEventDestroy(bumpFlagEv_);
    for (auto &s : slots_) {
        cudaEventDestroy(s.evStart);
        cudaEventDestroy(s.evH2DDone);
        cudaEventDestroy(s.evBeforeCheck);
        cudaEventDestroy(s.evAfterCheck);
        cudaEventDestroy(s.evEnd);
    }
    if (stream_ && ownStream_) cudaStreamDestroy(stream_);
}

void Sharer::useDevice() { GSS_CUDA(cudaSetDevice(device_)); }

void Sharer::gpuMemInfo(size_t *freeB, size_t *totalB) {
    useDevice();
    GSS_CUDA(cudaMemGetInfo(freeB, totalB));
}

void Sharer::setVarCount(int n) {
    varCount_ = std::max(varCount_, n);
    assigs_->setVarCount(n);
    for (auto &w : workers_) w->setVarCount(n);
}

void Sharer::setCpuSolverCount(int n) {
    // GpuClauseSharerImpl.cu:96-103
    if (n > kMaxGroups * kMaxSolversPerGroup) GSS_DIE("too many cpu solvers (max 256)");
    for (auto &w : workers_) w->setCpuSolverCount(n);
    assigs_->growSolvers(n);
    for (int s = 0; s < assigs_->solverCount(); s++) assigs_->solver(s).setAllocDevice(device_);
    reported_->setSolverCount(n);
    if ((int)toUnset_.size() < n) toUnset_.resize(n);
    size_t c = oneSolverStats_.size();
    if ((size_t)n > c) {
        oneSolverStats_.resize(n);
        for (size_t i = c; i < (size_t)n; i++) oneSolverStats_[i].assign(S_COUNT, 0);
    }
}

int64_t Sharer::addClause(int solver, const int *lits, int n) {
    int64_t id;
    if (workers_.empty()) {
        id = db_->addClause(lits, n);
    } else { // every device's database takes the clause, under one lock so that the ids agree
        std::lock_guard<std::mutex> g(multiAddLock_);
        id = db_->addClause(lits, n);
        for (auto &w : workers_)
            if (w->db_->addClause(lits, n) != id) GSS_DIE("the clause databases of a multi-device sharer disagree");
    }
    // GpuClauseSharerImpl.cu:124-128 registers the echo suppression even for a rejected clause
    if (solver != -1) reported_->clauseWasAdded(solver, id);
    return id;
}

int64_t Sharer::addClausesBulk(const int64_t *offsets, const int *lits, int64_t n) {
    if (workers_.empty()) return db_->addClausesBulk(offsets, lits, n);
    std::lock_guard<std::mutex> g(multiAddLock_);
    const int64_t id = db_->addClausesBulk(offsets, lits, n);
    for (auto &w : workers_)
        if (w->db_->addClausesBulk(offsets, lits, n) != id) GSS_DIE("the clause databases of a multi-device sharer disagree");
    return id;
}

void Sharer::unsetPendingLocked(int solver) {
    std::vector<int> &u = toUnset_[solver];
    SolverAssigs &sa = assigs_->solver(solver);
    for (int l : u) sa.setVarLocked(litVar(l), V_UNDEF);
    oneSolverStats_[solver][S_varUpdatesSentToGpu] += u.size();
    u.clear();
}

bool Sharer::trySetSolverValues(int solver, const int *lits, int n) {
    // GpuClauseSharerImpl.cu:130-147: all or nothing
    SolverAssigs &sa = assigs_->solver(solver);
    bool ok = false;
    sa.lock();
    if (sa.isAssignmentAvailableLocked()) {
        unsetPendingLocked(solver);
        for (int i = 0; i < n; i++) sa.setVarLocked(litVar(lits[i]), litSign(lits[i]) ? V_FALSE : V_TRUE);
        oneSolverStats_[solver][S_varUpdatesSentToGpu] += n;
        ok = true;
    } else {
        oneSolverStats_[solver][S_failuresToFindAssig]++;
    }
    sa.unlock();
    return ok;
}

void Sharer::unsetSolverValues(int solver, const int *lits, int n) {
    // GpuClauseSharerImpl.cu:159-176: buffered when no slot is free
    SolverAssigs &sa = assigs_->solver(solver);
    sa.lock();
    if (sa.isAssignmentAvailableLocked()) {
        unsetPendingLocked(solver);
        for (int i = 0; i < n; i++) sa.setVarLocked(litVar(lits[i]), V_UNDEF);
        oneSolverStats_[solver][S_varUpdatesSentToGpu] += n;
    } else {
        toUnset_[solver].insert(toUnset_[solver].end(), lits, lits + n);
    }
    sa.unlock();
}

int64_t Sharer::trySendAssignment(int solver) {
    // GpuClauseSharerImpl.cu:178-191
    SolverAssigs &sa = assigs_->solver(solver);
    int64_t r = -1;
    sa.lock();
    if (sa.isAssignmentAvailableLocked()) {
        r = sa.assignmentDoneLocked();
        oneSolverStats_[solver][S_assigsSentToGpu]++;
        reported_->assigWasSent(solver, r);
    } else {
        oneSolverStats_[solver][S_failuresToFindAssig]++;
    }
    sa.unlock();
    return r;
}

bool Sharer::popReportedClause(int solver, int *&lits, int &count, int64_t &id) {
    return reported_->pop(solver, lits, count, id);
}

void Sharer::currentAssignment(int solver, uint8_t *assig) {
    // GpuClauseSharerImpl.cu:238-244: pending unsets are overlaid
    assigs_->solver(solver).getCurrentAssignment(assig);
    for (int l : toUnset_[solver]) assig[litVar(l)] = V_UNDEF;
}

int64_t Sharer::globalStat(int stat) {
    switch (stat) {
    case G_clauseTestsOnAssigs: {
        uint64_t v = globalStats_[stat];
        for (auto &w : workers_) v += w->globalStats_[stat];
        return (int64_t)v;
    }
    case G_gpuClauses: return db_->stats().clauses;
    case G_gpuClauseLengthSum: return db_->stats().lengthSum;
    case G_gpuClausesAdded: return db_->stats().added;
    case G_gpuReduceDbs: return db_->reduceDbCount();
    default: return (int64_t)globalStats_[stat];
    }
}

// --------------------------------------------------------------------------------------------
// run pipeline
// --------------------------------------------------------------------------------------------

void Sharer::gpuRun() {
    // GpuClauseSharerImpl.cu:83-94
    int64_t t0 = nowMicros();
    // clauses streamed in since the arenas were last put in first-literal order: once their share has grown
    // enough, re-sort on the device (clause indices change: finish the run in flight first, as for reduceDb)
    if (workers_.empty() && !peer_ && db_->resortDue()) {
        wholeRun(false);
        useDevice();
        materializeLastHits();
        db_->resortOnDevice(stream_);
        lastStarted_ = -1;
        resorts_++;
    }
    wholeRun(true);
    // The reference sleeps until minGpuLatencyMicros have passed and surfaces the run's hits in the NEXT
    // call (GpuRunner.cu:233-242).  Here the waiting time is used: a run that completes within it is
    // gathered and handed to the solvers in THIS call ("not guaranteed" either way, GpuClauseSharer.h:80-83),
    // which takes one whole call period off the import latency (BASELINE config 5).
    if (eagerResults_ && cur_ >= 0 && workers_.empty()) {
        RunSlot &slot = slots_[cur_];
        while (nowMicros() - t0 < opts_.minGpuLatencyMicros) {
            cudaError_t e = cudaEventQuery(slot.evEnd);
            if (e == cudaSuccess) {
                wholeRun(false);
                break;
            }
            if (e != cudaErrorNotReady) GSS_CUDA(e);
        }
    }
    int64_t passed = nowMicros() - t0;
    if (passed < opts_.minGpuLatencyMicros)
        std::this_thread::sleep_for(std::chrono::microseconds(opts_.minGpuLatencyMicros - passed));
}

void Sharer::reduceDb() {
    // GpuClauseSharerImpl.cu:105-110: no run may be in flight, clause indices change
    wholeRun(false);
    useDevice();
    materializeLastHits(); // clause indices are about to change
    TimeAdder t(globalStats_[G_timeSpentReduceGpuDb], true);
    if (workers_.empty()) {
        db_->reduceDb(stream_);
    } else {
        reduceDbMulti();
    }
    lastStarted_ = -1;
}

void Sharer::wholeRun(bool canStart) {
    if (!workers_.empty()) return wholeRunMulti(canStart);
    useDevice();
    RunSlot *prev = cur_ >= 0 ? &slots_[cur_] : nullptr;
    {
        PhaseTimer t(hostPhases_[0]);
        if (prev) { // run k is complete and every hit is on the host
            if (prev->direct) finishRunDirect(*prev);
            else finishRun(*prev);
        }
    }
    int startedSlot = -1;
    bool outOfMemory = false;
    if (canStart) {
        int next = cur_ >= 0 ? 1 - cur_ : (collapseSlot_ >= 0 ? 1 - collapseSlot_ : 0);
        PhaseTimer t(hostPhases_[1]);
        db_->drainPending();
        if (db_->stats().clauses > 0) { // GpuRunner.cu:284-288: nothing starts on an empty database
            if (startRun(slots_[next])) startedSlot = next;
            else outOfMemory = true;
        }
    }
    {
        PhaseTimer t(hostPhases_[2]);
        if (prev) { // overlaps with run k+1 on the GPU
            if (prev->direct) processResultsDirect(*prev);
            else processResults(*prev);
        }
    }
    cur_ = startedSlot;
    if (outOfMemory) {
        // GpuRunner.cu:243-246
        logger_.log(1, "c gpushare_b200: out of GPU memory, reducing the clause database\n");
        materializeLastHits();
        TimeAdder t(globalStats_[G_timeSpentReduceGpuDb], true);
        db_->reduceDb(stream_);
        lastStarted_ = -1;
        ranOutOfMemory_ = true;
    }
}

bool Sharer::ensureTables(bool &rebuild) {
    int needVars = std::max(std::max(varCount_, db_->maxVarPlusOne()), 1);
    int nSolvers = assigs_->solverCount();
    rebuild = false;
    if (tablesValid_ && needVars <= tables_.varCap && nSolvers == tableSolvers_) return true;
    // (Re)create the tables; their contents are rebuilt from the host state by a full update
    // list in this run (SolverAssigs::collectLocked with fullRebuild).
    int stride = nSolvers <= 2 ? nSolvers : (nSolvers + 3) / 4 * 4;
    int groups = (nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    int varCap = tablesValid_ ? std::max(needVars, tables_.varCap) : needVars;
    if (!a1_.tryReserve((size_t)groups * 2 * varCap, 0, stream_, true)) return false;
    if (!t2_.tryReserve((size_t)varCap * stride, 0, stream_, true)) return false;
    tables_.a1 = a1_.data();
    tables_.t2 = t2_.data();
    tables_.varCap = varCap;
    tables_.solverStride = stride;
    tables_.nGroups = groups;
    tableSolvers_ = nSolvers;
    tablesValid_ = true;
    launchFillTables(tables_, 0, stream_, &launches_);
    collapseSlot_ = -1; // superseded by the rebuild
    rebuild = true;
    return true;
}

void Sharer::ensureResultBuffers() {
    int groups = std::max(1, tables_.nGroups);
    resDev_.reserve(sizeof(Counters) + hitCap_ * sizeof(HitRecord), 0, stream_);
    survDev_.reserve((size_t)groups * survCap_, 0, stream_);
}

CheckArgs Sharer::checkArgs(const RunSlot &slot, int g, bool recs) const {
    CheckArgs a;
    if (recs) {
        a.solverCtr = const_cast<unsigned long long *>(slot.ctrDev.data());
        a.recKeys = const_cast<unsigned long long *>(slot.recKeys.data());
        a.recMasks = const_cast<uint32_t *>(slot.recMasks.data());
        a.recCap = slot.recCap;
        a.totalClauses = std::max<int64_t>(1, db_->localClauses());
    }
    a.dir = slot.dirDev();
    a.nDir = slot.nDir;
    a.totalTiles = slot.totalTiles;
    a.shardRank = db_->shardRank();
    a.shardWorld = db_->shardWorld();
    a.params = slot.paramsDev();
    a.groupBase = g * kMaxSolversPerGroup;
    a.groupSolvers = std::min(kMaxSolversPerGroup, slot.nSolvers - a.groupBase);
    a.aggStart = slot.aggStart[g];
    a.aggStartOnDevice = slot.aggOnDevice ? 1 : 0;
    a.tables = tables_;
    a.survivors = const_cast<Survivor *>(survDev_.data()) + (size_t)g * survCap_;
    a.survCap = (unsigned int)survCap_;
    a.hits = hitsOverride_ ? hitsOverride_ : (HitRecord *)(const_cast<uint8_t *>(resDev_.data()) + sizeof(Counters));
    a.hitCap = hitsOverride_ ? hitCapOverride_ : (unsigned int)hitCap_;
    a.counters = (Counters *)const_cast<uint8_t *>(resDev_.data());
    return a;
}

bool Sharer::launchCheckKernels(RunSlot &slot, bool dense, bool filterOnly) {
    // direct pipeline: k_exact appends to per-solver record lists (dense mode always uses the global hit buffer)
    const bool recs = slot.direct && !dense;
    if (slot.direct && !recs) ensureResultBuffers();
    if (slot.countersZeroed && recs) {
        slot.countersZeroed = false; // (k_apply_direct of this run has done it)
    } else {
        GSS_CUDA(cudaMemsetAsync(resDev_.data(), 0, sizeof(Counters), stream_));
        if (recs) GSS_CUDA(cudaMemsetAsync(slot.ctrDev.data(), 0, (size_t)slot.nSolvers * kRecBuckets * kCtrStride * sizeof(unsigned long long), stream_));
    }
    int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    int lastGroup = -1;
    for (int g = 0; g < groups; g++)
        if (slot.aggStart[g] != 0) lastGroup = g;
    bool published = false;
    for (int g = 0; g < groups; g++) {
        if (slot.aggStart[g] == 0) continue; // no frozen slot in this group
        CheckArgs a = checkArgs(slot, g, recs);
        if (filterOnly) launchFilterOnly(a, dims_, numSMs_, stream_, &launches_);
        else if (dense) {
            if (denseSliced_ && a.groupSolvers > 8) {
                // slices of this batch's tables, built once per batch and group (a real dense-mode library
                // would have k_apply_updates write this layout; here it is a copy, timed by gss_debug_time_check mode 7)
                const size_t per = (size_t)tables_.varCap * 32;
                t2Sliced_.reserve(per * (size_t)groups, 0, stream_);
                if (!(denseSlicesValid_ & (1u << g))) {
                    launchSliceTables(tables_, a.groupBase, a.groupSolvers, t2Sliced_.data() + per * (size_t)g, numSMs_, stream_, &launches_);
                    denseSlicesValid_ |= 1u << g;
                }
                launchCheckDenseSliced(a, t2Sliced_.data() + per * (size_t)g, dims_, numSMs_, stream_, &launches_);
            } else {
                launchCheckDense(a, dims_, numSMs_, stream_, &launches_);
            }
        }
        else {
            if (fusedPublish_.peerDone && g == lastGroup && a.totalTiles > 0) {
                a.peerHdr = fusedPublish_.peerHdr;
                a.peerDone = fusedPublish_.peerDone;
                a.peerTicket = fusedPublish_.peerTicket;
                a.peerSeq = fusedPublish_.peerSeq;
                a.peerGroups = groups;
                published = true;
            }
            launchCheck(a, dims_, numSMs_, stream_, &launches_);
        }
    }
    return published;
}

void Sharer::enqueueResultCopy(RunSlot &slot) {
    size_t bytes = sizeof(Counters) + resultChunk() * sizeof(HitRecord);
    slot.resHost.resize(bytes);
    GSS_CUDA(cudaMemcpyAsync(slot.resHost.data(), resDev_.data(), bytes, cudaMemcpyDeviceToHost, stream_));
    lastD2H_ = (int64_t)bytes;
}

// tables, clause upload, length directory (everything of a run that is local to this device)
bool Sharer::prepareRun(RunSlot &slot, bool &rebuild, int64_t &h2d) {
    GSS_CUDA(cudaEventRecord(slot.evStart, stream_));
    rebuild = false;
    denseSlicesValid_ = 0; // a new batch: the tables are about to change
    HostProf hpUp("  prepare: tables+upload+dir");
    // Removing clauses cannot make room for the assignment tables (V x S words): fail loudly instead
    // of reducing the database run after run.  Clause arenas that do not fit take the reduceDb path.
    if (!ensureTables(rebuild))
        GSS_DIE("out of device memory for the assignment tables (" + std::to_string(varCount_) + " variables x " +
                std::to_string(assigs_->solverCount()) + " solvers)");
    if (!db_->uploadDirty(stream_, &h2d)) return false;
    std::vector<LenDir> dir;
    slot.totalTiles = db_->buildDirectory(dir);
    slot.nDir = (int)dir.size();
    slot.nSolvers = assigs_->solverCount();
    slot.dirBytes = ((size_t)slot.nDir * sizeof(LenDir) + 15) / 16 * 16;
    slot.headHost.resize(slot.dirBytes + (size_t)slot.nSolvers * sizeof(SolverRunParams));
    memcpy(slot.headHost.data(), dir.data(), dir.size() * sizeof(LenDir));
    return true;
}

// reference HostAssigs::fillAssigsAsync, Assigs.cu:326-372: every solver's deltas and run masks
void Sharer::collectBatch(RunSlot &slot, bool rebuild) {
    SolverRunParams *params = (SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
    // the update staging starts with room for [PayloadHeader][SolverRunParams x nSolvers], so that
    // the multi-GPU path can ship header + parameters + deltas as one contiguous payload
    const size_t P = payloadPrefixRecords(slot.nSolvers);
    slot.updHost.clear();
    slot.updHost.append(P);
    slot.ids.assign(slot.nSolvers, AssigIds{});
    slot.assigCount = 0;
    TimeAdder t(globalStats_[G_timeSpentFillingAssigs], opts_.quickProf != 0);
    // a solver busy writing its assignment is skipped for this run (Assigs.cu:350);
    // during a table rebuild every solver must contribute, so wait for it
    std::vector<char> locked(slot.nSolvers, 0);
    std::vector<size_t> offset(slot.nSolvers + 1, 0);
    size_t total = 0;
    for (int s = 0; s < slot.nSolvers; s++) {
        SolverAssigs &sa = assigs_->solver(s);
        memset(&params[s], 0, sizeof(SolverRunParams));
        locked[s] = rebuild ? (sa.lock(), 1) : (sa.tryLock() ? 1 : 0);
        offset[s] = total;
        if (locked[s] && !rebuild) total += sa.pendingUpdatesLocked();
    }
    offset[slot.nSolvers] = total;
    if (rebuild) {
        for (int s = 0; s < slot.nSolvers; s++) {
            SolverRunParams &p = params[s];
            assigs_->solver(s).collectLocked(slot.updHost, p, slot.ids[s], true);
            p.updStart -= (int32_t)P;
        }
    } else {
        // every solver's deltas go to their own range of the staging buffer: the copies are
        // independent, large batches spread them over the worker pool
        VarUpdate *base = slot.updHost.append(total);
        auto one = [&](int s) {
            SolverRunParams &p = params[s];
            p.updStart = (int32_t)offset[s];
            if (locked[s]) assigs_->solver(s).collectIntoLocked(base + offset[s], (int32_t)offset[s], p, slot.ids[s]);
        };
        if (total >= 65536 && slot.nSolvers >= 4) pool_.get().parallelFor(slot.nSolvers, one);
        else for (int s = 0; s < slot.nSolvers; s++) one(s);
    }
    for (int s = 0; s < slot.nSolvers; s++) {
        if (!locked[s]) continue;
        assigs_->solver(s).unlock(); // by the thread that locked it
        slot.assigCount += slot.ids[s].count;
    }
    slot.nUpdates = (int64_t)(slot.updHost.size() - P);
}

// device half of a run.  The run parameters are already in slot.headHost; `updSrc` may be a host
// or a device pointer (multi-GPU: the broadcast payload), cudaMemcpyDefault sorts it out.
void Sharer::launchRun(RunSlot &slot, const void *updSrc, int64_t nUpdates, int64_t &h2d) {
    const SolverRunParams *params = (const SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
    int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    slot.aggStart.assign(groups, 0u);
    slot.aggOnDevice = false;
    slot.maxUpd = 0;
    for (int s = 0; s < slot.nSolvers; s++) {
        slot.aggStart[s / kMaxSolversPerGroup] |= params[s].usedAggBits;
        slot.maxUpd = std::max(slot.maxUpd, (int)params[s].updCount);
    }
    slot.nUpdates = nUpdates;
    slot.dense = dense_;

    slot.headDev.reserve(slot.headHost.size(), 0, stream_);
    GSS_CUDA(cudaMemcpyAsync(slot.headDev.data(), slot.headHost.data(), slot.headHost.size(), cudaMemcpyHostToDevice, stream_));
    h2d += (int64_t)slot.headHost.size();
    if (slot.nUpdates) {
        slot.updDev.reserve((size_t)slot.nUpdates, 0, stream_);
        GSS_CUDA(cudaMemcpyAsync(slot.updDev.data(), updSrc, (size_t)slot.nUpdates * sizeof(VarUpdate), cudaMemcpyDefault, stream_));
        h2d += slot.nUpdates * (int64_t)sizeof(VarUpdate);
    }
    ensureResultBuffers();
    GSS_CUDA(cudaEventRecord(slot.evH2DDone, stream_));

    // the previous batch collapses to its last slot first (deferred dSetAllAssigsToLast)
    if (collapseSlot_ >= 0) {
        RunSlot &c = slots_[collapseSlot_];
        launchCollapse(c.updDev.data(), c.paramsDev(), c.nSolvers, c.maxUpd, c.nUpdates, tables_, numSMs_, stream_, &launches_);
        collapseSlot_ = -1;
    }
    launchApplyUpdates(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_, stream_, &launches_);
    GSS_CUDA(cudaEventRecord(slot.evBeforeCheck, stream_));
    launchCheckKernels(slot, slot.dense);
    GSS_CUDA(cudaEventRecord(slot.evAfterCheck, stream_));
    enqueueResultCopy(slot);
    GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
    slot.inFlight = true;
    if (slot.nUpdates) collapseSlot_ = (int)(&slot - slots_);
    lastStarted_ = (int)(&slot - slots_);
    lastH2D_ = h2d;
}

bool Sharer::startRun(RunSlot &slot) {
    if (directEnabled_ && !dense_ && !peer_) return startRunDirect(slot);
    slot.direct = false;
    int64_t h2d = 0;
    bool rebuild = false;
    if (!prepareRun(slot, rebuild, h2d)) return false;
    {
        PhaseTimer t(hostPhases_[3]);
        collectBatch(slot, rebuild);
    }
    launchRun(slot, slot.updHost.data() + payloadPrefixRecords(slot.nSolvers), slot.nUpdates, h2d);
    return true;
}

// --------------------------------------------------------------------------------------------
// multi-GPU: one front-end (rank 0: solver threads, slot machines, hand-over) and one clause
// shard per device.  Per batch: rank 0 collects -> the payload (run parameters + deltas) is
// broadcast -> every rank runs its shard -> hits are gathered -> rank 0 hands them over.
// --------------------------------------------------------------------------------------------

int Sharer::nextSlot() const { return cur_ >= 0 ? 1 - cur_ : (collapseSlot_ >= 0 ? 1 - collapseSlot_ : 0); }

int Sharer::mgpuCollect(const void **params, int64_t *paramsBytes, const void **updates, int64_t *nUpdates) {
    useDevice();
    GSS_CHECK(cur_ < 0 && mgpuPending_ < 0);
    db_->drainPending();
    if (db_->stats().clauses == 0) return -1;
    RunSlot &slot = slots_[nextSlot()];
    bool rebuild = false;
    mgpuH2D_ = 0;
    if (!prepareRun(slot, rebuild, mgpuH2D_)) GSS_DIE("out of device memory (multi-GPU mode does not reduce the database by itself)");
    collectBatch(slot, rebuild);
    mgpuPending_ = (int)(&slot - slots_);
    mgpuRebuild_ = rebuild;
    *params = slot.headHost.data() + slot.dirBytes;
    *paramsBytes = (int64_t)slot.nSolvers * (int64_t)sizeof(SolverRunParams);
    *updates = slot.updHost.data() + payloadPrefixRecords(slot.nSolvers);
    *nUpdates = slot.nUpdates;
    return rebuild ? 1 : 0;
}

// rank 0: collect and ship the batch as ONE contiguous payload [header][params][deltas] into a
// caller-provided device buffer (the NCCL broadcast source), H2D on the library's stream
int64_t Sharer::mgpuCollectTo(void *devDst, int64_t capBytes) {
    const void *params, *updates;
    int64_t paramsBytes, nUpdates;
    int r = mgpuCollect(&params, &paramsBytes, &updates, &nUpdates);
    PayloadHeader hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.magic = kPayloadMagic;
    if (r < 0) {
        hdr.status = -1;
        hdr.totalBytes = sizeof(hdr);
        mgpuHdrHost_.resize(sizeof(hdr));
        memcpy(mgpuHdrHost_.data(), &hdr, sizeof(hdr));
        GSS_CUDA(cudaMemcpyAsync(devDst, mgpuHdrHost_.data(), sizeof(hdr), cudaMemcpyHostToDevice, stream_));
        return (int64_t)sizeof(hdr);
    }
    RunSlot &slot = slots_[mgpuPending_];
    const size_t P = payloadPrefixRecords(slot.nSolvers);
    hdr.status = r;
    hdr.nSolvers = slot.nSolvers;
    hdr.prefixRecords = (int32_t)P;
    hdr.nUpdates = nUpdates;
    hdr.totalBytes = (int64_t)((P + (size_t)nUpdates) * sizeof(VarUpdate));
    uint8_t *base = reinterpret_cast<uint8_t *>(slot.updHost.data());
    memcpy(base, &hdr, sizeof(hdr));
    memcpy(base + sizeof(hdr), params, (size_t)paramsBytes);
    if (hdr.totalBytes > capBytes) GSS_DIE("multi-GPU payload buffer too small");
    GSS_CUDA(cudaMemcpyAsync(devDst, base, (size_t)hdr.totalBytes, cudaMemcpyHostToDevice, stream_));
    mgpuH2D_ += hdr.totalBytes;
    return hdr.totalBytes;
}

void Sharer::mgpuRun(const void *params, int64_t paramsBytes, const void *updates, int64_t nUpdates, int rebuild) {
    useDevice();
    GSS_CHECK(cur_ < 0);
    int64_t h2d = 0;
    RunSlot *slot;
    if (mgpuPending_ >= 0) { // rank 0: the slot was prepared by mgpuCollect
        slot = &slots_[mgpuPending_];
        mgpuPending_ = -1;
        h2d = mgpuH2D_;
    } else {
        db_->drainPending();
        slot = &slots_[nextSlot()];
        bool myRebuild = false;
        if (!prepareRun(*slot, myRebuild, h2d)) GSS_DIE("out of device memory (multi-GPU mode)");
        if (myRebuild && !rebuild) GSS_DIE("multi-GPU ranks disagree about a table rebuild (call setVarCount / setCpuSolverCount identically on every rank)");
        slot->ids.assign(slot->nSolvers, AssigIds{});
        slot->assigCount = 0;
    }
    GSS_CHECK(paramsBytes == (int64_t)slot->nSolvers * (int64_t)sizeof(SolverRunParams));
    void *dst = slot->headHost.data() + slot->dirBytes;
    if (params != dst) {
        GSS_CUDA(cudaMemcpyAsync(dst, params, (size_t)paramsBytes, cudaMemcpyDefault, stream_));
        GSS_CUDA(cudaStreamSynchronize(stream_));
    }
    launchRun(*slot, updates, nUpdates, h2d);
    cur_ = (int)(slot - slots_);
}

// every rank: run the batch whose packed payload sits in device memory (after the broadcast)
int Sharer::mgpuRunPayload(const void *devPayload, int64_t payloadBytes) {
    useDevice();
    PayloadHeader hdr;
    if (mgpuPending_ >= 0) { // rank 0 wrote it itself: no device round trip at all
        RunSlot &slot = slots_[mgpuPending_];
        memcpy(&hdr, slot.updHost.data(), sizeof(hdr));
        if (hdr.status < 0) { mgpuPending_ = -1; return -1; }
        const uint8_t *base = static_cast<const uint8_t *>(devPayload);
        mgpuRun(slot.headHost.data() + slot.dirBytes, (int64_t)hdr.nSolvers * (int64_t)sizeof(SolverRunParams),
                base + (size_t)hdr.prefixRecords * sizeof(VarUpdate), hdr.nUpdates, hdr.status);
        return hdr.status;
    } else {
        GSS_CUDA(cudaMemcpyAsync(&hdr, devPayload, sizeof(hdr), cudaMemcpyDeviceToHost, stream_));
        GSS_CUDA(cudaStreamSynchronize(stream_));
    }
    if (hdr.magic != kPayloadMagic) GSS_DIE("multi-GPU payload header corrupt");
    if (hdr.status < 0) return -1;
    GSS_CHECK(hdr.totalBytes <= payloadBytes);
    const uint8_t *base = static_cast<const uint8_t *>(devPayload);
    mgpuRun(base + sizeof(hdr), (int64_t)hdr.nSolvers * (int64_t)sizeof(SolverRunParams),
            base + (size_t)hdr.prefixRecords * sizeof(VarUpdate), hdr.nUpdates, hdr.status);
    return hdr.status;
}

// receivers, asynchronous: enqueue the whole batch without ever looking at the payload on the host.
// Counts and masks are read by the kernels from the device copy of the run parameters; grids are
// sized from the upper bound the payload size gives.  Returns -1 when there is no clause yet.
int Sharer::mgpuEnqueuePayload(const void *devPayload, int64_t validBytes) {
    useDevice();
    GSS_CHECK(cur_ < 0 && mgpuPending_ < 0);
    db_->drainPending();
    if (db_->stats().clauses == 0) return -1;
    RunSlot &slot = slots_[nextSlot()];
    bool rebuild = false;
    int64_t h2d = 0;
    if (!prepareRun(slot, rebuild, h2d)) GSS_DIE("out of device memory (multi-GPU mode)");
    slot.ids.assign(slot.nSolvers, AssigIds{});
    slot.assigCount = 0;
    enqueueFromDevicePayload(slot, devPayload, validBytes, true, h2d);
    cur_ = (int)(&slot - slots_);
    return rebuild ? 1 : 0;
}

void Sharer::enqueueFromDevicePayload(RunSlot &slot, const void *devPayload, int64_t validBytes, bool collapsePrev, int64_t h2d) {
    const uint8_t *base = static_cast<const uint8_t *>(devPayload);
    const size_t P = payloadPrefixRecords(slot.nSolvers);
    int64_t nUpper = validBytes / (int64_t)sizeof(VarUpdate) - (int64_t)P;
    if (nUpper < 0) nUpper = 0;
    int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    slot.aggStart.assign(groups, ~0u);
    slot.aggOnDevice = true;
    slot.maxUpd = (int)std::min<int64_t>(nUpper, 1 << 30);
    slot.nUpdates = nUpper;
    slot.dense = dense_;
    slot.headDev.reserve(slot.headHost.size(), 0, stream_);
    GSS_CUDA(cudaMemcpyAsync(slot.headDev.data(), slot.headHost.data(), slot.dirBytes, cudaMemcpyHostToDevice, stream_));
    GSS_CUDA(cudaMemcpyAsync(slot.headDev.data() + slot.dirBytes, base + sizeof(PayloadHeader),
                             (size_t)slot.nSolvers * sizeof(SolverRunParams), cudaMemcpyDeviceToDevice, stream_));
    if (nUpper) {
        slot.updDev.reserve((size_t)nUpper, 0, stream_);
        GSS_CUDA(cudaMemcpyAsync(slot.updDev.data(), base + P * sizeof(VarUpdate), (size_t)nUpper * sizeof(VarUpdate),
                                 cudaMemcpyDeviceToDevice, stream_));
    }
    ensureResultBuffers();
    GSS_CUDA(cudaEventRecord(slot.evH2DDone, stream_));
    if (collapsePrev && collapseSlot_ >= 0) {
        RunSlot &c = slots_[collapseSlot_];
        launchCollapse(c.updDev.data(), c.paramsDev(), c.nSolvers, c.maxUpd, c.nUpdates, tables_, numSMs_, stream_, &launches_);
        collapseSlot_ = -1;
    }
    launchApplyUpdates(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_, stream_, &launches_);
    GSS_CUDA(cudaEventRecord(slot.evBeforeCheck, stream_));
    launchCheckKernels(slot, slot.dense);
    GSS_CUDA(cudaEventRecord(slot.evAfterCheck, stream_));
    enqueueResultCopy(slot);
    GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
    slot.inFlight = true;
    collapseSlot_ = (int)(&slot - slots_);
    lastStarted_ = (int)(&slot - slots_);
    lastH2D_ = h2d;
}

// receivers: the broadcast was truncated (the batch outgrew the predicted size); run the same batch
// again on the complete payload.  apply is idempotent and the previous batch is already collapsed.
void Sharer::mgpuRedoPayload(const void *devPayload, int64_t totalBytes) {
    useDevice();
    GSS_CHECK(cur_ < 0 && mgpuLast_ >= 0);
    RunSlot &slot = slots_[mgpuLast_];
    GSS_CUDA(cudaEventRecord(slot.evStart, stream_));
    collapseSlot_ = -1;
    enqueueFromDevicePayload(slot, devPayload, totalBytes, false, 0);
    cur_ = mgpuLast_;
}

// enqueue [64 B header {nHits, overflow}][hits x capRecords] of the run in flight into devDst
int64_t Sharer::mgpuEnqueueResult(void *devDst, int64_t capRecords) {
    useDevice();
    int slotIdx = cur_ >= 0 ? cur_ : mgpuLast_;
    GSS_CHECK(slotIdx >= 0);
    RunSlot &slot = slots_[slotIdx];
    int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    launchFinalize((const Counters *)resDev_.data(), (unsigned int)hitCap_, (unsigned int)survCap_, groups,
                   static_cast<long long *>(devDst), stream_, &launches_);
    int64_t n = std::min<int64_t>(capRecords, (int64_t)hitCap_);
    if (n > 0)
        GSS_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(devDst) + 64, resDev_.data() + sizeof(Counters),
                                 (size_t)n * sizeof(HitRecord), cudaMemcpyDeviceToDevice, stream_));
    return 64 + n * (int64_t)sizeof(HitRecord);
}

// after the caller has synchronised: bookkeeping of the finished run; returns 1 when this rank had
// to run again with larger buffers (what it contributed to the gather is stale), else 0
int Sharer::mgpuFinish() {
    useDevice();
    GSS_CHECK(cur_ >= 0);
    finishReran_ = false;
    finishRun(slots_[cur_], false);
    mgpuLast_ = cur_;
    cur_ = -1;
    return finishReran_ ? 1 : 0;
}

int64_t Sharer::mgpuHitsToDevice(void *devDst, int64_t capRecords) {
    useDevice();
    int64_t n = std::min<int64_t>((int64_t)hits_.size(), capRecords);
    if (n > 0)
        GSS_CUDA(cudaMemcpyAsync(devDst, resDev_.data() + sizeof(Counters), (size_t)n * sizeof(HitRecord),
                                 cudaMemcpyDeviceToDevice, stream_));
    return (int64_t)hits_.size();
}

void Sharer::setStream(void *stream) {
    useDevice();
    GSS_CUDA(cudaStreamSynchronize(stream_));
    if (ownStream_) cudaStreamDestroy(stream_);
    stream_ = static_cast<cudaStream_t>(stream);
    ownStream_ = false;
}

int64_t Sharer::mgpuWait(const HitRecord **hits) {
    useDevice();
    GSS_CHECK(cur_ >= 0);
    // hits == nullptr: the caller gathers the hits from device memory (gss_mgpu_hits_to_device),
    // only the count is needed on the host
    finishRun(slots_[cur_], hits != nullptr, false);
    mgpuLast_ = cur_;
    cur_ = -1;
    if (hits) *hits = hits_.data();
    return (int64_t)hits_.size();
}

// rank 0: the all-gathered result blocks are still on the device ([64 B header][hits] per rank).
// Concatenate the valid parts on the device and treat the union like the hit list of a local run:
// large unions are sorted / resolved on the device (every rank holds the whole clause arena).
void Sharer::mgpuImportGathered(const void *devGathered, int world, int64_t slotBytes, const int64_t *counts) {
    useDevice();
    GSS_CHECK(mgpuLast_ >= 0);
    RunSlot &slot = slots_[mgpuLast_];
    size_t total = 0;
    for (int r = 0; r < world; r++) total += (size_t)counts[r];
    unionDev_.reserve(std::max<size_t>(total, 1) * sizeof(HitRecord), 0, stream_);
    size_t off = 0;
    const uint8_t *g = static_cast<const uint8_t *>(devGathered);
    for (int r = 0; r < world; r++) {
        if (counts[r] == 0) continue;
        GSS_CUDA(cudaMemcpyAsync(unionDev_.data() + off * sizeof(HitRecord), g + (size_t)r * (size_t)slotBytes + 64,
                                 (size_t)counts[r] * sizeof(HitRecord), cudaMemcpyDeviceToDevice, stream_));
        off += (size_t)counts[r];
    }
    postValid_ = false;
    finishedD2H_ = 0;
    if (total >= kPostprocessHits) {
        hits_.clear();
        postprocessOnDevice(slot, total, (const HitRecord *)unionDev_.data());
        parkHitsForBump((const uint8_t *)postDev_.data() + postSortedOffset_, (int)sizeof(SortedHit), total);
    } else {
        hits_.resize(total);
        if (total) {
            GSS_CUDA(cudaMemcpyAsync(hits_.data(), unionDev_.data(), total * sizeof(HitRecord), cudaMemcpyDeviceToHost, stream_));
            GSS_CUDA(cudaStreamSynchronize(stream_));
        }
        finishedD2H_ = (int64_t)(total * sizeof(HitRecord));
        parkHitsForBump(unionDev_.data(), (int)sizeof(HitRecord), total);
    }
    processResults(slot);
}

void Sharer::mgpuImport(const HitRecord *hits, int64_t n) {
    GSS_CHECK(mgpuLast_ >= 0);
    if (hits != hits_.data()) hits_.assign(hits, hits + n);
    postValid_ = false;
    parkHitsForBump(hits_.data(), (int)sizeof(HitRecord), hits_.size()); // the union lives on the host
    processResults(slots_[mgpuLast_]);
}

void Sharer::finishRun(RunSlot &slot, bool fetchAllHits, bool allowPostprocess) {
    {
        PhaseTimer t(hostPhases_[4]);
        GSS_CUDA(cudaEventSynchronize(slot.evEnd));
    }
    float msCopy = 0, msApply = 0, msCheck = 0, msTotal = 0;
    cudaEventElapsedTime(&msCopy, slot.evStart, slot.evH2DDone);
    cudaEventElapsedTime(&msApply, slot.evH2DDone, slot.evBeforeCheck);
    cudaEventElapsedTime(&msCheck, slot.evBeforeCheck, slot.evAfterCheck);
    cudaEventElapsedTime(&msTotal, slot.evStart, slot.evEnd);
    lastTimes_[0] = msCopy * 1000.0;
    lastTimes_[1] = msApply * 1000.0;
    lastTimes_[2] = msCheck * 1000.0;
    lastTimes_[3] = msTotal * 1000.0;
    haveTimes_ = true;
    if (opts_.quickProf) globalStats_[G_timeSpentTestingClauses] += (uint64_t)(msCheck * 1000.0f);

    Counters c;
    for (int attempt = 0;; attempt++) {
        memcpy(&c, slot.resHost.data(), sizeof(c));
        size_t maxSurv = 0;
        for (int g = 0; g < kMaxGroups; g++) maxSurv = std::max(maxSurv, (size_t)c.nSurvivors[g]);
        if (c.nHits <= hitCap_ && maxSurv <= survCap_) break;
        GSS_CHECK(attempt < 8);
        finishReran_ = true;
        // Overflow: the tables of this run are still intact (collapse is deferred), so grow the
        // buffers and run the check again.  Nothing is dropped.
        if (c.nHits > hitCap_) hitCap_ = std::max(hitCap_ * 2, (size_t)c.nHits + c.nHits / 4);
        if (maxSurv > survCap_) survCap_ = std::max(survCap_ * 2, maxSurv + maxSurv / 4);
        hitCap_ = std::max(hitCap_, survCap_ / 4);
        ensureResultBuffers();
        launchCheckKernels(slot, slot.dense);
        enqueueResultCopy(slot);
        GSS_CUDA(cudaStreamSynchronize(stream_));
    }
    globalStats_[G_clauseTestsOnAssigs] += c.exactTests;
    slot.inFlight = false;
    postValid_ = false;
    finishedD2H_ = (int64_t)slot.resHost.size();
    if (fetchAllHits && allowPostprocess && c.nHits >= kPostprocessHits) {
        // large result: sort and resolve it on the device, fetch only the finished product
        hits_.clear();
        PhaseTimer t(hostPhases_[5]);
        postprocessOnDevice(slot, c.nHits);
        parkHitsForBump((const uint8_t *)postDev_.data() + postSortedOffset_, (int)sizeof(SortedHit), c.nHits);
        return;
    }
    hits_.resize(c.nHits);
    size_t first = std::min((size_t)c.nHits, (slot.resHost.size() - sizeof(Counters)) / sizeof(HitRecord));
    if (first) memcpy(hits_.data(), slot.resHost.data() + sizeof(Counters), first * sizeof(HitRecord));
    if (c.nHits > first && fetchAllHits) {
        size_t rest = c.nHits - first;
        GSS_CUDA(cudaMemcpyAsync(hits_.data() + first, resDev_.data() + sizeof(Counters) + first * sizeof(HitRecord),
                                 rest * sizeof(HitRecord), cudaMemcpyDeviceToHost, stream_));
        GSS_CUDA(cudaStreamSynchronize(stream_));
        finishedD2H_ += (int64_t)(rest * sizeof(HitRecord));
    }
    if (fetchAllHits) parkHitsForBump(resDev_.data() + sizeof(Counters), (int)sizeof(HitRecord), c.nHits);
}

// The records are on the device already (result buffer or post-processing buffer); both are reused
// by the next run, so copy them aside (device to device).
void Sharer::parkHitsForBump(const void *devRecs, int stride, size_t n) {
    bumpN_ = n;
    bumpStride_ = stride;
    if (n == 0) return;
    bumpRecs_.reserve(n * (size_t)stride, 0, stream_);
    GSS_CUDA(cudaMemcpyAsync(bumpRecs_.data(), devRecs, n * (size_t)stride, cudaMemcpyDefault, stream_));
}

void Sharer::bumpParkedHits() {
    // a bump of the previous batch overflowed an activity: rescale like Clauses.cu:231-237 does
    if (bumpFlagPending_) {
        // wait for that bump's flag only -- not for whatever has been queued since (the next run)
        GSS_CUDA(cudaEventSynchronize(bumpFlagEv_));
        if (bumpFlagHost_[0]) {
            // Clauses.cu:231-237 rescales activities and increment together: the device copies must be
            // scaled down BEFORE the next bump uses the scaled-down increment
            db_->rescaleAfterDeviceOverflow();
            db_->applyPendingDeviceRescales(stream_);
        }
        bumpFlagPending_ = false;
    }
    if (bumpN_ == 0) return;
    std::vector<LenDir> dir;
    db_->buildDirectory(dir);
    bumpDirHost_.resize(dir.size() * sizeof(LenDir));
    memcpy(bumpDirHost_.data(), dir.data(), dir.size() * sizeof(LenDir));
    bumpDirDev_.reserve(bumpDirHost_.size(), 0, stream_);
    GSS_CUDA(cudaMemcpyAsync(bumpDirDev_.data(), bumpDirHost_.data(), bumpDirHost_.size(), cudaMemcpyHostToDevice, stream_));
    bumpFlagDev_.reserve(1, 0, stream_);
    bumpFlagHost_.resize(1);
    GSS_CUDA(cudaMemsetAsync(bumpFlagDev_.data(), 0, sizeof(int), stream_));
    launchBumpActivity(bumpRecs_.data(), bumpStride_, (unsigned int)bumpN_, (const LenDir *)bumpDirDev_.data(), (int)dir.size(),
                       db_->activityIncrement(), bumpFlagDev_.data(), stream_, &launches_);
    GSS_CUDA(cudaMemcpyAsync(bumpFlagHost_.data(), bumpFlagDev_.data(), sizeof(int), cudaMemcpyDeviceToHost, stream_));
    if (!bumpFlagEv_) GSS_CUDA(cudaEventCreateWithFlags(&bumpFlagEv_, cudaEventDisableTiming));
    GSS_CUDA(cudaEventRecord(bumpFlagEv_, stream_));
    bumpFlagPending_ = true;
    bumpN_ = 0;
}

void Sharer::postprocessOnDevice(RunSlot &slot, size_t n, const HitRecord *hitsDevOverride) {
    auto align = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t tempBytes = postprocessTempBytes((unsigned int)n);
    size_t off = 0;
    const size_t oKeysIn = off; off += align(n * 8);
    const size_t oKeysOut = off; off += align(n * 8);
    const size_t oValsIn = off; off += align(n * 4);
    const size_t oValsOut = off; off += align(n * 4);
    const size_t oLitPos = off; off += align((n + 1) * 8);
    const size_t oSorted = off; off += align(n * sizeof(SortedHit));
    const size_t oTemp = off; off += align(tempBytes);
    postDev_.reserve(off, 0, stream_);
    uint8_t *base = postDev_.data();
    PostBuffers b;
    b.keysIn = (unsigned long long *)(base + oKeysIn);
    b.keysOut = (unsigned long long *)(base + oKeysOut);
    b.valsIn = (unsigned int *)(base + oValsIn);
    b.valsOut = (unsigned int *)(base + oValsOut);
    b.litPos = (long long *)(base + oLitPos);
    b.sorted = (SortedHit *)(base + oSorted);
    postSortedOffset_ = oSorted;
    b.temp = base + oTemp;
    b.tempBytes = tempBytes;
    b.lits = nullptr;
    b.litCap = 0;
    const HitRecord *hitsDev = hitsDevOverride ? hitsDevOverride : (const HitRecord *)(resDev_.data() + sizeof(Counters));
    auto bitsFor = [](uint64_t maxValue) { int b = 1; while (b < 63 && (maxValue >> b)) b++; return b; };
    launchPostSort(hitsDev, (unsigned int)n, b, bitsFor((uint64_t)std::max(1, slot.nSolvers - 1)), bitsFor((uint64_t)db_->maxLen()),
                   bitsFor((uint64_t)std::max<int64_t>(1, db_->stats().clauses)), stream_, &launches_);
    // The literal stream is emitted into a buffer sized from the previous large result (its total is
    // only known on the device): sort, scan, emit and all three copies are queued back to back and
    // there is ONE synchronisation.  A result that outgrew the guess is emitted again.
    postTotalHost_.resize(1);
    GSS_CUDA(cudaMemcpyAsync(postTotalHost_.data(), b.litPos + n, sizeof(long long), cudaMemcpyDeviceToHost, stream_));
    int64_t guess = std::max<int64_t>(postLitGuess_ + postLitGuess_ / 8, (int64_t)n * 3);
    postSortedHost_.resize(n);
    int64_t total = 0;
    for (int attempt = 0;; attempt++) {
        postLitsDev_.reserve((size_t)std::max<int64_t>(guess, 1), 0, stream_);
        b.lits = postLitsDev_.data();
        b.litCap = guess;
        launchPostEmit(hitsDev, (unsigned int)n, slot.dirDev(), slot.nDir, b, stream_, &launches_);
        if (attempt == 0)
            GSS_CUDA(cudaMemcpyAsync(postSortedHost_.data(), b.sorted, n * sizeof(SortedHit), cudaMemcpyDeviceToHost, stream_));
        postLitsHost_.resize((size_t)std::max<int64_t>(guess, 1));
        GSS_CUDA(cudaMemcpyAsync(postLitsHost_.data(), b.lits, (size_t)guess * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        GSS_CUDA(cudaStreamSynchronize(stream_));
        total = postTotalHost_[0];
        finishedD2H_ += (int64_t)((size_t)guess * sizeof(int32_t));
        if (total <= guess) break;
        GSS_CHECK(attempt == 0);
        guess = total;
    }
    postLitGuess_ = total;
    finishedD2H_ += (int64_t)(n * sizeof(SortedHit) + sizeof(long long));
    postValid_ = true;
    postN_ = n;
    postLits_ = total;
}

void Sharer::processResults(RunSlot &slot) {
    // reference gatherGpuRunResults, GpuRunner.cu:360-383 (64-bit arithmetic)
    int64_t clCount = db_->stats().clauses;
    globalStats_[G_gpuRuns]++;
    globalStats_[G_totalAssigClauseTested] += (uint64_t)clCount * (uint64_t)slot.assigCount;
    globalStats_[G_clauseTestsOnGroups] += (uint64_t)clCount;
    globalStats_[G_gpuReports] += postValid_ ? postN_ : hits_.size();
    lastHitsValid_ = false; // gss_debug_last_hits converts hits_ on demand
    lastDirect_ = nullptr;
    bumpParkedHits();
    TimeAdder t(globalStats_[G_timeSpentFillingReported], opts_.quickProf != 0);
    if (postValid_) reported_->handOverSorted(postSortedHost_.data(), postN_, postLitsHost_.data(), postLits_, slot.ids, slot.nSolvers);
    else reported_->handOver(hits_, slot.ids, slot.nSolvers);
}

void Sharer::materializeLastHits() {
    if (lastHitsValid_) return;
    if (lastDirect_) {
        // direct pipeline: ids are in the result buffers (host), the masks of the sorted record lists on the device(s)
        lastHits_.clear();
        const int idx = (int)(lastDirect_ - slots_);
        appendDirectHits(*lastDirect_, lastHits_);
        for (auto &w : workers_) w->appendDirectHits(w->slots_[idx], lastHits_);
        for (auto &fb : lastForeign_) { // other ranks' results (multi-process exchange): masks lie next to the ids
            const RunHdr *h = fb->hdr();
            if (!fb->withRecords) continue; // (that rank kept its masks: gss_debug_set_peer_records)
            for (int s = 0; s < kMaxSolvers; s++) {
                const RunHdr::PerSolver &ps = h->solver[s];
                for (int32_t i = 0; i < ps.n; i++)
                    lastHits_.push_back(gss_hit{fb->ids()[ps.entryBase + i], s, fb->masks()[ps.entryBase + i]});
            }
        }
        useDevice();
    } else if (postValid_) {
        lastHits_.resize(postN_);
        for (size_t i = 0; i < postN_; i++)
            lastHits_[i] = gss_hit{postSortedHost_[i].id, postSortedHost_[i].solver, postSortedHost_[i].mask};
    } else {
        lastHits_.resize(hits_.size());
        for (size_t i = 0; i < hits_.size(); i++)
            lastHits_[i] = gss_hit{db_->clauseId(hits_[i].len, hits_[i].idx), hits_[i].solver, hits_[i].mask};
    }
    std::sort(lastHits_.begin(), lastHits_.end(), [](const gss_hit &a, const gss_hit &b) {
        return a.clause_id != b.clause_id ? a.clause_id < b.clause_id : a.solver_id < b.solver_id;
    });
    lastHitsValid_ = true;
}

int64_t Sharer::lastHits(gss_hit *out, int64_t cap) {
    materializeLastHits();
    int64_t n = (int64_t)lastHits_.size();
    if (out && cap > 0) memcpy(out, lastHits_.data(), (size_t)std::min(n, cap) * sizeof(gss_hit));
    return n;
}

// mode: 0 production check (k_filter + k_exact), 1 dense kernel, 2 k_filter alone, 3 k_exact alone
// (on the survivors of the last filter), 4 k_apply_updates alone (idempotent), 5 k_collapse alone
// (followed by one apply, which restores the tables).  Average microseconds per launch.
double Sharer::timeCheck(int iters, int mode) {
    useDevice();
    if (lastStarted_ < 0 || iters < 1) return -1.0;
    RunSlot &slot = slots_[lastStarted_];
    GSS_CUDA(cudaStreamSynchronize(stream_));
    const bool dense = mode == 1, filterOnly = mode == 2;
    int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    auto once = [&]() {
        if (mode <= 2) {
            launchCheckKernels(slot, dense, filterOnly);
        } else if (mode == 3) {
            if (slot.direct) GSS_CUDA(cudaMemsetAsync(slot.ctrDev.data(), 0, (size_t)slot.nSolvers * kRecBuckets * kCtrStride * sizeof(unsigned long long), stream_));
            for (int g = 0; g < groups; g++)
                if (slot.aggStart[g]) launchExactOnly(checkArgs(slot, g, slot.direct), dims_, numSMs_, stream_, &launches_);
        } else if (mode == 6) {
            launchEmitFor(slot);
        } else if (mode == 7) {
            t2Sliced_.reserve((size_t)tables_.varCap * 32 * (size_t)groups, 0, stream_);
            launchSliceTables(tables_, 0, std::min(kMaxSolversPerGroup, slot.nSolvers), t2Sliced_.data(), numSMs_, stream_, &launches_);
        } else if (mode == 4) {
            launchApplyUpdates(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_,
                               stream_, &launches_);
        } else {
            launchCollapse(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_,
                           stream_, &launches_);
        }
    };
    cudaEvent_t e0, e1;
    GSS_CUDA(cudaEventCreate(&e0));
    GSS_CUDA(cudaEventCreate(&e1));
    if (mode == 6 && !(slot.direct && slot.checked)) return -1.0;
    if (mode == 1 && slot.direct) { // the dense kernel reports into the global hit buffer: this run finishes on the staged path
        slot.direct = false;
        if (lastDirect_ == &slot) lastDirect_ = nullptr;
    }
    if (mode == 3) launchCheckKernels(slot, false, true); // a fresh survivor list
    if (mode == 6) launchCheckKernels(slot, false, false); // fresh record lists
    once(); // warm-up
    GSS_CUDA(cudaEventRecord(e0, stream_));
    for (int i = 0; i < iters; i++) once();
    GSS_CUDA(cudaEventRecord(e1, stream_));
    if (mode == 5)
        launchApplyUpdates(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_,
                           stream_, &launches_);
    if (slot.direct) {
        launchDirectCheck(slot); // leave a complete result behind
    } else {
        if (mode >= 2) launchCheckKernels(slot, slot.dense, false);
        enqueueResultCopy(slot);
    }
    GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
    GSS_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    GSS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (mode <= 1) slot.dense = dense;
    return (double)ms * 1000.0 / iters;
}

int Sharer::lastRunTimes(double out[4]) {
    if (!haveTimes_) return 0;
    for (int i = 0; i < 4; i++) out[i] = lastTimes_[i];
    return 1;
}

double Sharer::lop3Peak() {
    useDevice();
    GSS_CUDA(cudaStreamSynchronize(stream_));
    return measureLop3Peak(numSMs_, stream_, &launches_);
}

} // namespace gss
