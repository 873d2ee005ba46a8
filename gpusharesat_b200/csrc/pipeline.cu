// pipeline.cu -- the direct run pipeline of a single-device sharer (the default path of gss_gpu_run).
//
// What the reference does around its kernel (GpuRunner.cu:274-383): copy every solver's deltas into a
// packed staging buffer, one H2D; after the run one D2H of the hit records; then, per hit, fetch the
// clause's literals from the host clause mirror into the solver's batch (Reported.cu:160-204).
// Here the host moves nothing:
//   * deltas: a solver thread writes its VarUpdate records into its own page-locked buffer
//     (assigs.h); collecting a run swaps that buffer out (O(1) under the solver's lock) and the
//     kernel that applies the deltas (k_apply_direct) reads it in place over PCIe -- the transfer is
//     the kernel's load stream, there is no staging copy and no separate H2D of the deltas;
//   * hits: k_exact appends (clause, mask) records to PER-SOLVER lists; k_emit sorts each list into
//     the reproducible hand-over order and writes the finished result of every solver -- clause ids,
//     literal positions, literal stream -- straight into a page-locked result buffer; the host
//     builds each solver's ClauseBatch as a VIEW over that buffer (no sort, no literal copies, no
//     D2H of unknown size) and the buffer returns to the pool when the last batch lets go of it.
// Overflow (survivor list, a solver's record list, the result buffer) is reported in the header and
// the check is launched again with larger buffers: the tables are intact (collapse is deferred), so
// nothing is ever dropped (reference: Reporter.cuh:46-48 drops).
#include "sharer.h"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>

namespace gss {

// ---- result buffers in page-locked host memory ----
class Sharer::RunBufPool : public std::enable_shared_from_this<Sharer::RunBufPool> {
public:
    ~RunBufPool() {
        for (RunBuf *b : free_) destroy(b);
    }
    // a buffer with room for at least that many entries / literals; returned to the pool by the last owner
    std::shared_ptr<RunBuf> acquire(int64_t entryCap, int64_t litCap) {
        RunBuf *b = nullptr;
        {
            std::lock_guard<std::mutex> g(m_);
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i]->entryCap >= entryCap && free_[i]->litCap >= litCap) {
                    b = free_[i];
                    free_[i] = free_.back();
                    free_.pop_back();
                    break;
                }
            if (!b && free_.size() >= 4) { // all too small for today's results: do not hoard them
                for (RunBuf *f : free_) destroy(f);
                free_.clear();
            }
            outstanding_++;
        }
        if (!b) {
            b = new RunBuf();
            b->entryCap = entryCap;
            b->litCap = litCap;
            b->bytes = RunBuf::bytesFor(entryCap, litCap);
            void *p = nullptr;
            // page-locked + mapped: with unified addressing the device writes through the same pointer
            if (cudaHostAlloc(&p, b->bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
                GSS_DIE("out of page-locked host memory for a result buffer of " + std::to_string(b->bytes) + " bytes");
            b->base = static_cast<uint8_t *>(p);
            memset(b->base, 0, RunBuf::hdrBytes());
        }
        std::shared_ptr<RunBufPool> self = shared_from_this();
        return std::shared_ptr<RunBuf>(b, [self](RunBuf *q) { self->release(q); });
    }
    int outstanding() {
        std::lock_guard<std::mutex> g(m_);
        return outstanding_;
    }

private:
    static void destroy(RunBuf *b) {
        if (b->base) cudaFreeHost(b->base);
        delete b;
    }
    void release(RunBuf *b) {
        std::lock_guard<std::mutex> g(m_);
        outstanding_--;
        free_.push_back(b);
    }
    std::mutex m_;
    std::vector<RunBuf *> free_;
    int outstanding_ = 0;
};

std::shared_ptr<Sharer::RunBufPool> makeRunBufPool();
std::shared_ptr<Sharer::RunBufPool> makeRunBufPool() { return std::make_shared<Sharer::RunBufPool>(); }

static size_t pow2AtLeast(size_t x) {
    size_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

void Sharer::ensureDirectBuffers(RunSlot &slot) {
    const size_t S = (size_t)std::max(1, slot.nSolvers);
    if (slot.ctrDev.capacity() < (size_t)kMaxSolvers * kRecBuckets * kCtrStride) {
        slot.ctrDev.reserve((size_t)kMaxSolvers * kRecBuckets * kCtrStride, 0, stream_);
        slot.solverInfo.reserve(kMaxSolvers, 0, stream_);
        slot.ticketDev.reserve(4, 0, stream_);
        slot.solverDone.reserve(kMaxSolvers, 0, stream_);
        GSS_CUDA(cudaMemsetAsync(slot.solverDone.data(), 0, (size_t)kMaxSolvers * sizeof(unsigned int), stream_));
        GSS_CUDA(cudaMemsetAsync(slot.ctrDev.data(), 0, (size_t)kMaxSolvers * kRecBuckets * kCtrStride * sizeof(unsigned long long), stream_));
        GSS_CUDA(cudaMemsetAsync(slot.solverInfo.data(), 0, (size_t)kMaxSolvers * sizeof(EmitSolver), stream_));
        GSS_CUDA(cudaMemsetAsync(slot.ticketDev.data(), 0, 4 * sizeof(unsigned int), stream_));
    }
    slot.recCap = (unsigned int)recCap_;
    HostProf hpR("    edb: reserves");
    slot.recKeys.reserve(S * recCap_, 0, stream_);
    slot.recMasks.reserve(S * recCap_, 0, stream_);
    slot.sortKeys.reserve(S * recCap_, 0, stream_);
    slot.sortMasks.reserve(S * recCap_, 0, stream_);
    slot.recPos.reserve(S * (recCap_ + 1), 0, stream_);
    slot.bucketBase.reserve(2 * (S * kRecBuckets + S), 0, stream_);
    survDev_.reserve((size_t)std::max(1, tables_.nGroups) * survCap_, 0, stream_);
    resDev_.reserve(sizeof(Counters) + hitCap_ * sizeof(HitRecord), 0, stream_);
}

// Collect: every solver's delta buffer is swapped out and read where it lies.  The one run that
// rebuilds the device tables lists every variable of every solver instead; that list is built by
// the GPU thread in the slot's (page-locked) staging buffer and read from there the same way.
// Fills slot.headHost = [directory][run parameters][per-solver delta pointers] and the slot's counts.
void Sharer::collectDirect(RunSlot &slot, bool rebuild) {
    const int S = slot.nSolvers;
    const int groups = (S + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    slot.srcOff = (slot.dirBytes + (size_t)S * sizeof(SolverRunParams) + 15) / 16 * 16;
    slot.headHost.resize(slot.srcOff + (size_t)S * sizeof(void *));
    slot.aggStart.assign(groups, 0u);
    slot.aggOnDevice = false;
    slot.maxUpd = 0;
    slot.staged.clear();
    slot.stagedOff.clear();
    PhaseTimer t(hostPhases_[3]);
    if (rebuild) {
        collectBatch(slot, true); // full update lists in slot.updHost, behind the payload prefix
        SolverRunParams *params = (SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
        const VarUpdate **src = (const VarUpdate **)(slot.headHost.data() + slot.srcOff);
        const VarUpdate *base = slot.updHost.data() + payloadPrefixRecords(S);
        for (int s = 0; s < S; s++) {
            src[s] = base + params[s].updStart;
            if (!slot.updHost.pinned() && params[s].updCount > 0) slot.staged.push_back({s, src[s]});
            slot.aggStart[s / kMaxSolversPerGroup] |= params[s].usedAggBits;
            slot.maxUpd = std::max(slot.maxUpd, (int)params[s].updCount);
        }
        return;
    }
    SolverRunParams *params = (SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
    const VarUpdate **src = (const VarUpdate **)(slot.headHost.data() + slot.srcOff);
    slot.ids.assign(S, AssigIds{});
    slot.assigCount = 0;
    slot.updHost.clear();
    int64_t total = 0;
    TimeAdder ta(globalStats_[G_timeSpentFillingAssigs], opts_.quickProf != 0);
    // one solver at a time, its lock held only for the O(1) swap (reference: locks, copies and unlocks
    // one solver after the other, Assigs.cu:346-352; a busy solver is skipped for this run)
    for (int s = 0; s < S; s++) {
        SolverAssigs &sa = assigs_->solver(s);
        memset(&params[s], 0, sizeof(SolverRunParams));
        params[s].updStart = (int32_t)total;
        src[s] = nullptr;
        if (!sa.tryLock()) continue;
        bool pinned = true;
        sa.takeUpdatesLocked(src[s], (int32_t)total, params[s], slot.ids[s], &pinned);
        sa.unlock();
        const int n = params[s].updCount;
        if (!pinned && n > 0) { // ordinary memory (page-locked budget exhausted): stage it like the reference does
            slot.stagedOff.push_back({s, slot.updHost.size()}); // (updHost may still move: pointers once it is complete)
            memcpy(slot.updHost.append((size_t)n), src[s], (size_t)n * sizeof(VarUpdate));
        }
        total += n;
        slot.assigCount += slot.ids[s].count;
        slot.aggStart[s / kMaxSolversPerGroup] |= params[s].usedAggBits;
        slot.maxUpd = std::max(slot.maxUpd, n);
    }
    slot.nUpdates = total;
    for (auto &so : slot.stagedOff) slot.staged.push_back({so.first, slot.updHost.data() + so.second});
    slot.stagedOff.clear();
}

// Launch the run described by slot.headHost on this device: header H2D, deferred collapse of the
// previous batch, k_apply_direct (the deltas cross PCIe inside it), check kernels, k_emit.
void Sharer::launchDirect(RunSlot &slot, int64_t h2d) {
    const int S = slot.nSolvers;
    SolverRunParams *params = (SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
    const VarUpdate **src = (const VarUpdate **)(slot.headHost.data() + slot.srcOff);
    HostProf hpAll("launchDirect");
    slot.direct = true;
    slot.dense = false;
    {
        HostProf hp("  updDev reserve");
        slot.updDev.reserve((size_t)std::max<int64_t>(slot.nUpdates, 1), 0, stream_);
    }
    for (auto &st : slot.staged) { // deltas that are not in page-locked memory go up as ordinary copies
        const int s = st.first;
        VarUpdate *dst = slot.updDev.data() + params[s].updStart;
        GSS_CUDA(cudaMemcpyAsync(dst, st.second, (size_t)params[s].updCount * sizeof(VarUpdate), cudaMemcpyHostToDevice, stream_));
        src[s] = dst;
    }
    {
        HostProf hp("  headDev reserve");
        slot.headDev.reserve(slot.headHost.size(), 0, stream_);
    }
    // The run header travels inside the first kernel (k_apply_direct reads it from page-locked host memory and
    // leaves the device copy behind) and the same kernel zeroes the run's counters; only a header that is not
    // page-locked goes up as an ordinary copy.
    const bool fused = fuseHeader_ && slot.headHost.pinned();
    if (!fused) {
        HostProf hp("  head H2D");
        GSS_CUDA(cudaMemcpyAsync(slot.headDev.data(), slot.headHost.data(), slot.headHost.size(), cudaMemcpyHostToDevice, stream_));
    }
    h2d += (int64_t)slot.headHost.size() + slot.nUpdates * (int64_t)sizeof(VarUpdate);
    {
        HostProf hp("  ensureDirectBuffers");
        ensureDirectBuffers(slot);
    }
    {
        HostProf hp("  event record");
        GSS_CUDA(cudaEventRecord(slot.evH2DDone, stream_));
    }

    // the previous batch collapses to its last slot first (deferred dSetAllAssigsToLast)
    if (collapseSlot_ >= 0) {
        HostProf hp("  launch collapse");
        RunSlot &c = slots_[collapseSlot_];
        launchCollapse(c.updDev.data(), c.paramsDev(), c.nSolvers, c.maxUpd, c.nUpdates, tables_, numSMs_, stream_, &launches_);
        collapseSlot_ = -1;
    }
    // A run with nothing new -- no delta, no frozen assignment (a GPU thread that calls gpuRun() faster than the
    // solvers export) -- has nothing to apply and nothing to check: only the deferred collapse above was due.
    bool anyFrozen = false;
    for (uint32_t g : slot.aggStart) anyFrozen = anyFrozen || g != 0;
    const bool idleRun = slot.nUpdates == 0 && !anyFrozen && slot.staged.empty();
    if (!idleRun) {
        HostProf hp("  launch apply");
        ApplyExtra x;
        const uint8_t *head = slot.headDev.data();
        if (fused) {
            head = slot.headHost.data(); // (this kernel's own blocks read their parameters where the host wrote them)
            x.headHost = reinterpret_cast<const uint32_t *>(slot.headHost.data());
            x.headDev = reinterpret_cast<uint32_t *>(slot.headDev.data());
            x.headWords = (int)((slot.headHost.size() + 3) / 4);
            x.zeroA = reinterpret_cast<uint32_t *>(resDev_.data());
            x.zeroAWords = (int)(sizeof(Counters) / 4);
            x.zeroB = reinterpret_cast<uint32_t *>(slot.ctrDev.data());
            x.zeroBWords = (int)((size_t)S * kRecBuckets * kCtrStride * sizeof(unsigned long long) / 4);
            slot.countersZeroed = true;
        }
        launchApplyDirect((const VarUpdate *const *)(head + slot.srcOff), (const SolverRunParams *)(head + slot.dirBytes), S,
                          slot.maxUpd, tables_, slot.updDev.data(), numSMs_, stream_, &launches_, x);
    }
    GSS_CUDA(cudaEventRecord(slot.evBeforeCheck, stream_));
    launchDirectCheck(slot);
    slot.countersZeroed = false; // (only the check launched right behind the apply may rely on it)
    GSS_CUDA(cudaEventRecord(slot.evAfterCheck, stream_));
    GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
    slot.inFlight = true;
    if (slot.nUpdates) collapseSlot_ = (int)(&slot - slots_);
    lastStarted_ = (int)(&slot - slots_);
    lastH2D_ = h2d;
}

bool Sharer::startRunDirect(RunSlot &slot) {
    int64_t h2d = 0;
    bool rebuild = false;
    {
        HostProf hp("prepareRun");
        if (!prepareRun(slot, rebuild, h2d)) return false;
    }
    {
        HostProf hp("collectDirect");
        collectDirect(slot, rebuild);
    }
    launchDirect(slot, h2d);
    return true;
}

// k_filter + k_exact (per-solver record lists) for every solver group with a frozen slot, then k_emit
void Sharer::launchDirectCheck(RunSlot &slot) {
    const int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    bool any = false;
    for (int g = 0; g < groups; g++) any = any || slot.aggStart[g] != 0;
    slot.checked = any && slot.totalTiles > 0;
    if (!slot.checked) return;
    HostProf hpAll("  launchDirectCheck");
    ensureDirectBuffers(slot);
    if (lastDirect_ == &slot) lastDirect_ = nullptr; // its record lists are about to be overwritten
    slot.seq = ++directSeq_;
    // capacities in powers of two: a released buffer fits the next run although the guesses drift
    if (peer_ && peerAcquireResultBuf(slot)) {
        // (a worker rank of the multi-process exchange: a buffer of its shared-memory ring, which rank 0 reads)
    } else {
        // ... and sticky: a need that hovers around a power of two must not alternate between two buffer sizes
        // (a miss is a page-locked allocation: tens of milliseconds); shrink only below a quarter
        int64_t wantE = (int64_t)pow2AtLeast((size_t)entryGuess_), wantL = (int64_t)pow2AtLeast((size_t)litGuess_);
        if (wantE < runCapE_ && wantE * 4 > runCapE_) wantE = runCapE_;
        if (wantL < runCapL_ && wantL * 4 > runCapL_) wantL = runCapL_;
        runCapE_ = wantE;
        runCapL_ = wantL;
        HostProf hp("    acquire result buffer");
        slot.runBuf = runBufs_->acquire(wantE, wantL);
    }
    {
        HostProf hp("    launch filter+exact");
        launchCheckKernels(slot, false);
    }
    HostProf hp("    launch emit");
    launchEmitFor(slot);
}

void Sharer::launchEmitFor(RunSlot &slot) {
    const int S = slot.nSolvers;
    const int groups = (S + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    EmitArgs e;
    e.dir = slot.dirDev();
    e.nDir = slot.nDir;
    e.nSolvers = S;
    e.solverCtr = slot.ctrDev.data();
    e.recKeys = slot.recKeys.data();
    e.recMasks = slot.recMasks.data();
    e.sortKeys = slot.sortKeys.data();
    e.sortMasks = slot.sortMasks.data();
    e.recPos = slot.recPos.data();
    e.bucketBase = slot.bucketBase.data();
    e.solverInfo = slot.solverInfo.data();
    e.recCap = slot.recCap;
    e.counters = (Counters *)resDev_.data();
    e.survCap = (unsigned int)survCap_;
    e.groups = groups;
    e.ticket = slot.ticketDev.data();
    e.solverDone = slot.solverDone.data();
    e.seq = slot.seq;
    RunBuf &rb = *slot.runBuf;
    e.hdr = rb.hdr();
    e.ids = rb.ids();
    e.pos = rb.pos();
    e.lits = rb.lits();
    e.entryCap = rb.entryCap;
    e.litCap = rb.litCap;
    if (rb.withRecords) {
        e.keysOut = rb.keys();
        e.masksOut = rb.masks();
    }
    launchEmit(e, stream_, &launches_);
}

void Sharer::finishRunDirect(RunSlot &slot) {
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
    slot.inFlight = false;
    postValid_ = false;
    hits_.clear();
    finishedD2H_ = 0;
    if (!slot.checked) {
        slot.runBuf.reset();
        return;
    }
    for (int attempt = 0;; attempt++) {
        const RunHdr *h = slot.runBuf->hdr();
        if (*reinterpret_cast<const volatile uint32_t *>(&h->seq) != slot.seq) GSS_DIE("result header of a finished run is missing");
        const uint32_t flags = h->flags;
        if (flags == 0) break;
        GSS_CHECK(attempt < 10);
        finishReran_ = true;
        // Overflow: the tables of this run are still intact (collapse is deferred): grow what was too small
        // and launch the check again.  Nothing is dropped.
        if (flags & 1u) {
            size_t maxSurv = 0;
            for (int g = 0; g < kMaxGroups; g++) maxSurv = std::max(maxSurv, (size_t)h->nSurvivors[g]);
            survCap_ = std::max(survCap_ * 2, maxSurv + maxSurv / 4);
        }
        if (flags & 2u) // (maxRec = the fullest bucket of any solver)
            recCap_ = pow2AtLeast(std::max<size_t>(recCap_ * 2, ((size_t)h->maxRec + h->maxRec / 2) * kRecBuckets));
        if (flags & 4u) {
            entryGuess_ = std::max<int64_t>(entryGuess_ * 2, h->nTotal + h->nTotal / 2);
            litGuess_ = std::max<int64_t>(litGuess_ * 2, h->litTotal + h->litTotal / 2);
        }
        slot.runBuf.reset();
        launchDirectCheck(slot);
        GSS_CUDA(cudaStreamSynchronize(stream_));
    }
    const RunHdr *h = slot.runBuf->hdr();
    globalStats_[G_clauseTestsOnAssigs] += h->exactTests;
    // Head room for the batches to come: an overflow is repaired by a second pass over the same batch plus larger
    // buffers (tens of milliseconds of cudaMalloc / cudaFree / synchronisation) -- with capacities that follow the
    // observed counts closely, one batch in a few hundred overflowed by fluctuation alone (a 60 ms hiccup in the middle
    // of a timed region, profiles/r02o_bench.json).  Grow ahead, while the counts are still far from the capacities.
    {
        size_t maxSurv = 0;
        for (int g = 0; g < kMaxGroups; g++) maxSurv = std::max(maxSurv, (size_t)h->nSurvivors[g]);
        if (maxSurv * 2 > survCap_) survCap_ = maxSurv * 3;
        if ((size_t)h->maxRec * 3 > recCap_ / kRecBuckets) recCap_ = pow2AtLeast((size_t)h->maxRec * 4 * kRecBuckets);
    }
    // size the next result buffer from this result (with head room), the hit buffer guess follows the trend
    entryGuess_ = std::max<int64_t>(4096, std::max(h->nTotal + h->nTotal / 2, entryGuess_ - entryGuess_ / 16));
    litGuess_ = std::max<int64_t>(16384, std::max(h->litTotal + h->litTotal / 2, litGuess_ - litGuess_ / 16));
    finishedD2H_ = (int64_t)(sizeof(RunHdr) + (size_t)h->nTotal * 12 + (size_t)slot.nSolvers * 4 + (size_t)h->litTotal * 4);
}

bool Sharer::waitBumpFlag() {
    if (!bumpFlagPending_) return false;
    // wait for that bump's flag only -- not for whatever has been queued since (the next run)
    GSS_CUDA(cudaEventSynchronize(bumpFlagEv_));
    bumpFlagPending_ = false;
    return bumpFlagHost_[0] != 0;
}

// reference: one host-side bump per hit record (Clauses.cu:231-237, GpuRunner.cu:375-378); here one
// kernel per device over its sorted per-solver record lists, with the increment the reference would
// use at this point (after the next batch of clauses has been drained).  The activities live on THIS
// device; the record lists of other devices of the process are read in place through peer access.
void Sharer::bumpDirect(const std::vector<DevicePart> &parts) {
    HostProf hpAll("bumpDirect");
    if (waitBumpFlag()) {
        db_->rescaleAfterDeviceOverflow();
        db_->applyPendingDeviceRescales(stream_);
    }
    bumpOwners_.clear(); // (the bump that read them has completed: waitBumpFlag)
    bool any = false;
    for (const DevicePart &p : parts) any = any || (p.view() && p.view()->hdr()->nTotal > 0);
    if (!any) return;
    std::vector<LenDir> dir;
    db_->buildDirectory(dir);
    bumpDirHost_.resize(dir.size() * sizeof(LenDir));
    memcpy(bumpDirHost_.data(), dir.data(), dir.size() * sizeof(LenDir));
    bumpDirDev_.reserve(bumpDirHost_.size(), 0, stream_);
    GSS_CUDA(cudaMemcpyAsync(bumpDirDev_.data(), bumpDirHost_.data(), bumpDirHost_.size(), cudaMemcpyHostToDevice, stream_));
    bumpFlagDev_.reserve(1, 0, stream_);
    bumpFlagHost_.resize(1);
    GSS_CUDA(cudaMemsetAsync(bumpFlagDev_.data(), 0, sizeof(int), stream_));
    for (const DevicePart &p : parts) {
        const RunBuf *rb = p.view();
        if (!rb || rb->hdr()->nTotal == 0) continue;
        if (p.slot) {
            RunSlot &slot = *p.slot;
            launchBumpFromRecs(slot.sortKeys.data(), slot.recCap, slot.solverInfo.data(), slot.nSolvers, slot.recCap,
                               (const LenDir *)bumpDirDev_.data(), (int)dir.size(), db_->activityIncrement(), bumpFlagDev_.data(),
                               stream_, &launches_);
        } else if (rb->withRecords) { // another process's result: its sorted record keys lie next to the ids, in host memory this device can read
            launchBumpFromKeys(rb->keys(), rb->hdr()->nTotal, (const LenDir *)bumpDirDev_.data(), (int)dir.size(),
                               db_->activityIncrement(), bumpFlagDev_.data(), stream_, &launches_);
            bumpOwners_.push_back(p.buf);
        }
    }
    GSS_CUDA(cudaMemcpyAsync(bumpFlagHost_.data(), bumpFlagDev_.data(), sizeof(int), cudaMemcpyDeviceToHost, stream_));
    if (!bumpFlagEv_) GSS_CUDA(cudaEventCreateWithFlags(&bumpFlagEv_, cudaEventDisableTiming));
    GSS_CUDA(cudaEventRecord(bumpFlagEv_, stream_)); // (also: the other devices' record lists have been read)
    bumpFlagPending_ = true;
}

void Sharer::processResultsDirect(RunSlot &slot) {
    std::vector<DevicePart> parts{DevicePart{this, &slot, nullptr}};
    processResultsParts(slot, parts);
}

// parts[0] is this sharer's own slot; with several devices in one process (multi.cu) every device
// contributes one slice per solver
void Sharer::processResultsParts(RunSlot &slot, const std::vector<DevicePart> &parts) {
    // reference gatherGpuRunResults, GpuRunner.cu:360-383 (64-bit arithmetic)
    const int64_t clCount = db_->stats().clauses;
    int64_t nTotal = 0;
    lastForeign_.clear();
    for (const DevicePart &p : parts) {
        if (p.view()) nTotal += p.view()->hdr()->nTotal;
        if (!p.slot && p.buf) lastForeign_.push_back(p.buf);
    }
    globalStats_[G_gpuRuns]++;
    globalStats_[G_totalAssigClauseTested] += (uint64_t)clCount * (uint64_t)slot.assigCount;
    globalStats_[G_clauseTestsOnGroups] += (uint64_t)clCount;
    globalStats_[G_gpuReports] += (uint64_t)nTotal;
    lastHitsValid_ = false;
    lastDirect_ = &slot;
    bumpN_ = 0; // (nothing parked by the staged path)
    bumpDirect(parts);
    TimeAdder t(globalStats_[G_timeSpentFillingReported], opts_.quickProf != 0);
    HostProf hpViews("handOverViews");
    std::vector<std::vector<ResultView>> views((size_t)slot.nSolvers);
    for (const DevicePart &p : parts) {
        if (!p.view()) continue;
        const RunHdr *h = p.view()->hdr();
        if (h->nTotal <= 0) continue;
        const RunBuf &rb = *p.view();
        // safety valve: a solver that does not pop keeps its batches, and with them whole result buffers,
        // alive; past a bound the slices are copied out and the buffer goes back to its pool at once
        // (another process's buffers come from a small ring: their slices are always copied out late)
        const bool copyOut = p.sh ? p.sh->runBufs_->outstanding() > 64 : foreignCopyOut_;
        for (int s = 0; s < slot.nSolvers; s++) {
            const RunHdr::PerSolver &ps = h->solver[s];
            if (ps.n <= 0) continue;
            ResultView v;
            v.n = ps.n;
            if (!copyOut) {
                v.ids = rb.ids() + ps.entryBase;
                v.pos = rb.pos() + ps.entryBase + s;
                v.lits = rb.lits() + ps.litBase;
                v.owner = p.owner();
            } else {
                const size_t bytes = (size_t)ps.n * 8 + ((size_t)ps.n + 1) * 4 + (size_t)ps.nLits * 4;
                std::shared_ptr<uint8_t> heap(new uint8_t[bytes + 8], std::default_delete<uint8_t[]>());
                int64_t *ids = reinterpret_cast<int64_t *>(heap.get());
                int32_t *pos = reinterpret_cast<int32_t *>(ids + ps.n);
                int32_t *lits = pos + ps.n + 1;
                memcpy(ids, rb.ids() + ps.entryBase, (size_t)ps.n * 8);
                memcpy(pos, rb.pos() + ps.entryBase + s, ((size_t)ps.n + 1) * 4);
                memcpy(lits, rb.lits() + ps.litBase, (size_t)ps.nLits * 4);
                v.ids = ids;
                v.pos = pos;
                v.lits = lits;
                v.owner = heap;
            }
            views[s].push_back(std::move(v));
        }
    }
    reported_->handOverViews(views, slot.ids, slot.nSolvers);
}

// (clause id, solver, mask) triples of a finished direct run on THIS device, appended to `out`
void Sharer::appendDirectHits(RunSlot &slot, std::vector<gss_hit> &out) {
    if (!slot.checked || !slot.runBuf) return;
    useDevice();
    const RunHdr *h = slot.runBuf->hdr();
    std::vector<uint32_t> masks;
    for (int s = 0; s < slot.nSolvers; s++) {
        const RunHdr::PerSolver &ps = h->solver[s];
        if (ps.n <= 0) continue;
        masks.resize((size_t)ps.n);
        GSS_CUDA(cudaMemcpyAsync(masks.data(), slot.sortMasks.data() + (size_t)s * slot.recCap, (size_t)ps.n * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, stream_));
        GSS_CUDA(cudaStreamSynchronize(stream_));
        const int64_t *ids = slot.runBuf->ids() + ps.entryBase;
        for (int32_t i = 0; i < ps.n; i++) out.push_back(gss_hit{ids[i], s, masks[(size_t)i]});
    }
}

} // namespace gss
