// sharer.h -- the engine behind the C ABI: owns the clause database, the assignment state
// machines, the device tables and the run pipeline (replaces the reference's
// GpuClauseSharerImpl + GpuRunner, gpuShareLib/GpuClauseSharerImpl.cu, GpuRunner.cu).
#pragma once
#include "../../include/gpushare_b200.h"
#include "assigs.h"
#include "clause_db.h"
#include "kernels.cuh"
#include "reported.h"
#include "stats.h"
#include <chrono>
#include <memory>
#include <mutex>
#include <vector>

namespace gss {

inline int64_t nowMicros() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct PhaseTimer { // host-side wall time of a phase of gss_gpu_run (gss_debug_host_phases)
    double &acc;
    int64_t t0;
    explicit PhaseTimer(double &a) : acc(a), t0(nowMicros()) {}
    ~PhaseTimer() { acc += (double)(nowMicros() - t0); }
};
struct TimeAdder { // reference TimeGauge, gpuShareLib/Profiler.h:28-46
    uint64_t &acc;
    bool on;
    int64_t t0;
    TimeAdder(uint64_t &a, bool enabled) : acc(a), on(enabled), t0(enabled ? nowMicros() : 0) {}
    ~TimeAdder() { if (on) acc += (uint64_t)(nowMicros() - t0); }
};

class Sharer {
public:
    // workerOfDevice >= 0: a device worker of another sharer of this process (multi.cu), on that device
    Sharer(const gss_options &opts, gss_log_fn log, void *logCtx, int workerOfDevice = -1);
    ~Sharer();

    // ---- GpuClauseSharer.h API (see include/gpushare_b200.h for the line map) ----
    void gpuRun();
    void reduceDb();
    int64_t addedClauseCount() { return db_->stats().added; }
    int64_t addedClauseCountAtLastReduceDb() { return db_->addedAtLastReduceDb(); }
    bool hasRunOutOfGpuMemoryOnce() const { return ranOutOfMemory_; }
    void gpuMemInfo(size_t *freeB, size_t *totalB);
    int64_t globalStat(int stat);
    void writeClausesInCnf(FILE *f) { db_->writeCnf(f, varCount_); }
    void setVarCount(int n);
    void setCpuSolverCount(int n);
    int64_t addClause(int solver, const int *lits, int n);
    bool trySetSolverValues(int solver, const int *lits, int n);
    void unsetSolverValues(int solver, const int *lits, int n);
    int64_t trySendAssignment(int solver);
    bool popReportedClause(int solver, int *&lits, int &count, int64_t &id);
    int64_t lastAssigAllReported(int solver) { return reported_->lastAssigAllReported(solver); }
    void currentAssignment(int solver, uint8_t *assig);
    int64_t oneSolverStat(int solver, int stat) { return (int64_t)oneSolverStats_[solver][stat]; }

    // ---- parity / bench hooks ----
    int64_t lastHits(gss_hit *out, int64_t cap);
    int64_t addClausesBulk(const int64_t *offsets, const int *lits, int64_t n);
    void setMaxClauseLen(int n) {
        db_->setMaxLen(n);
        for (auto &w : workers_) w->db_->setMaxLen(n);
    }
    void setDense(bool d) { dense_ = d; }
    double timeCheck(int iters, int mode);
    int lastRunTimes(double out[4]);
    double lop3Peak();
    void lastRunBytes(int64_t *h2d, int64_t *d2h) { *h2d = lastH2D_; *d2h = finishedD2H_; }
    int64_t kernelLaunches() const { return launches_; }
    // cumulative host wall time (us) of the phases of gss_gpu_run: [0] finish the previous run (wait +
    // device-side sort/resolve + D2H), [1] start the next run (drain, upload, collect, enqueue),
    // [2] hand-over, [3] collect alone (part of 1), [4] wait for the GPU alone (part of 0),
    // [5] sort/resolve/D2H of a large hit list (part of 0)
    void hostPhases(double out[6]) const {
        for (int i = 0; i < 6; i++) out[i] = hostPhases_[i];
        if (HostProf::enabled()) {
            HostProf::dump();
            fprintf(stderr, "host prof ---- (reset)\n");
            HostProf::reset();
        }
    }
    // ---- multi-GPU (see include/gpushare_b200.h) ----
    void setShard(int rank, int world) { db_->setShard(rank, world); }
    int mgpuCollect(const void **params, int64_t *paramsBytes, const void **updates, int64_t *nUpdates);
    void mgpuRun(const void *params, int64_t paramsBytes, const void *updates, int64_t nUpdates, int rebuild);
    int64_t mgpuCollectTo(void *devDst, int64_t capBytes);
    int mgpuRunPayload(const void *devPayload, int64_t payloadBytes);
    int mgpuEnqueuePayload(const void *devPayload, int64_t validBytes);
    void mgpuRedoPayload(const void *devPayload, int64_t totalBytes);
    int64_t mgpuEnqueueResult(void *devDst, int64_t capRecords);
    int mgpuFinish();
    int64_t mgpuHitsToDevice(void *devDst, int64_t capRecords);
    void setStream(void *stream);
    int64_t mgpuWait(const HitRecord **hits);
    void mgpuImport(const HitRecord *hits, int64_t n);
    void mgpuImportGathered(const void *devGathered, int world, int64_t slotBytes, const int64_t *counts);
    void dbSize(int64_t *ncl, int64_t *nlits) { *ncl = db_->stats().clauses; *nlits = db_->stats().lengthSum; }
    void dbOrder(int64_t *unsorted, int64_t *resorts) { *unsorted = db_->unsortedClauses(); *resorts = resorts_; }
    // ---- multi-GPU over peer memory (CUDA IPC windows, no collective on the data path: peer.cu) ----
    int64_t peerInit(int rank, int world, int64_t payloadCap, int64_t slotHits, void *blobOut, int64_t blobCap);
    void peerConnect(const void *blobs, int64_t blobBytes);
    int peerEnqueue();
    int64_t peerFinish();
    void setPeerRecords(bool on);
    void peerClose();

    struct RunBuf { // one run's result buffer in page-locked host memory, written by k_emit (pipeline.cu)
        uint8_t *base = nullptr;
        size_t bytes = 0;
        int64_t entryCap = 0, litCap = 0;
        RunHdr *hdr() const { return reinterpret_cast<RunHdr *>(base); }
        int64_t *ids() const { return reinterpret_cast<int64_t *>(base + hdrBytes()); }
        int32_t *pos() const { return reinterpret_cast<int32_t *>(base + hdrBytes() + (size_t)entryCap * 8); }
        int32_t *lits() const { return pos() + (size_t)entryCap + kMaxSolvers; }
        // buffers another process reads (multi-process exchange, peer.cu) also carry the sorted records
        bool withRecords = false;
        unsigned long long *keys() const { return reinterpret_cast<unsigned long long *>(lits() + (size_t)litCap); }
        uint32_t *masks() const { return reinterpret_cast<uint32_t *>(keys() + (size_t)entryCap); }
        static size_t hdrBytes() { return (sizeof(RunHdr) + 255) / 256 * 256; }
        static size_t bytesFor(int64_t entryCap, int64_t litCap, bool withRecords = false) {
            return hdrBytes() + (size_t)entryCap * 8 + ((size_t)entryCap + kMaxSolvers) * 4 + (size_t)litCap * 4 +
                   (withRecords ? (size_t)entryCap * 12 : 0);
        }
    };
    class RunBufPool;

private:
    struct RunSlot {
        HostBuf<uint8_t> headHost; // [LenDir x nDir][SolverRunParams x nSolvers]
        HostBuf<VarUpdate> updHost;
        DevBuf<uint8_t> headDev;
        DevBuf<VarUpdate> updDev;
        HostBuf<uint8_t> resHost;  // [Counters][HitRecord x chunk]
        std::vector<AssigIds> ids;
        std::vector<uint32_t> aggStart; // per solver group
        int nDir = 0, totalTiles = 0, nSolvers = 0, maxUpd = 0, assigCount = 0;
        int64_t nUpdates = 0;
        bool dense = false;
        bool inFlight = false;
        bool aggOnDevice = false; // multi-GPU receiver: the host never saw the run parameters
        cudaEvent_t evStart = nullptr, evH2DDone = nullptr, evBeforeCheck = nullptr, evAfterCheck = nullptr, evEnd = nullptr;
        // ---- direct pipeline (pipeline.cu) ----
        bool direct = false;        // this run used the direct pipeline (per-solver records, k_emit into host memory)
        bool checked = false;       // check kernels + k_emit were launched (some solver had a frozen slot)
        bool countersZeroed = false; // this run's k_apply_direct has reset the counters: no memsets before the check
        uint32_t seq = 0;           // what k_emit writes into the header last
        std::shared_ptr<RunBuf> runBuf;
        DevBuf<unsigned long long> ctrDev;  // [kMaxSolvers][kRecBuckets] x kCtrStride: records | literals << 32
        DevBuf<EmitSolver> solverInfo;      // [kMaxSolvers] written by k_emit_sort
        DevBuf<unsigned long long> recKeys; // [nSolvers][kRecBuckets][recCap / kRecBuckets] as k_exact appends them
        DevBuf<uint32_t> recMasks;
        DevBuf<unsigned long long> sortKeys; // [nSolvers][recCap] every solver's records in canonical order (k_emit_sort)
        DevBuf<uint32_t> sortMasks;
        DevBuf<int32_t> recPos;             // [nSolvers][recCap + 1]
        DevBuf<unsigned int> solverDone;    // [kMaxSolvers] (k_emit_fused)
        DevBuf<long long> bucketBase;       // [nSolvers * kRecBuckets + nSolvers][2] (k_emit_scan)
        DevBuf<unsigned int> ticketDev;     // 4 words
        unsigned int recCap = 0;
        size_t srcOff = 0;          // offset of the per-solver delta pointers in headHost / headDev
        std::vector<std::pair<int, const VarUpdate *>> staged; // solvers whose deltas are not page-locked: {solver, host copy}
        std::vector<std::pair<int, size_t>> stagedOff;
        const LenDir *dirDev() const { return (const LenDir *)headDev.data(); }
        const SolverRunParams *paramsDev() const { return (const SolverRunParams *)(headDev.data() + dirBytes); }
        size_t dirBytes = 0;
    };

    // ---- direct pipeline (pipeline.cu): deltas are read by the GPU from the solver threads' own
    // page-locked buffers, results are written by the GPU into page-locked result buffers that the
    // solvers' ClauseBatches view in place ----
    bool startRunDirect(RunSlot &slot);
    void collectDirect(RunSlot &slot, bool rebuild);
    void launchDirect(RunSlot &slot, int64_t h2d);
    void launchDirectCheck(RunSlot &slot);
    void launchEmitFor(RunSlot &slot);
    void finishRunDirect(RunSlot &slot);
    void processResultsDirect(RunSlot &slot);
    // One device's share of a finished run: a result buffer in host memory (this process's or, in the
    // multi-process exchange, another rank's shared-memory buffer) and where the sorted records are for
    // the activity bumps / the parity hook: on a device of this process (slot) or next to the ids (buf).
    struct DevicePart {
        Sharer *sh = nullptr;        // the sharer whose pool the buffer came from (nullptr: another process)
        RunSlot *slot = nullptr;     // device-resident record lists (nullptr: use buf->keys() / masks())
        std::shared_ptr<RunBuf> buf; // nullptr: the slot's own (slot->runBuf), if it checked anything
        const RunBuf *view() const { return buf ? buf.get() : (slot && slot->checked ? slot->runBuf.get() : nullptr); }
        std::shared_ptr<RunBuf> owner() const { return buf ? buf : slot->runBuf; }
    };
    void processResultsParts(RunSlot &slot, const std::vector<DevicePart> &parts);
    void appendDirectHits(RunSlot &slot, std::vector<gss_hit> &out);
    bool foreignCopyOut_ = false; // multi-process exchange: copy the slices of other ranks' results out of their ring buffers
    std::vector<std::shared_ptr<RunBuf>> bumpOwners_;    // foreign result buffers a queued bump kernel still reads
    std::vector<std::shared_ptr<RunBuf>> lastForeign_;   // foreign parts of the last processed run (parity hook)
    void ensureDirectBuffers(RunSlot &slot);
    void bumpDirect(const std::vector<DevicePart> &parts);
    bool waitBumpFlag();
    // ---- several devices in one process (GPUSHARE_DEVICES=N, multi.cu): this sharer is the front-end
    // and device 0; workers_[r-1] drives device r.  Every device keeps the whole clause database and
    // checks its share of the tiles; the batch reaches every device straight from the solver threads'
    // page-locked buffers (every GPU reads them over its own PCIe link), every device writes its
    // finished per-solver results into page-locked result buffers, the front-end stitches views ----
    class WorkerThreads;
    void multiInit(int nDevices);
    void multiShutdown();
    void wholeRunMulti(bool canStart);
    void reduceDbMulti();
    void workerStart(const RunSlot &rootSlot, const uint8_t *paramsAndSrc, int slotIdx, bool rebuild);
    std::vector<uint8_t> multiSnap_;
    void workerFinish(int slotIdx);
    std::vector<std::unique_ptr<Sharer>> workers_;
    WorkerThreads *wthreads_ = nullptr; // (owned; created / deleted in multi.cu)
    std::mutex multiAddLock_;
    bool isWorker_ = false;
    cudaEvent_t peerReadEv_ = nullptr; // front-end: its bump kernels have read the workers' record lists
    bool peerReadRecorded_ = false;
    Sharer *root_ = nullptr;

    std::shared_ptr<RunBufPool> runBufs_;
    DevBuf<uint2> t2Sliced_;          // dense mode (bench): the level-2 table cut into L2-sized slices of 8 solvers
    uint32_t denseSlicesValid_ = 0;   // per solver group: the slices match the current tables
    bool denseSliced_ = true;
    bool fuseHeader_ = true; // direct pipeline: the run header and the counter reset ride in k_apply_direct
    int64_t resorts_ = 0; // device-side re-sorts of streamed clauses (ClauseDb::resortOnDevice)
    bool directEnabled_ = true;
    bool eagerResults_ = true; // surface a run's hits in the call that started it when it completes within minGpuLatencyMicros
    size_t recCap_ = 4096;       // per-solver record capacity (power of two, >= kRecBuckets)
    int64_t entryGuess_ = 4096, litGuess_ = 16384; // result buffer sizing (from previous runs)
    int64_t runCapE_ = 0, runCapL_ = 0;            // capacities of the previous run's result buffer
    uint32_t directSeq_ = 0;
    RunSlot *lastDirect_ = nullptr; // finished direct run whose records are still on the device (parity hook)

    void useDevice();
    void wholeRun(bool canStart);
    bool startRun(RunSlot &slot);      // false: nothing started
    bool prepareRun(RunSlot &slot, bool &rebuild, int64_t &h2d);
    void collectBatch(RunSlot &slot, bool rebuild);
    void launchRun(RunSlot &slot, const void *updSrc, int64_t nUpdates, int64_t &h2d);
    int nextSlot() const;
    void enqueueFromDevicePayload(RunSlot &slot, const void *devPayload, int64_t validBytes, bool collapsePrev, int64_t h2d);
    void finishRun(RunSlot &slot, bool fetchAllHits = true, bool allowPostprocess = true);     // wait, re-run on overflow, pull every hit to the host
    void processResults(RunSlot &slot);
    // returns true when the last k_exact launch also published the peer result (fusedPublish_ set)
    bool launchCheckKernels(RunSlot &slot, bool dense, bool filterOnly = false);
    void enqueueResultCopy(RunSlot &slot);
    void materializeLastHits();
    bool ensureTables(bool &rebuild);
    void ensureResultBuffers();
    void unsetPendingLocked(int solver);
    CheckArgs checkArgs(const RunSlot &slot, int group, bool recs = false) const;
    size_t resultChunk() const { return hitCap_ < 4096 ? hitCap_ : 4096; }

    gss_options opts_;
    Logger logger_;
    int device_ = 0, numSMs_ = 1;
    cudaStream_t stream_ = nullptr;
    LaunchDims dims_;

    std::unique_ptr<ClauseDb> db_;
    std::unique_ptr<HostAssigs> assigs_;
    std::unique_ptr<Reported> reported_;
    std::vector<std::vector<uint64_t>> oneSolverStats_;
    std::vector<std::vector<int>> toUnset_;
    uint64_t globalStats_[G_COUNT] = {};
    int varCount_ = 0;

    // device state
    DeviceTables tables_;
    DevBuf<uint2> a1_, t2_;
    bool tablesValid_ = false;
    int tableSolvers_ = 0;
    DevBuf<uint8_t> resDev_; // [Counters][HitRecord x hitCap]
    DevBuf<Survivor> survDev_; // [nGroups][survCap]
    size_t hitCap_ = 0, survCap_ = 0;
    int peerRecordsWanted_ = -1; // -1: GPUSHARE_PEER_RECORDS / the default decides

    RunSlot slots_[2];
    int cur_ = -1;          // slot of the run in flight
    int collapseSlot_ = -1; // slot whose updates still have to be collapsed on the device
    int lastStarted_ = -1;  // slot whose tables are still intact (for timeCheck)
    int mgpuPending_ = -1;  // multi-GPU: slot collected but not yet launched
    int mgpuLast_ = -1;     // multi-GPU: slot of the last finished run
    int64_t mgpuH2D_ = 0;
    bool mgpuRebuild_ = false;
    bool finishReran_ = false;
    bool ownStream_ = true;
    HostBuf<uint8_t> mgpuHdrHost_;

    // multi-GPU payload = [PayloadHeader][SolverRunParams x nSolvers][VarUpdate x n], the first two
    // padded to a whole number of VarUpdate records
    struct PayloadHeader {
        uint32_t magic;
        int32_t status; // -1 nothing to run, 0 batch, 1 batch that rebuilds the tables
        int32_t nSolvers;
        int32_t prefixRecords;
        int64_t nUpdates;
        int64_t totalBytes;
        int64_t pad[4];
    };
    static constexpr uint32_t kPayloadMagic = 0x47535331u; // "GSS1"
    static size_t payloadPrefixRecords(int nSolvers) {
        return (sizeof(PayloadHeader) + (size_t)nSolvers * sizeof(SolverRunParams) + sizeof(VarUpdate) - 1) / sizeof(VarUpdate);
    }

    // ---- peer-memory exchange state (peer.cu) ----
    struct PeerState;
    PeerState *peer_ = nullptr;
    int64_t peerPayloadBytes_ = 0;
    HitRecord *hitsOverride_ = nullptr; // the check kernels append their hits here (peer mode: this rank's
    unsigned int hitCapOverride_ = 0;   // slot in rank 0's gather window) instead of resDev_
    CheckArgs fusedPublish_;            // peer mode: peerHdr / peerDone / peerTicket / peerSeq for the last k_exact
    void peerLaunchCheckAndFinalize(RunSlot &slot);
    bool peerAcquireResultBuf(RunSlot &slot); // worker ranks: a buffer of the rank's shared-memory result ring
    void peerWaitFlag(const uint32_t *flag, uint32_t value);

    std::vector<HitRecord> hits_;   // hits of the run being processed
    // large hit lists are sorted / resolved on the device (see kernels.cuh: PostBuffers)
    static constexpr size_t kPostprocessHits = 8192;
    bool postValid_ = false;
    size_t postSortedOffset_ = 0;
    size_t postN_ = 0;
    int64_t postLits_ = 0;
    int64_t postLitGuess_ = 0; // literal count of the previous large result
    DevBuf<uint8_t> postDev_;            // keys, values, positions, sorted records, CUB scratch
    DevBuf<int32_t> postLitsDev_;
    HostBuf<SortedHit> postSortedHost_;
    HostBuf<int32_t> postLitsHost_;
    HostBuf<long long> postTotalHost_;
    void postprocessOnDevice(RunSlot &slot, size_t n, const HitRecord *hitsDevOverride = nullptr);
    DevBuf<uint8_t> unionDev_; // multi-GPU rank 0: the ranks' hits, concatenated
    // activity bumps on the device: the hit records of the finished run are parked in bumpRecs_
    // before the next run may overwrite the result buffers, and bumped once the next batch of
    // clauses has been drained (same increment as the reference uses at that point)
    void parkHitsForBump(const void *hostOrDevRecs, int stride, size_t n);
    void bumpParkedHits();
    DevBuf<uint8_t> bumpRecs_;
    int bumpStride_ = 0;
    size_t bumpN_ = 0;
    HostBuf<uint8_t> bumpDirHost_;
    DevBuf<uint8_t> bumpDirDev_;
    DevBuf<int> bumpFlagDev_;
    HostBuf<int> bumpFlagHost_;
    bool bumpFlagPending_ = false;
    cudaEvent_t bumpFlagEv_ = nullptr;
    LazyPool pool_; // host-side worker pool (collect of large batches, hand-over of large hit lists)
    std::vector<gss_hit> lastHits_; // sorted, for gss_debug_last_hits (built on demand)
    bool lastHitsValid_ = true;
    bool dense_ = false;
    bool ranOutOfMemory_ = false;
    int64_t launches_ = 0;
    int64_t lastH2D_ = 0, lastD2H_ = 0, finishedD2H_ = 0;
    double lastTimes_[4] = {0, 0, 0, 0};
    double hostPhases_[6] = {0, 0, 0, 0, 0, 0};
    bool haveTimes_ = false;
};

} // namespace gss
