// peer.cu -- multi-GPU exchange over peer memory: no collective on the data path.
//
// One process per GPU.  Every rank allocates a WINDOW in its device memory and exports it with CUDA
// IPC; rank 0 (the front-end: solver threads, slot machines, hand-over) maps every worker's window,
// the workers map rank 0's.  NVLink / NVSwitch then carries plain loads and stores:
//
//   rank 0 window   [control: done flag of every rank][payload: header, run parameters, deltas]
//                   [gather: one slot per rank = 64 B header + hit records]
//   worker window   [control: mailbox][payload]
//
// Data is always PUSHED (posted writes run at link speed; reading a peer's memory is latency-bound:
// measured 2.3 MB in ~30 us when the workers pulled the batch, whichever way the loads were shaped).
// Per batch (sequence number seq):
//   rank 0   collect the batch, ONE H2D copy into its payload area; then k_peer_push stores the
//            payload into every worker's window and, from the block that finishes last, seq into
//            every mailbox; then its own apply / check.  (Measured alternatives, profiles/: copy-engine
//            transfers and a second stream both deliver the batch ~25 us later.)
//   worker   stream waits for mailbox >= seq (cuStreamWaitValue32: nothing spins on an SM), then
//            k_apply_updates (counts come from the payload itself, the host never reads it; it keeps
//            a copy for the deferred collapse), k_filter / k_exact on this rank's tiles, which append
//            the hits straight into this rank's slot of rank 0's gather area (peer stores); the block
//            of k_exact that finishes last writes the slot header and the done flag.
//   workers  as soon as a batch cannot be run again (no overflow), peerFinish enqueues its collapse
//            (dSetAllAssigsToLast): off the critical path of the next batch, on a warm GPU.  Rank 0
//            collapses right after the next push instead: it does not have to wait for the batch
//            to arrive, so its collapse hides behind the workers' later start.
//   rank 0   stream waits for every done flag, reads the headers (one small D2H), and sorts /
//            resolves / hands over the union of the hits on its device (every rank holds the whole
//            clause arena, so rank 0 can resolve any hit)
// A rank whose survivor buffer overflowed flags it in its header (done = 2*seq-1), grows the buffer,
// runs again and then reports done = 2*seq; rank 0 waits for that second flag.  Nothing is dropped.
//
// (The reference drives one device only: gpuShareLib/GpuClauseSharerImpl.cu:52.)
#include "sharer.h"
#include <chrono>
#include <cstring>
#include <thread>
#include <unistd.h>

namespace gss {

namespace {

constexpr uint32_t kBlobMagic = 0x47535350u; // "GSSP"
constexpr size_t kCtlBytes = 4096;           // control block at the start of every window
constexpr size_t kFlagStride = 128;          // one flag per 128 B line
constexpr size_t kMailboxOff = 0;
constexpr size_t kDoneOff = 128;             // done flag of rank r at kDoneOff + r * kFlagStride
constexpr size_t kErrOff = kCtlBytes - 64;   // error word of the polling fallback
constexpr size_t kTicketOff = kCtlBytes - 256; // this device: finished-block counter of the publishing k_exact
constexpr size_t kPushTicketOff = kCtlBytes - 192; // rank 0: finished-block counter of k_peer_push
static_assert(kDoneOff + kMaxPeers * kFlagStride <= kTicketOff, "control block too small");

struct PeerBlob {
    uint32_t magic;
    int32_t rank, world, pid;
    int64_t windowBytes, payloadCap, slotHits;
    cudaIpcMemHandle_t handle;
};
static_assert(sizeof(PeerBlob) <= 128, "blob grew");

inline size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

// driver entry point of the stream memory operation, resolved through the runtime (no libcuda link)
using WaitValue32Fn = int (*)(cudaStream_t, unsigned long long, uint32_t, unsigned int);

} // namespace

struct Sharer::PeerState {
    int rank = 0, world = 1;
    int64_t payloadCap = 0, slotHits = 0;
    size_t slotBytes = 0, windowBytes = 0, payloadArea = 0;
    uint8_t *window = nullptr;          // this rank's window
    uint8_t *rootWindow = nullptr;      // rank 0's window (mapped on the workers)
    std::vector<uint8_t *> mapped;      // windows opened through IPC (to be closed)
    PeerPushList push{};                // rank 0: every worker's payload area and mailbox
    cudaEvent_t evPushed = nullptr;     // rank 0: the batch has been stored into every worker's window
    uint32_t seq = 0;
    WaitValue32Fn waitFn = nullptr;
    bool connected = false;
    HostBuf<long long> headsHost;       // rank 0: world x 8 int64
    HostBuf<int> errHost;
    double timeoutS = 120.0;
    cudaEvent_t evC0 = nullptr, evC1 = nullptr; // around the collapse of this batch (end of peerFinish)
    cudaEvent_t evGathered = nullptr;           // rank 0: every rank's hits are in its memory
    bool collapsed = false, collapsePending = false;
    cudaEvent_t evWait = nullptr;               // workers: the mailbox wait has been satisfied
    bool trace = false;                         // GSS_PEER_TRACE: per-batch phase times on stderr

    uint8_t *payload() const { return (rank == 0 ? rootWindow : window) + kCtlBytes; } // this rank's copy of the batch
    uint8_t *slot(int r) const { return rootWindow + kCtlBytes + payloadArea + (size_t)r * slotBytes; }
    uint32_t *done(int r) const { return reinterpret_cast<uint32_t *>(rootWindow + kDoneOff + (size_t)r * kFlagStride); }
    uint32_t *mailbox() const { return reinterpret_cast<uint32_t *>(window + kMailboxOff); }
    int *err() const { return reinterpret_cast<int *>(window + kErrOff); }
    unsigned int *ticket() const { return reinterpret_cast<unsigned int *>(window + kTicketOff); }
    unsigned int *pushTicket() const { return reinterpret_cast<unsigned int *>(window + kPushTicketOff); }
};

int64_t Sharer::peerInit(int rank, int world, int64_t payloadCap, int64_t slotHits, void *blobOut, int64_t blobCap) {
    useDevice();
    GSS_CHECK(!peer_ && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world);
    GSS_CHECK(blobCap >= (int64_t)sizeof(PeerBlob) && payloadCap >= 4096 && slotHits >= 1);
    GSS_CHECK(db_->shardRank() == rank && db_->shardWorld() == world);
    peer_ = new PeerState();
    PeerState &P = *peer_;
    P.rank = rank;
    P.world = world;
    P.payloadCap = payloadCap;
    P.slotHits = slotHits;
    P.payloadArea = alignUp((size_t)payloadCap, 256);
    P.slotBytes = alignUp(64 + (size_t)slotHits * sizeof(HitRecord), 256);
    P.windowBytes = kCtlBytes + P.payloadArea + (rank == 0 ? (size_t)world * P.slotBytes : 0);
    GSS_CUDA(cudaMalloc(reinterpret_cast<void **>(&P.window), P.windowBytes));
    GSS_CUDA(cudaMemset(P.window, 0, kCtlBytes));
    GSS_CUDA(cudaDeviceSynchronize());
    if (rank == 0) P.rootWindow = P.window;
    if (const char *t = getenv("GPUSHARE_PEER_TIMEOUT_S")) P.timeoutS = atof(t);

    const char *mode = getenv("GPUSHARE_PEER_WAIT"); // "kernel": poll from a one-thread kernel instead
    if (!mode || strcmp(mode, "kernel") != 0) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && fn)
            P.waitFn = reinterpret_cast<WaitValue32Fn>(fn);
        else
            cudaGetLastError();
    }
    logger_.log(1, std::string("c gpushare_b200 peer exchange: rank ") + std::to_string(rank) + "/" + std::to_string(world) +
                       (P.waitFn ? ", stream memory operations\n" : ", polling kernel\n"));

    PeerBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kBlobMagic;
    b.rank = rank;
    b.world = world;
    b.pid = (int32_t)getpid();
    b.windowBytes = (int64_t)P.windowBytes;
    b.payloadCap = payloadCap;
    b.slotHits = slotHits;
    GSS_CUDA(cudaIpcGetMemHandle(&b.handle, P.window));
    memcpy(blobOut, &b, sizeof(b));
    return (int64_t)sizeof(PeerBlob);
}

// blobs: the world blobs of peerInit in rank order, `blobBytes` apart
void Sharer::peerConnect(const void *blobs, int64_t blobBytes) {
    useDevice();
    GSS_CHECK(peer_ && !peer_->connected && blobBytes >= (int64_t)sizeof(PeerBlob));
    PeerState &P = *peer_;
    auto blobOf = [&](int r) {
        PeerBlob b;
        memcpy(&b, static_cast<const uint8_t *>(blobs) + (size_t)r * (size_t)blobBytes, sizeof(b));
        if (b.magic != kBlobMagic || b.rank != r || b.world != P.world || b.payloadCap != P.payloadCap || b.slotHits != P.slotHits)
            GSS_DIE("peer exchange: the ranks disagree about the window parameters");
        return b;
    };
    auto open = [&](const PeerBlob &b) {
        void *p = nullptr;
        GSS_CUDA(cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess));
        P.mapped.push_back(static_cast<uint8_t *>(p));
        return static_cast<uint8_t *>(p);
    };
    if (P.rank == 0) {
        for (int r = 1; r < P.world; r++) {
            uint8_t *w = open(blobOf(r));
            P.push.dst[P.push.n] = reinterpret_cast<uint4 *>(w + kCtlBytes);
            P.push.mailbox[P.push.n] = reinterpret_cast<uint32_t *>(w + kMailboxOff);
            P.push.n++;
        }
        GSS_CUDA(cudaEventCreate(&P.evPushed));
        P.headsHost.resize((size_t)P.world * 8);
    } else {
        P.rootWindow = open(blobOf(0));
    }
    P.errHost.resize(1);
    GSS_CUDA(cudaEventCreate(&P.evC0));
    GSS_CUDA(cudaEventCreate(&P.evC1));
    GSS_CUDA(cudaEventCreate(&P.evGathered));
    GSS_CUDA(cudaEventCreate(&P.evWait));
    P.trace = getenv("GSS_PEER_TRACE") != nullptr;
    P.connected = true;
}

void Sharer::peerClose() {
    if (!peer_) return;
    for (uint8_t *p : peer_->mapped) cudaIpcCloseMemHandle(p);
    if (peer_->evC0) cudaEventDestroy(peer_->evC0);
    if (peer_->evC1) cudaEventDestroy(peer_->evC1);
    if (peer_->evGathered) cudaEventDestroy(peer_->evGathered);
    if (peer_->evWait) cudaEventDestroy(peer_->evWait);
    if (peer_->evPushed) cudaEventDestroy(peer_->evPushed);
    if (peer_->window) cudaFree(peer_->window);
    delete peer_;
    peer_ = nullptr;
}

void Sharer::peerLaunchCheckAndFinalize(RunSlot &slot) {
    PeerState &P = *peer_;
    uint8_t *mySlot = P.slot(P.rank);
    hitsOverride_ = reinterpret_cast<HitRecord *>(mySlot + 64);
    hitCapOverride_ = (unsigned int)P.slotHits;
    fusedPublish_.peerHdr = reinterpret_cast<long long *>(mySlot);
    fusedPublish_.peerDone = P.done(P.rank);
    fusedPublish_.peerTicket = P.ticket();
    fusedPublish_.peerSeq = P.seq;
    const bool published = launchCheckKernels(slot, slot.dense);
    hitsOverride_ = nullptr;
    fusedPublish_.peerDone = nullptr;
    if (published) return; // the last k_exact wrote the slot header and the done flag itself
    // nothing was launched that could publish (no tile on this rank, no frozen slot, dense mode)
    int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    launchPeerFinalize((const Counters *)resDev_.data(), (unsigned int)P.slotHits, (unsigned int)survCap_, groups,
                       reinterpret_cast<long long *>(mySlot), P.done(P.rank), P.seq, stream_, &launches_);
}

// make the stream wait until *flag >= value (flag lives in THIS device's memory; another GPU stores it)
void Sharer::peerWaitFlag(const uint32_t *flag, uint32_t value) {
    PeerState &P = *peer_;
    if (P.waitFn) {
        int r = P.waitFn(stream_, (unsigned long long)(uintptr_t)flag, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (r != 0) GSS_DIE("cuStreamWaitValue32 failed with driver error " + std::to_string(r));
    } else {
        launchPeerWait(flag, value, (unsigned long long)(P.timeoutS * 1e9), P.err(), stream_, &launches_);
    }
}

// Start one batch on this rank.  Returns -1 when there is no clause yet (every rank holds the same
// clause stream, so every rank returns -1 together), 1 when this batch rebuilt the tables, else 0.
int Sharer::peerEnqueue() {
    useDevice();
    GSS_CHECK(peer_ && peer_->connected && cur_ < 0 && mgpuPending_ < 0);
    PeerState &P = *peer_;
    const bool root = P.rank == 0;
    db_->drainPending();
    if (db_->stats().clauses == 0) return -1;
    RunSlot &slot = slots_[nextSlot()];
    // The collapse of the previous batch (dSetAllAssigsToLast).  A worker did it in peerFinish, as soon
    // as that batch could no longer be run again (only a batch left over from before peer mode goes
    // here).  Rank 0 does it below, right after the push: it is the one rank that does not wait for the
    // batch to arrive, so there the collapse hides behind the workers' later start.
    if (!root && collapseSlot_ >= 0) {
        RunSlot &c = slots_[collapseSlot_];
        launchCollapse(c.updDev.data(), c.paramsDev(), c.nSolvers, c.maxUpd, c.nUpdates, tables_, numSMs_, stream_, &launches_);
        collapseSlot_ = -1;
    }
    bool rebuild = false;
    int64_t h2d = 0;
    if (!prepareRun(slot, rebuild, h2d)) GSS_DIE("out of device memory (multi-GPU mode does not reduce the database by itself)");
    P.seq++;
    const size_t PR = payloadPrefixRecords(slot.nSolvers);
    GSS_CHECK((int64_t)(PR * sizeof(VarUpdate)) <= P.payloadCap);
    uint8_t *payload = P.payload(); // this device's copy of the batch (rank 0 pushes it into the workers' windows)
    const int groups = (slot.nSolvers + kMaxSolversPerGroup - 1) / kMaxSolversPerGroup;
    slot.dense = dense_;
    slot.headDev.reserve(slot.headHost.size(), 0, stream_);

    if (root) {
        collectBatch(slot, rebuild);
        PayloadHeader hdr;
        memset(&hdr, 0, sizeof(hdr));
        hdr.magic = kPayloadMagic;
        hdr.status = rebuild ? 1 : 0;
        hdr.nSolvers = slot.nSolvers;
        hdr.prefixRecords = (int32_t)PR;
        hdr.nUpdates = slot.nUpdates;
        hdr.totalBytes = (int64_t)((PR + (size_t)slot.nUpdates) * sizeof(VarUpdate));
        if (hdr.totalBytes > P.payloadCap) GSS_DIE("peer exchange: the batch does not fit the payload area (raise payload_cap)");
        const SolverRunParams *params = (const SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
        uint8_t *base = reinterpret_cast<uint8_t *>(slot.updHost.data());
        memcpy(base, &hdr, sizeof(hdr));
        memcpy(base + sizeof(hdr), params, (size_t)slot.nSolvers * sizeof(SolverRunParams));
        GSS_CUDA(cudaMemcpyAsync(payload, base, (size_t)hdr.totalBytes, cudaMemcpyHostToDevice, stream_));
        h2d += hdr.totalBytes;
        peerPayloadBytes_ = hdr.totalBytes;
        slot.aggStart.assign(groups, 0u);
        slot.aggOnDevice = false;
        slot.maxUpd = 0;
        for (int s = 0; s < slot.nSolvers; s++) {
            slot.aggStart[s / kMaxSolversPerGroup] |= params[s].usedAggBits;
            slot.maxUpd = std::max(slot.maxUpd, (int)params[s].updCount);
        }
        GSS_CUDA(cudaMemcpyAsync(slot.headDev.data(), slot.headHost.data(), slot.headHost.size(), cudaMemcpyHostToDevice, stream_));
        h2d += (int64_t)slot.headHost.size();
    } else {
        slot.ids.assign(slot.nSolvers, AssigIds{});
        slot.assigCount = 0;
        int64_t nUpper = P.payloadCap / (int64_t)sizeof(VarUpdate) - (int64_t)PR;
        slot.aggStart.assign(groups, ~0u);
        slot.aggOnDevice = true; // the host never sees the run parameters: the kernels read them
        slot.maxUpd = (int)std::min<int64_t>(nUpper, 1 << 30);
        slot.nUpdates = nUpper;
        // everything that does not need the batch goes ahead of the wait: the length directory,
        // buffer growth (the run parameters are not copied here at all: k_apply_updates reads them
        // from the pulled payload and leaves a copy in headDev for the kernels that follow)
        GSS_CUDA(cudaMemcpyAsync(slot.headDev.data(), slot.headHost.data(), slot.dirBytes, cudaMemcpyHostToDevice, stream_));
        h2d += (int64_t)slot.dirBytes;
        slot.updDev.reserve((size_t)std::max<int64_t>(slot.nUpdates, 1), 0, stream_);
        ensureResultBuffers();
        // rank 0 has pushed this batch into this rank's window once the mailbox says so
        peerWaitFlag(P.mailbox(), P.seq);
        GSS_CUDA(cudaEventRecord(P.evWait, stream_));
    }
    slot.updDev.reserve((size_t)std::max<int64_t>(slot.nUpdates, 1), 0, stream_);
    ensureResultBuffers();
    GSS_CUDA(cudaEventRecord(slot.evH2DDone, stream_));
    if (root && P.push.n) {
        // push the batch into every worker's window and signal: one kernel, FIRST on rank 0's stream
        // (on a second stream it reached the workers ~25 us later: it competed with rank 0's own table
        // kernels).  The workers start ~12 us after rank 0 and have no push to do: the ranks finish
        // within a few microseconds of each other.
        launchPeerPush(payload, peerPayloadBytes_, P.push, P.seq, P.pushTicket(), numSMs_, stream_, &launches_);
        GSS_CUDA(cudaEventRecord(P.evPushed, stream_));
    }
    if (root && collapseSlot_ >= 0) {
        RunSlot &c = slots_[collapseSlot_];
        launchCollapse(c.updDev.data(), c.paramsDev(), c.nSolvers, c.maxUpd, c.nUpdates, tables_, numSMs_, stream_, &launches_);
        collapseSlot_ = -1;
    }

    const SolverRunParams *paramsSrc = root ? slot.paramsDev() : reinterpret_cast<const SolverRunParams *>(payload + sizeof(PayloadHeader));
    SolverRunParams *paramsKeep = root ? nullptr : const_cast<SolverRunParams *>(slot.paramsDev());
    launchApplyUpdates(reinterpret_cast<const VarUpdate *>(payload + PR * sizeof(VarUpdate)), paramsSrc, slot.nSolvers,
                       slot.maxUpd, slot.nUpdates, tables_, numSMs_, stream_, &launches_, slot.updDev.data(), paramsKeep);
    GSS_CUDA(cudaEventRecord(slot.evBeforeCheck, stream_));
    peerLaunchCheckAndFinalize(slot);
    GSS_CUDA(cudaEventRecord(slot.evAfterCheck, stream_));
    if (root) {
        if (P.world > 2) { // all workers' flags with one launch
            PeerFlagList all{};
            for (int r = 1; r < P.world; r++) all.p[all.n++] = P.done(r);
            launchPeerWaitAll(all, 2u * P.seq - 1u, (unsigned long long)(P.timeoutS * 1e9), P.err(), stream_, &launches_);
        } else if (P.world == 2) {
            peerWaitFlag(P.done(1), 2u * P.seq - 1u);
        }
    }
    GSS_CUDA(cudaEventRecord(P.evGathered, stream_)); // rank 0: the hits of every rank are in its memory
    // (small copies cost ~10 us of latency each: they go after the point the batch is complete)
    slot.resHost.resize(sizeof(Counters));
    GSS_CUDA(cudaMemcpyAsync(slot.resHost.data(), resDev_.data(), sizeof(Counters), cudaMemcpyDeviceToHost, stream_));
    if (root)
        GSS_CUDA(cudaMemcpy2DAsync(P.headsHost.data(), 64, P.slot(0), P.slotBytes, 64, (size_t)P.world, cudaMemcpyDeviceToHost, stream_));
    GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
    slot.inFlight = true;
    P.collapsePending = slot.nUpdates != 0;
    lastStarted_ = (int)(&slot - slots_);
    lastH2D_ = h2d;
    cur_ = (int)(&slot - slots_);
    return rebuild ? 1 : 0;
}

// Close the batch: wait, repair overflows, and (rank 0) hand the union of the hits to the solver
// queues.  Returns the number of hit records: of all ranks on rank 0, of this rank on a worker.
int64_t Sharer::peerFinish() {
    useDevice();
    GSS_CHECK(peer_ && cur_ >= 0);
    PeerState &P = *peer_;
    RunSlot &slot = slots_[cur_];
    const bool root = P.rank == 0;
    auto waitEnd = [&]() {
        // bounded wait: a protocol error must not hang the device box
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            cudaError_t e = cudaEventQuery(slot.evEnd);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) GSS_DIE(std::string("CUDA error ") + cudaGetErrorString(e) + " while waiting for the peer exchange");
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > P.timeoutS)
                GSS_DIE("peer exchange timed out (rank " + std::to_string(P.rank) + ", batch " + std::to_string(P.seq) + ")");
            std::this_thread::yield();
        }
        if (!P.waitFn || (root && P.world > 2)) {
            GSS_CUDA(cudaMemcpyAsync(P.errHost.data(), P.err(), sizeof(int), cudaMemcpyDeviceToHost, stream_));
            GSS_CUDA(cudaStreamSynchronize(stream_));
            if (P.errHost[0]) GSS_DIE("peer exchange timed out on the device (rank " + std::to_string(P.rank) + ")");
        }
    };
    waitEnd();

    // this rank's own overflow: grow and run again (the tables are intact: collapse is deferred)
    auto repairOwn = [&]() -> bool {
        Counters c;
        memcpy(&c, slot.resHost.data(), sizeof(c));
        if ((int64_t)c.nHits > P.slotHits)
            GSS_DIE("peer exchange: " + std::to_string(c.nHits) + " hits of rank " + std::to_string(P.rank) +
                    " do not fit its gather slot (raise slot_hits)");
        size_t maxSurv = 0;
        for (int g = 0; g < kMaxGroups; g++) maxSurv = std::max(maxSurv, (size_t)c.nSurvivors[g]);
        if (maxSurv <= survCap_) return false;
        survCap_ = std::max(survCap_ * 2, maxSurv + maxSurv / 4);
        ensureResultBuffers();
        peerLaunchCheckAndFinalize(slot);
        GSS_CUDA(cudaMemcpyAsync(slot.resHost.data(), resDev_.data(), sizeof(Counters), cudaMemcpyDeviceToHost, stream_));
        return true;
    };

    // Once the batch cannot be run again, its tables are no longer needed: collapse every slot to the
    // solver's last one right away (dSetAllAssigsToLast) -- the GPU is warm and otherwise idle while
    // the host hands the hits over / rank 0 prepares the next batch.  Its time is charged to this batch.
    float msFirst[5] = {0, 0, 0, 0, 0}; // first attempt: copy, apply, check, total, tail
    cudaEventElapsedTime(&msFirst[0], slot.evStart, slot.evH2DDone);
    cudaEventElapsedTime(&msFirst[1], slot.evH2DDone, slot.evBeforeCheck);
    cudaEventElapsedTime(&msFirst[2], slot.evBeforeCheck, slot.evAfterCheck);
    cudaEventElapsedTime(&msFirst[3], slot.evStart, P.evGathered);
    cudaEventElapsedTime(&msFirst[4], slot.evAfterCheck, P.evGathered);
    auto collapseNow = [&]() {
        if (root) { // rank 0: at the start of the next batch, behind the push (see peerEnqueue)
            P.collapsed = false;
            collapseSlot_ = P.collapsePending ? cur_ : -1;
            P.collapsePending = false;
            return;
        }
        P.collapsed = P.collapsePending;
        if (!P.collapsed) return;
        GSS_CUDA(cudaEventRecord(P.evC0, stream_));
        launchCollapse(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_, stream_, &launches_);
        GSS_CUDA(cudaEventRecord(P.evC1, stream_));
        P.collapsePending = false;
        collapseSlot_ = -1;
        lastStarted_ = -1; // the tables of this batch are gone
    };
    auto recordTimes = [&]() {
        float msCollapse = 0;
        if (P.collapsed) {
            GSS_CUDA(cudaEventSynchronize(P.evC1));
            cudaEventElapsedTime(&msCollapse, P.evC0, P.evC1);
        }
        // [0] prepare + H2D (+ wait for the batch on a worker), [1] table kernels (push on rank 0, apply,
        // collapse), [2] check kernels, [3] everything from the start to "hits gathered" plus the
        // collapse: [3] - [0] = device time of the batch on this rank
        lastTimes_[0] = msFirst[0] * 1000.0;
        lastTimes_[1] = (msFirst[1] + msCollapse) * 1000.0;
        lastTimes_[2] = msFirst[2] * 1000.0;
        lastTimes_[3] = (msFirst[3] + msCollapse) * 1000.0;
        haveTimes_ = true;
        if (opts_.quickProf) globalStats_[G_timeSpentTestingClauses] += (uint64_t)(msFirst[2] * 1000.0f);
        if (P.trace) {
            float msWait = 0, msAfterWait = 0, msPush = 0;
            if (!root) {
                cudaEventElapsedTime(&msWait, slot.evStart, P.evWait);
                cudaEventElapsedTime(&msAfterWait, P.evWait, slot.evH2DDone);
            } else if (P.push.n) {
                cudaEventElapsedTime(&msPush, slot.evH2DDone, P.evPushed);
            }
            fprintf(stderr, "peer trace rank %d batch %u: start->batch-arrived %.1f | ->ready %.1f | push %.1f | push+apply %.1f | check %.1f | tail %.1f | collapse %.1f us\n",
                    P.rank, P.seq, msWait * 1e3, msAfterWait * 1e3, msPush * 1e3, msFirst[1] * 1e3, msFirst[2] * 1e3, msFirst[4] * 1e3, msCollapse * 1e3);
        }
    };

    int64_t total = 0;
    if (!root) {
        for (int attempt = 0; repairOwn(); attempt++) {
            GSS_CHECK(attempt < 8);
            GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
            waitEnd();
        }
        collapseNow();
        Counters c;
        memcpy(&c, slot.resHost.data(), sizeof(c));
        globalStats_[G_clauseTestsOnAssigs] += c.exactTests;
        total = (int64_t)c.nHits;
        slot.inFlight = false;
        mgpuLast_ = cur_;
        cur_ = -1;
        recordTimes();
        return total;
    }

    std::vector<int64_t> counts(P.world, 0);
    for (int attempt = 0;; attempt++) {
        GSS_CHECK(attempt < 16);
        bool again = repairOwn();
        for (int r = 1; r < P.world; r++) {
            const long long flags = P.headsHost[(size_t)r * 8 + 1];
            if (flags & 2) GSS_DIE("peer exchange: the hits of rank " + std::to_string(r) + " do not fit its gather slot (raise slot_hits)");
            if (flags & 1) { // that rank is running again: its second flag follows
                peerWaitFlag(P.done(r), 2u * P.seq);
                again = true;
            }
        }
        if (!again) break;
        GSS_CUDA(cudaMemcpy2DAsync(P.headsHost.data(), 64, P.slot(0), P.slotBytes, 64, (size_t)P.world, cudaMemcpyDeviceToHost, stream_));
        GSS_CUDA(cudaEventRecord(slot.evEnd, stream_));
        waitEnd();
    }
    for (int r = 0; r < P.world; r++) {
        GSS_CHECK(P.headsHost[(size_t)r * 8 + 3] == (long long)P.seq);
        counts[r] = P.headsHost[(size_t)r * 8];
        total += counts[r];
        globalStats_[G_clauseTestsOnAssigs] += (uint64_t)P.headsHost[(size_t)r * 8 + 2];
    }
    collapseNow();
    slot.inFlight = false;
    mgpuLast_ = cur_;
    cur_ = -1;
    finishedD2H_ = (int64_t)((size_t)P.world * 64 + sizeof(Counters));
    mgpuImportGathered(P.slot(0), P.world, (int64_t)P.slotBytes, counts.data());
    recordTimes();
    return total;
}

} // namespace gss
