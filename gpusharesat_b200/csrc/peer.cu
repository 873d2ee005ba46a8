// peer.cu -- multi-GPU exchange over peer memory: no collective on the data path.
//
// One process per GPU.  Every rank allocates a WINDOW in its device memory and exports it with CUDA
// IPC; rank 0 (the front-end: solver threads, slot machines, hand-over) maps every worker's window,
// the workers map rank 0's.  NVLink / NVSwitch then carries plain loads and stores:
//
//   rank 0 window   [control][payload: header, run parameters, deltas]
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
//            a copy for the deferred collapse), k_filter / k_exact on this rank's tiles (results: below)
//   workers  as soon as a batch cannot be run again (no overflow), peerFinish enqueues its collapse
//            (dSetAllAssigsToLast): off the critical path of the next batch, on a warm GPU.  Rank 0
//            collapses right after the next push instead: it does not have to wait for the batch
//            to arrive, so its collapse hides behind the workers' later start.
//   results  (round 2) every rank runs the direct pipeline's result path on its own device: per-solver
//            record buckets, k_emit_sort, k_emit_write -- and k_emit_write writes the rank's FINISHED
//            per-solver results (clause ids, literal positions, literal stream, plus the sorted
//            records for rank 0's activity bumps) straight into HOST memory over the rank's own PCIe
//            link: a ring of buffers in a POSIX shared-memory segment the rank owns and rank 0 maps.
//            Rank 0 only stitches views: a solver's ClauseBatch views one slice per rank.  Nothing
//            is funnelled through rank 0's GPU or its PCIe link any more (round 1: hits by peer stores
//            into rank 0's window, then sort / resolve / D2H of the UNION on rank 0: e2e weak
//            efficiency 0.31 at 8 GPUs).  A rank whose buffers overflowed repairs that by itself
//            (re-launch with larger buffers) before it publishes; rank 0 just waits for it.
//   hand-shake  a worker publishes (sequence number, buffer index) in its ring's control block after
//            its own event has completed (so its GPU's writes are visible to every CPU); rank 0's
//            host polls that word.
//
// (The reference drives one device only: gpuShareLib/GpuClauseSharerImpl.cu:52.)
#include "sharer.h"
#include <chrono>
#include <cstring>
#include <atomic>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
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

// ---- a rank's result buffers in a POSIX shared-memory segment ----
constexpr int kRingBufs = 4;
constexpr uint32_t kRingMagic = 0x47535352u; // "GSSR"
struct RingCtl {
    uint32_t magic;
    int32_t nBufs;
    int64_t bufBytes, entryCap, litCap;
    std::atomic<uint64_t> published; // (sequence number << 8) | buffer index, written by the owner after its event completed
    uint8_t pad0[64 - 40];
    struct {
        std::atomic<uint32_t> state; // 0 free, 1 owned by the producer / still viewed by rank 0
        uint8_t pad[60];
    } buf[kRingBufs];
};
static_assert(sizeof(RingCtl) == 64 * (1 + kRingBufs), "control block layout");
constexpr size_t kRingCtlBytes = 4096;

struct ShmRing {
    std::string name;
    uint8_t *base = nullptr;
    size_t bytes = 0;
    bool owner = false, registered = false;
    RingCtl *ctl() const { return reinterpret_cast<RingCtl *>(base); }
    uint8_t *buf(int k) const { return base + kRingCtlBytes + (size_t)k * (size_t)ctl()->bufBytes; }

    static std::string nameOf(int pid, int rank) { return "/gss_b200_" + std::to_string(pid) + "_" + std::to_string(rank); }
    void create(int pid, int rank, int64_t entryCap, int64_t litCap) {
        name = nameOf(pid, rank);
        const size_t bufBytes = (Sharer::RunBuf::bytesFor(entryCap, litCap, true) + 4095) / 4096 * 4096;
        bytes = kRingCtlBytes + kRingBufs * bufBytes;
        shm_unlink(name.c_str());
        int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        // (posix_fallocate: a full /dev/shm must be an error here, not a SIGBUS on the first store)
        if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0 || posix_fallocate(fd, 0, (off_t)bytes) != 0) {
            if (fd >= 0) shm_unlink(name.c_str());
            GSS_DIE("cannot create the shared-memory result ring " + name + " (" + std::to_string(bytes) + " bytes; is /dev/shm large enough?)");
        }
        map(fd);
        memset(base, 0, kRingCtlBytes);
        RingCtl *c = ctl();
        c->nBufs = kRingBufs;
        c->bufBytes = (int64_t)bufBytes;
        c->entryCap = entryCap;
        c->litCap = litCap;
        c->magic = kRingMagic;
        owner = true;
    }
    void open(int pid, int rank) {
        name = nameOf(pid, rank);
        int fd = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd < 0) GSS_DIE("cannot open the shared-memory result ring " + name);
        struct stat st;
        if (fstat(fd, &st) != 0) GSS_DIE("cannot stat " + name);
        bytes = (size_t)st.st_size;
        map(fd);
        if (ctl()->magic != kRingMagic) GSS_DIE("shared-memory result ring " + name + " is not initialised");
    }
    void map(int fd) {
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
        ::close(fd);
        if (p == MAP_FAILED) GSS_DIE("cannot map the shared-memory result ring " + name);
        base = static_cast<uint8_t *>(p);
        // page-locked + mapped: this process's device reads / writes it in place (unified addressing)
        GSS_CUDA(cudaHostRegister(base, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        registered = true;
    }
    ~ShmRing() { close(); }
    void close() {
        if (!base) return;
        if (registered) cudaHostUnregister(base);
        munmap(base, bytes);
        if (owner) shm_unlink(name.c_str());
        base = nullptr;
    }
};

// driver entry point of the stream memory operation, resolved through the runtime (no libcuda link)
using WaitValue32Fn = int (*)(cudaStream_t, unsigned long long, uint32_t, unsigned int);

} // namespace

struct Sharer::PeerState {
    int rank = 0, world = 1;
    int64_t payloadCap = 0, slotHits = 0;
    size_t windowBytes = 0, payloadArea = 0;
    uint8_t *window = nullptr;          // this rank's window
    uint8_t *rootWindow = nullptr;      // rank 0's window (mapped on the workers)
    std::vector<uint8_t *> mapped;      // windows opened through IPC (to be closed)
    PeerPushList push{};                // rank 0: every worker's payload area and mailbox
    cudaEvent_t evPushed = nullptr;     // rank 0: the batch has been stored into every worker's window
    uint32_t seq = 0;
    WaitValue32Fn waitFn = nullptr;
    bool connected = false;
    HostBuf<int> errHost;
    double timeoutS = 120.0;
    cudaEvent_t evC0 = nullptr, evC1 = nullptr; // around the collapse of this batch (end of peerFinish)
    cudaEvent_t evGathered = nullptr;           // rank 0: every rank's hits are in its memory
    bool collapsed = false, collapsePending = false;
    cudaEvent_t evWait = nullptr;               // workers: the mailbox wait has been satisfied
    bool trace = false;                         // GSS_PEER_TRACE: per-batch phase times on stderr
    std::shared_ptr<ShmRing> ring;              // workers: this rank's result buffers
    std::vector<std::shared_ptr<ShmRing>> rings; // rank 0: every worker's ring (index = rank; [0] unused)
    int ringBuf = -1;                           // workers: buffer of the batch in flight (not yet published)
    bool exportRecords = false;                 // workers: sorted record keys + masks go out with the results (GPUSHARE_PEER_RECORDS=1, gss_debug_set_peer_records)
    bool directPush = true;                     // rank 0: deltas read in place by the push kernel (GPUSHARE_PEER_STAGED: staged copy + H2D)

    uint8_t *payload() const { return (rank == 0 ? rootWindow : window) + kCtlBytes; } // this rank's copy of the batch
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
    P.windowBytes = kCtlBytes + P.payloadArea; // (round 1 also had one gather slot per rank here: results go to host memory now)
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

    if (rank > 0) {
        P.ring = std::make_shared<ShmRing>();
        P.ring->create((int)getpid(), rank, slotHits, slotHits * 8);
    }
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
        P.rings.resize((size_t)P.world);
        for (int r = 1; r < P.world; r++) {
            P.rings[r] = std::make_shared<ShmRing>();
            P.rings[r]->open(blobOf(r).pid, r);
        }
        foreignCopyOut_ = getenv("GPUSHARE_PEER_COPY_OUT") != nullptr;
        for (int r = 1; r < P.world; r++) {
            uint8_t *w = open(blobOf(r));
            P.push.dst[P.push.n] = reinterpret_cast<uint4 *>(w + kCtlBytes);
            P.push.mailbox[P.push.n] = reinterpret_cast<uint32_t *>(w + kMailboxOff);
            P.push.n++;
        }
        GSS_CUDA(cudaEventCreate(&P.evPushed));
    } else {
        P.rootWindow = open(blobOf(0));
    }
    P.errHost.resize(1);
    GSS_CUDA(cudaEventCreate(&P.evC0));
    GSS_CUDA(cudaEventCreate(&P.evC1));
    GSS_CUDA(cudaEventCreate(&P.evGathered));
    GSS_CUDA(cudaEventCreate(&P.evWait));
    P.trace = getenv("GSS_PEER_TRACE") != nullptr;
    P.directPush = directEnabled_ && getenv("GPUSHARE_PEER_STAGED") == nullptr;
    if (const char *r = getenv("GPUSHARE_PEER_RECORDS")) P.exportRecords = atoi(r) != 0;
    if (peerRecordsWanted_ >= 0) P.exportRecords = peerRecordsWanted_ != 0;
    P.connected = true;
}

void Sharer::peerClose() {
    if (!peer_) return;
    bumpOwners_.clear();
    lastForeign_.clear();
    peer_->ring.reset(); // (the mappings live until the last batch that views them has gone)
    peer_->rings.clear();
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

// Workers: also write the sorted record keys and masks next to the results (rank 0 then bumps the activities of every
// rank's hits on its device and gss_debug_last_hits covers every rank); off: every rank bumps its own hits.
void Sharer::setPeerRecords(bool on) {
    peerRecordsWanted_ = on ? 1 : 0;
    if (peer_) peer_->exportRecords = on;
}

// worker ranks: the batch's result buffer comes from the rank's shared-memory ring (rank 0 reads it there)
bool Sharer::peerAcquireResultBuf(RunSlot &slot) {
    PeerState &P = *peer_;
    if (P.rank == 0 || !P.ring) return false;
    ShmRing &R = *P.ring;
    RingCtl *c = R.ctl();
    if (entryGuess_ > c->entryCap || litGuess_ > c->litCap)
        GSS_DIE("peer exchange: the hits of rank " + std::to_string(P.rank) + " do not fit its result buffers (raise slot_hits)");
    if (P.ringBuf >= 0) c->buf[P.ringBuf].state.store(0, std::memory_order_release); // a repeated batch: its first buffer was never published
    P.ringBuf = -1;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        for (int k = 0; k < c->nBufs && P.ringBuf < 0; k++) {
            uint32_t expect = 0;
            if (c->buf[k].state.compare_exchange_strong(expect, 1u, std::memory_order_acq_rel)) P.ringBuf = k;
        }
        if (P.ringBuf >= 0) break;
        // every buffer is still viewed by batches of rank 0's solvers: wait for one to be retired
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > P.timeoutS)
            GSS_DIE("peer exchange: no free result buffer on rank " + std::to_string(P.rank) + " (rank 0's solvers do not pop)");
        std::this_thread::yield();
    }
    RunBuf *b = new RunBuf();
    b->base = R.buf(P.ringBuf);
    b->bytes = (size_t)c->bufBytes;
    b->entryCap = c->entryCap;
    b->litCap = c->litCap;
    b->withRecords = P.exportRecords;
    std::shared_ptr<ShmRing> keep = P.ring;
    slot.runBuf = std::shared_ptr<RunBuf>(b, [keep](RunBuf *q) { delete q; }); // (rank 0 frees the ring buffer)
    return true;
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
    PhaseTimer tEnqueue(hostPhases_[1]);
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
    bool pushedDirect = false;

    if (root && P.directPush) {
        // Direct variant (default): no staging copy, no separate H2D of the deltas.  The solvers' delta buffers are
        // swapped out (collectDirect) and k_peer_push_direct reads them in place over PCIe, storing every record into
        // this rank's payload area AND every worker's window in the same pass.
        collectDirect(slot, rebuild); // slot.headHost = [directory][run parameters][per-solver delta pointers]
        const int S = slot.nSolvers;
        const size_t prefixOff = alignUp(slot.headHost.size(), 16), prefixBytes = PR * sizeof(VarUpdate);
        slot.headHost.resize(prefixOff + prefixBytes); // (may move the buffer: pointers into it are taken below)
        SolverRunParams *params = (SolverRunParams *)(slot.headHost.data() + slot.dirBytes);
        const VarUpdate **src = (const VarUpdate **)(slot.headHost.data() + slot.srcOff);
        PayloadHeader hdr;
        memset(&hdr, 0, sizeof(hdr));
        hdr.magic = kPayloadMagic;
        hdr.status = rebuild ? 1 : 0;
        hdr.nSolvers = S;
        hdr.prefixRecords = (int32_t)PR;
        hdr.nUpdates = slot.nUpdates;
        hdr.totalBytes = (int64_t)((PR + (size_t)slot.nUpdates) * sizeof(VarUpdate));
        if (hdr.totalBytes > P.payloadCap) GSS_DIE("peer exchange: the batch does not fit the payload area (raise payload_cap)");
        memcpy(slot.headHost.data() + prefixOff, &hdr, sizeof(hdr));
        memcpy(slot.headHost.data() + prefixOff + sizeof(hdr), params, (size_t)S * sizeof(SolverRunParams));
        VarUpdate *updLocal = reinterpret_cast<VarUpdate *>(payload) + PR;
        for (auto &st : slot.staged) { // delta buffers outside page-locked memory go up as ordinary copies, straight into place
            VarUpdate *dst = updLocal + params[st.first].updStart;
            GSS_CUDA(cudaMemcpyAsync(dst, st.second, (size_t)params[st.first].updCount * sizeof(VarUpdate), cudaMemcpyHostToDevice, stream_));
            src[st.first] = dst;
        }
        for (int sIdx = 0; sIdx < S; sIdx++)
            if (!src[sIdx]) src[sIdx] = updLocal + params[sIdx].updStart; // (a skipped solver: zero records)
        slot.headDev.reserve(slot.headHost.size(), 0, stream_);
        GSS_CUDA(cudaMemcpyAsync(slot.headDev.data(), slot.headHost.data(), slot.headHost.size(), cudaMemcpyHostToDevice, stream_));
        h2d += (int64_t)slot.headHost.size() + hdr.totalBytes;
        peerPayloadBytes_ = hdr.totalBytes;
        GSS_CUDA(cudaEventRecord(slot.evH2DDone, stream_));
        launchPeerPushDirect((const VarUpdate *const *)(slot.headDev.data() + slot.srcOff), slot.paramsDev(), S, slot.maxUpd, payload,
                             slot.headDev.data() + prefixOff, (long long)(prefixBytes / 4), P.push, P.seq, P.pushTicket(), numSMs_,
                             stream_, &launches_);
        GSS_CUDA(cudaEventRecord(P.evPushed, stream_));
        pushedDirect = true;
    } else if (root) {
        {
            PhaseTimer tCollect(hostPhases_[3]);
            collectBatch(slot, rebuild);
        }
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
    if (!pushedDirect) GSS_CUDA(cudaEventRecord(slot.evH2DDone, stream_));
    if (root && P.push.n && !pushedDirect) {
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
    // the direct pipeline's result path on this rank's own tiles: record buckets, k_emit_sort, k_emit_write into
    // host memory (rank 0: a buffer of its pool; workers: a buffer of their shared-memory ring)
    slot.direct = true;
    launchDirectCheck(slot);
    GSS_CUDA(cudaEventRecord(slot.evAfterCheck, stream_));
    GSS_CUDA(cudaEventRecord(P.evGathered, stream_));
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
    PhaseTimer tFinish(hostPhases_[0]);
    {
        // bounded wait: a protocol error must not hang the device box
        PhaseTimer tWait(hostPhases_[4]);
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            cudaError_t e = cudaEventQuery(slot.evEnd);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) GSS_DIE(std::string("CUDA error ") + cudaGetErrorString(e) + " while waiting for the peer exchange");
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > P.timeoutS)
                GSS_DIE("peer exchange timed out (rank " + std::to_string(P.rank) + ", batch " + std::to_string(P.seq) + ")");
            std::this_thread::yield();
        }
        if (!P.waitFn && !root) {
            GSS_CUDA(cudaMemcpyAsync(P.errHost.data(), P.err(), sizeof(int), cudaMemcpyDeviceToHost, stream_));
            GSS_CUDA(cudaStreamSynchronize(stream_));
            if (P.errHost[0]) GSS_DIE("peer exchange timed out on the device (rank " + std::to_string(P.rank) + ")");
        }
    }
    float msFirst[5] = {0, 0, 0, 0, 0}; // first attempt: copy, apply, check, total, tail
    cudaEventElapsedTime(&msFirst[0], slot.evStart, slot.evH2DDone);
    cudaEventElapsedTime(&msFirst[1], slot.evH2DDone, slot.evBeforeCheck);
    cudaEventElapsedTime(&msFirst[2], slot.evBeforeCheck, slot.evAfterCheck);
    cudaEventElapsedTime(&msFirst[3], slot.evStart, P.evGathered);
    // this rank's own result: complete, or repaired here (overflow: larger buffers, same batch again --
    // the tables are intact because the collapse is deferred)
    finishRunDirect(slot);
    const RunHdr *mine = slot.checked ? slot.runBuf->hdr() : nullptr;
    int64_t total = mine ? mine->nTotal : 0;

    if (!root) {
        if (slot.checked) {
            // with the records: rank 0 bumps the activities of every rank's hits (on its device).  Without (1.7 MB per
            // step and rank less to write into host memory, 12 of the 32 bytes per hit): every rank bumps the activities
            // of its own hits on its own device.
            const bool exported = slot.runBuf->withRecords;
            const_cast<RunHdr *>(mine)->hasRecords = exported ? 1u : 0u;
            if (!exported && mine->nTotal > 0) {
                std::vector<DevicePart> own{DevicePart{this, &slot, nullptr}};
                bumpDirect(own);
            }
        }
        // publish: this rank's event has completed, so everything its GPU wrote is visible to rank 0's CPU
        RingCtl *c = P.ring->ctl();
        const uint64_t word = ((uint64_t)P.seq << 8) | (uint64_t)(slot.checked ? P.ringBuf : 0xFF);
        P.ringBuf = -1;
        slot.runBuf.reset(); // (rank 0 owns the buffer from here on)
        c->published.store(word, std::memory_order_release);
    }

    // Once the batch cannot be run again, its tables are no longer needed: a worker collapses every slot
    // to the solver's last one right away (dSetAllAssigsToLast; the GPU is warm and otherwise idle);
    // rank 0 does it at the start of the next batch, behind the push (see peerEnqueue).
    float msCollapse = 0;
    if (root) {
        collapseSlot_ = P.collapsePending ? cur_ : -1;
    } else if (P.collapsePending) {
        GSS_CUDA(cudaEventRecord(P.evC0, stream_));
        launchCollapse(slot.updDev.data(), slot.paramsDev(), slot.nSolvers, slot.maxUpd, slot.nUpdates, tables_, numSMs_, stream_, &launches_);
        GSS_CUDA(cudaEventRecord(P.evC1, stream_));
        collapseSlot_ = -1;
        lastStarted_ = -1; // the tables of this batch are gone
        GSS_CUDA(cudaEventSynchronize(P.evC1));
        cudaEventElapsedTime(&msCollapse, P.evC0, P.evC1);
    }
    P.collapsePending = false;
    // [0] prepare + H2D (+ wait for the batch on a worker), [1] table kernels (push on rank 0, apply, collapse),
    // [2] check kernels + emit, [3] everything from the start to "this rank's results are in host memory" plus
    // the collapse: [3] - [0] = device time of the batch on this rank
    lastTimes_[0] = msFirst[0] * 1000.0;
    lastTimes_[1] = (msFirst[1] + msCollapse) * 1000.0;
    lastTimes_[2] = msFirst[2] * 1000.0;
    lastTimes_[3] = (msFirst[3] + msCollapse) * 1000.0;
    haveTimes_ = true;
    if (P.trace)
        fprintf(stderr, "peer trace rank %d batch %u: copies %.1f | tables %.1f | check+emit %.1f | collapse %.1f us\n", P.rank, P.seq,
                msFirst[0] * 1e3, msFirst[1] * 1e3, msFirst[2] * 1e3, msCollapse * 1e3);
    slot.inFlight = false;
    mgpuLast_ = cur_;
    cur_ = -1;

    if (!root) return total;

    // rank 0: every rank's finished result, one slice per solver and rank
    std::vector<DevicePart> parts{DevicePart{this, &slot, nullptr}};
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 1; r < P.world; r++) {
        PhaseTimer tWorkers(hostPhases_[5]);
        std::shared_ptr<ShmRing> ring = P.rings[r];
        RingCtl *c = ring->ctl();
        uint64_t word;
        for (;;) {
            word = c->published.load(std::memory_order_acquire);
            if ((word >> 8) == (uint64_t)P.seq) break;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > P.timeoutS)
                GSS_DIE("peer exchange: rank " + std::to_string(r) + " did not publish batch " + std::to_string(P.seq));
            std::this_thread::yield();
        }
        const int k = (int)(word & 0xFF);
        if (k == 0xFF) continue; // that rank had nothing to check
        RunBuf *b = new RunBuf();
        b->base = ring->buf(k);
        b->bytes = (size_t)c->bufBytes;
        b->entryCap = c->entryCap;
        b->litCap = c->litCap;
        b->withRecords = b->hdr()->hasRecords != 0;
        std::shared_ptr<RunBuf> buf(b, [ring, k](RunBuf *q) {
            ring->ctl()->buf[k].state.store(0, std::memory_order_release); // the worker may fill it again
            delete q;
        });
        total += b->hdr()->nTotal;
        globalStats_[G_clauseTestsOnAssigs] += b->hdr()->exactTests;
        finishedD2H_ += (int64_t)(sizeof(RunHdr) + (size_t)b->hdr()->nTotal * (b->withRecords ? 24 : 12) + (size_t)b->hdr()->litTotal * 4);
        parts.push_back(DevicePart{nullptr, nullptr, buf});
    }
    PhaseTimer tHandOver(hostPhases_[2]);
    processResultsParts(slot, parts);
    return total;
}

} // namespace gss
