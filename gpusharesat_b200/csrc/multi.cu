// multi.cu -- several devices behind ONE sharer in ONE process (GPUSHARE_DEVICES=N).
//
// New functionality: the reference drives device 0 only (GpuClauseSharerImpl.cu:52).  The factory
// keeps its signature (GpuClauseSharer.h:165); the device count comes from the environment, so
// glucose-syrup and rel-newtech shard their clause database over the GPUs of a box unchanged.
//
// The sharer the solver threads talk to is the front-end and drives device 0; one worker sharer +
// one host thread + one stream per further device.  Every device keeps the whole clause database
// (host mirror, arenas) and CHECKS its contiguous share of the tiles of every length.  Per run:
//   * the front-end collects the batch (buffer swap per solver, assigs.h) -- the deltas stay in the
//     solver threads' page-locked buffers and EVERY device's k_apply_direct reads them from there over
//     its own PCIe link, in parallel: no root-GPU fan-out, no staging, no collective;
//   * every device runs k_filter / k_exact on its tiles and k_emit writes that device's finished
//     per-solver results into a page-locked result buffer of its own;
//   * the front-end hands each solver a ClauseBatch that views one slice per device.
// Clause activities are authoritative on device 0: its bump kernel reads the other devices' sorted
// record lists in place through peer access (NVLink); reduceDb hands the activities to the workers'
// mirrors so that every device compacts identically.
#include "sharer.h"
#include <atomic>
#include <condition_variable>
#include <thread>

namespace gss {

// One thread per worker device; a command is a function run on every worker thread at once.
class Sharer::WorkerThreads {
public:
    explicit WorkerThreads(int n) : n_(n) {
        for (int i = 0; i < n; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerThreads() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_.fetch_add(1);
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void start(const std::function<void(int)> &fn) {
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            pending_.store(n_);
            gen_.fetch_add(1);
        }
        cv_.notify_all();
    }
    void wait() {
        // the commands are short (enqueue a run / wait for a device): spin briefly, then block
        for (int i = 0; i < 20000 && pending_.load(std::memory_order_acquire) != 0; i++) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        std::unique_lock<std::mutex> lk(m_);
        doneCv_.wait(lk, [this] { return pending_.load() == 0; });
        fn_ = nullptr;
    }
    void run(const std::function<void(int)> &fn) {
        start(fn);
        wait();
    }

private:
    void loop(int i) {
        uint64_t seen = 0;
        for (;;) {
            // a run follows the previous one within microseconds when the solvers are busy: spin first
            for (int k = 0; k < 20000 && gen_.load(std::memory_order_acquire) == seen; k++) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            const std::function<void(int)> *fn;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_.load() != seen; });
                seen = gen_.load();
                if (stop_) return;
                fn = fn_;
            }
            if (fn) (*fn)(i);
            if (pending_.fetch_sub(1) == 1) {
                std::lock_guard<std::mutex> g(m_);
                doneCv_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, doneCv_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> pending_{0};
    std::atomic<uint64_t> gen_{0};
    bool stop_ = false;
};

void Sharer::multiInit(int nDevices) {
    int visible = 0;
    GSS_CUDA(cudaGetDeviceCount(&visible));
    if (nDevices > visible)
        GSS_DIE("GPUSHARE_DEVICES=" + std::to_string(nDevices) + " but only " + std::to_string(visible) + " CUDA devices are visible");
    if (!directEnabled_) GSS_DIE("GPUSHARE_DEVICES needs the direct pipeline (unset GPUSHARE_LEGACY_PIPELINE)");
    db_->setShard(0, nDevices);
    for (int r = 1; r < nDevices; r++) {
        const int dev = (device_ + r) % visible;
        int can = 0;
        GSS_CUDA(cudaDeviceCanAccessPeer(&can, device_, dev));
        if (!can) GSS_DIE("GPUSHARE_DEVICES: device " + std::to_string(device_) + " cannot access device " + std::to_string(dev) + " as a peer");
        useDevice();
        cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) GSS_CUDA(e);
        cudaGetLastError();
        gss_options o = opts_;
        o.verbosity = 0;
        workers_.push_back(std::unique_ptr<Sharer>(new Sharer(o, nullptr, nullptr, dev)));
        Sharer &w = *workers_.back();
        w.root_ = this;
        w.db_->setShard(r, nDevices);
        w.db_->setMaxLen(db_->maxLen());
    }
    useDevice();
    GSS_CUDA(cudaEventCreateWithFlags(&peerReadEv_, cudaEventDisableTiming));
    wthreads_ = new WorkerThreads(nDevices - 1);
    logger_.log(1, "c gpushare_b200: clause check sharded over " + std::to_string(nDevices) + " devices of this process\n");
}

void Sharer::multiShutdown() {
    if (workers_.empty()) return;
    delete wthreads_;
    wthreads_ = nullptr;
    workers_.clear();
    useDevice();
    if (peerReadEv_) cudaEventDestroy(peerReadEv_);
    peerReadEv_ = nullptr;
}

// worker: enqueue its share of the run the front-end has collected into rootSlot
void Sharer::workerStart(const RunSlot &rootSlot, const uint8_t *paramsAndSrc, int slotIdx, bool rebuild) {
    useDevice();
    db_->drainPending();
    RunSlot &slot = slots_[slotIdx];
    bool myRebuild = false;
    int64_t h2d = 0;
    if (!prepareRun(slot, myRebuild, h2d)) GSS_DIE("out of device memory on a worker device (GPUSHARE_DEVICES)");
    if (myRebuild != rebuild) GSS_DIE("the devices of a multi-device sharer disagree about a table rebuild");
    GSS_CHECK(slot.nSolvers == rootSlot.nSolvers);
    const int S = slot.nSolvers;
    slot.srcOff = (slot.dirBytes + (size_t)S * sizeof(SolverRunParams) + 15) / 16 * 16;
    slot.headHost.resize(slot.srcOff + (size_t)S * sizeof(void *));
    // (a snapshot: the front-end patches its own pointer array while it launches)
    memcpy(slot.headHost.data() + slot.dirBytes, paramsAndSrc, (size_t)S * sizeof(SolverRunParams));
    memcpy(slot.headHost.data() + slot.srcOff, paramsAndSrc + (size_t)S * sizeof(SolverRunParams), (size_t)S * sizeof(void *));
    slot.aggStart = rootSlot.aggStart;
    slot.aggOnDevice = false;
    slot.maxUpd = rootSlot.maxUpd;
    slot.nUpdates = rootSlot.nUpdates;
    slot.staged = rootSlot.staged; // (host copies in the front-end's staging buffer)
    slot.ids.assign(S, AssigIds{});
    slot.assigCount = 0;
    // the front-end's bump kernels read this slot's record lists of two runs ago in place
    if (root_->peerReadRecorded_) GSS_CUDA(cudaStreamWaitEvent(stream_, root_->peerReadEv_, 0));
    launchDirect(slot, h2d);
    cur_ = slotIdx;
}

void Sharer::workerFinish(int slotIdx) {
    useDevice();
    GSS_CHECK(cur_ == slotIdx);
    finishRunDirect(slots_[slotIdx]);
    cur_ = -1;
}

// the activities are authoritative on this device: every mirror takes them, then every device
// compacts its copy of the database the same way, in parallel
void Sharer::reduceDbMulti() {
    db_->syncActivitiesFromDevice(stream_);
    // (two steps: nobody may still be reading this device's activities when its own reduction rewrites them)
    // (every device drains its pending clauses at the start of the same runs, never here: the mirrors are in step)
    const std::function<void(int)> take = [&](int i) {
        Sharer &w = *workers_[i];
        w.useDevice();
        w.db_->copyActivitiesFrom(*db_);
    };
    wthreads_->start(take);
    wthreads_->wait();
    const std::function<void(int)> red = [&](int i) {
        Sharer &w = *workers_[i];
        w.useDevice();
        w.db_->reduceAfterSync(w.stream_);
        w.lastStarted_ = -1;
    };
    wthreads_->start(red);
    db_->reduceAfterSync(stream_);
    wthreads_->wait();
}

void Sharer::wholeRunMulti(bool canStart) {
    useDevice();
    const int prevIdx = cur_;
    RunSlot *prev = prevIdx >= 0 ? &slots_[prevIdx] : nullptr;
    if (prev) {
        PhaseTimer t(hostPhases_[0]);
        const std::function<void(int)> fin = [&](int w) { workers_[w]->workerFinish(prevIdx); };
        wthreads_->start(fin);
        finishRunDirect(*prev);
        wthreads_->wait();
        // device time of the run = the slowest device's
        for (auto &w : workers_)
            for (int i = 0; i < 4; i++) lastTimes_[i] = std::max(lastTimes_[i], w->lastTimes_[i]);
        for (auto &w : workers_) finishedD2H_ += w->finishedD2H_;
    }
    int started = -1;
    if (canStart) {
        const int next = prevIdx >= 0 ? 1 - prevIdx : (collapseSlot_ >= 0 ? 1 - collapseSlot_ : 0);
        PhaseTimer t(hostPhases_[1]);
        db_->drainPending();
        if (db_->stats().clauses > 0) {
            RunSlot &slot = slots_[next];
            int64_t h2d = 0;
            bool rebuild = false;
            if (!prepareRun(slot, rebuild, h2d)) GSS_DIE("out of device memory (a multi-device sharer does not reduce the database by itself)");
            collectDirect(slot, rebuild);
            const size_t S = (size_t)slot.nSolvers;
            multiSnap_.resize(S * (sizeof(SolverRunParams) + sizeof(void *)));
            memcpy(multiSnap_.data(), slot.headHost.data() + slot.dirBytes, S * sizeof(SolverRunParams));
            memcpy(multiSnap_.data() + S * sizeof(SolverRunParams), slot.headHost.data() + slot.srcOff, S * sizeof(void *));
            const std::function<void(int)> go = [&](int w) { workers_[w]->workerStart(slot, multiSnap_.data(), next, rebuild); };
            wthreads_->start(go);
            launchDirect(slot, h2d);
            wthreads_->wait();
            for (auto &w : workers_) lastH2D_ += w->lastH2D_;
            started = next;
        }
    }
    if (prev) {
        PhaseTimer t(hostPhases_[2]);
        std::vector<DevicePart> parts{DevicePart{this, prev, nullptr}};
        for (auto &w : workers_) parts.push_back(DevicePart{w.get(), &w->slots_[prevIdx], nullptr});
        processResultsParts(*prev, parts);
        GSS_CUDA(cudaEventRecord(peerReadEv_, stream_));
        peerReadRecorded_ = true;
    }
    cur_ = started;
}

} // namespace gss
