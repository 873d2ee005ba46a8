// pool.h -- a tiny persistent worker pool for the host side of a run.  Only used when a run
// returns a large hit list (thousands of hits): the per-solver hand-over batches are independent,
// so sorting them and copying their literals out of the host mirror is spread over a few threads.
// Small runs (the common case in a real portfolio) never touch it.
#pragma once
#include <atomic>
#include <condition_variable>
#include <algorithm>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace gss {

class WorkerPool {
public:
    explicit WorkerPool(int nWorkers) {
        for (int i = 0; i < nWorkers; i++) threads_.emplace_back([this] { workerLoop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            generation_++;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    WorkerPool(const WorkerPool &) = delete;
    WorkerPool &operator=(const WorkerPool &) = delete;

    // runs fn(0..nTasks-1) on the workers and the calling thread; returns when all are done
    void parallelFor(int nTasks, const std::function<void(int)> &fn) {
        if (nTasks <= 0) return;
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            nTasks_ = nTasks;
            next_.store(0);
            pending_.store(nTasks);
            generation_++;
        }
        cv_.notify_all();
        runTasks();
        std::unique_lock<std::mutex> lk(m_);
        doneCv_.wait(lk, [this] { return pending_.load() == 0 && active_ == 0; });
        fn_ = nullptr;
    }

private:
    void runTasks() {
        for (;;) {
            int i = next_.fetch_add(1);
            if (i >= nTasks_) break;
            (*fn_)(i);
            pending_.fetch_sub(1);
        }
    }
    void workerLoop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                if (!fn_) continue;
                active_++;
            }
            runTasks();
            {
                std::lock_guard<std::mutex> g(m_);
                active_--;
            }
            doneCv_.notify_all();
        }
    }

    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, doneCv_;
    const std::function<void(int)> *fn_ = nullptr;
    int nTasks_ = 0;
    std::atomic<int> next_{0}, pending_{0};
    int active_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

// created on first use (small runs never need it); shared by the collect and the hand-over side
struct LazyPool {
    std::unique_ptr<WorkerPool> p;
    WorkerPool &get() {
        if (!p) p = std::make_unique<WorkerPool>(std::max(1, std::min(15, (int)std::thread::hardware_concurrency() - 1)));
        return *p;
    }
};

} // namespace gss
