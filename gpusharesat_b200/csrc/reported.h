// reported.h -- host side hand-over of hits to the solver threads (reference Reported /
// ClauseBatch / ConcurrentQueue, gpuShareLib/Reported.{cuh,cu}, ConcurrentQueue.h; rules
// restated in SURVEY.md Appendix B).  Pure CPU logic; observable behaviour follows the
// reference call for call, including its re-report suppression rules -- with ONE deliberate
// deviation that is the default: a clause that was already handed over is skipped and the rest of
// its batch is still delivered, where the reference falls through and discards the rest of the
// batch (Reported.cu:113-129).  GPUSHARE_REFERENCE_DUP_QUIRK=1 restores the reference behaviour;
// tests/test_cpu_host_logic.py covers both settings.
#pragma once
#include "assigs.h"
#include "clause_db.h"
#include "common.h"
#include "pool.h"
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <vector>

namespace gss {

// One solver's share of a run's result exactly as the GPU wrote it into a result buffer in pinned
// host memory (k_emit): n clause ids, n + 1 literal positions, the literal stream.  A ClauseBatch
// built from views copies nothing; `owner` keeps the buffer alive until the batch is recycled.
struct ResultView {
    const int64_t *ids = nullptr;
    const int32_t *pos = nullptr; // n + 1 entries, relative to lits
    int32_t *lits = nullptr;      // writable: callers permute the literals they are handed in place
    int32_t n = 0;
    int32_t cur = 0;              // next entry to hand over (consumer side, ClauseBatch::pop)
    std::shared_ptr<void> owner;
};

// The clauses reported to one solver by one GPU run.
struct ClauseBatch {
    struct Entry {
        int64_t id;
        int32_t pos; // start in lits
    };
    std::vector<int> lits;
    std::vector<Entry> entries;
    // zero-copy part, after `entries`: one view per device that contributed, in device order.  Every view is
    // in the canonical order (length, index) and a device's share of every length array lies before the next
    // device's, so merging the views by clause length (ties: the earlier device) hands the clauses over in
    // exactly the order of a single-device run.
    std::vector<ResultView> views;
    size_t next = 0;
    AssigIds ids;
    uint32_t hadSomeReported = 0;
    int64_t assigWhichKnowsAboutThese = 0;

    void clear() {
        lits.clear();
        entries.clear();
        views.clear(); // drops the references to the result buffers
        next = 0;
        hadSomeReported = 0;
    }
    bool pop(int *&outLits, int &count, int64_t &id) {
        if (next < entries.size()) {
            const Entry &e = entries[next];
            int end = next + 1 < entries.size() ? entries[next + 1].pos : (int)lits.size();
            outLits = lits.data() + e.pos;
            count = end - e.pos;
            id = e.id;
            next++;
            return true;
        }
        ResultView *best = nullptr;
        int bestLen = 0;
        for (ResultView &v : views) {
            if (v.cur >= v.n) continue;
            const int len = v.pos[v.cur + 1] - v.pos[v.cur];
            if (!best || len < bestLen) {
                best = &v;
                bestLen = len;
            }
        }
        if (!best) return false;
        outLits = best->lits + best->pos[best->cur];
        count = bestLen;
        id = best->ids[best->cur];
        best->cur++;
        return true;
    }
    template <typename F> void forEachId(F f) const {
        for (const auto &e : entries) f(e.id);
        for (const auto &v : views)
            for (int32_t i = 0; i < v.n; i++) f(v.ids[i]);
    }
};

// Batches of one solver.  Producer: the GPU thread (begin/publish).  Consumer: that solver's
// thread (takeNext / oldest / retireOldest).  Three cursors like the reference's
// ConcurrentQueue: retired < handed-over < published.
class BatchQueue {
public:
    ClauseBatch &begin();  // a cleared batch, not yet visible to the consumer
    void publish();        // make the batch returned by begin() visible
    bool takeNext(ClauseBatch *&b);
    bool oldest(ClauseBatch *&b);
    void retireOldest();

private:
    std::mutex lock_;
    std::deque<std::unique_ptr<ClauseBatch>> live_; // [retired .. published (+1 being built))
    std::vector<std::unique_ptr<ClauseBatch>> spare_;
    size_t taken_ = 0;     // index in live_ of the next batch to hand over
    size_t published_ = 0; // number of visible batches in live_
};

class Reported {
public:
    Reported(ClauseDb &db, std::vector<std::vector<uint64_t>> &oneSolverStats)
        : db_(db), stats_(oneSolverStats), referenceDupQuirk_(getenv("GPUSHARE_REFERENCE_DUP_QUIRK") != nullptr) {}
    void setSolverCount(int n);
    // false: the caller bumps the activities itself (on the device); true (default): per hit on the host
    void setHostBumps(bool on) { hostBumps_ = on; }
    void setPool(LazyPool *p) { pool_ = p; }

    void clauseWasAdded(int solver, int64_t clauseId);                 // Reported.cu:97-103
    void assigWasSent(int solver, int64_t id) { lastSent_[solver] = id; } // Reported.cuh:118
    // GPU thread: Reported.cu:160-204
    void fill(const std::vector<AssigIds> &ids, const HitRecord *hits, size_t nHits);
    // Same result as fill() for hits already grouped by solver (bucket s = hits[start[s]..start[s+1]),
    // each bucket in hand-over order); the buckets are filled concurrently through `forEach`,
    // which must call its argument once for every solver index (any thread, any order).
    // bump(len, idx) is invoked once per hit from the filling thread.
    void fillBuckets(const std::vector<AssigIds> &ids, const HitRecord *hits, const std::vector<size_t> &start,
                     const std::function<void(const std::function<void(int)> &)> &forEach,
                     const std::function<void(int, int)> &bump);
    // GPU thread: everything that happens to the hits of a finished run -- sort (reproducible
    // hand-over order), activity bumps (Clauses.cu:231-237), batches.  Hit lists of kParallelHits or
    // more are grouped by solver and processed per solver on a small worker pool; `hits` is
    // reordered in place.
    void handOver(std::vector<HitRecord> &hits, const std::vector<AssigIds> &ids, int nSolvers);
    static constexpr size_t kParallelHits = 8192;
    // Same for a hit list the GPU has already put in hand-over order and resolved (ids, literal
    // stream): every solver's batch is one contiguous slice -- sequential copies only.
    void handOverSorted(const SortedHit *recs, size_t n, const int32_t *lits, int64_t totalLits,
                        const std::vector<AssigIds> &ids, int nSolvers);
    // Zero-copy hand-over of a run whose per-solver results the GPU(s) wrote into result buffers in
    // pinned host memory: views[s] = solver s's slices (one per device, possibly empty).  Every solver
    // with assignments in the run gets a batch, hits or not (progress marker, Reported.cu:166-174).
    void handOverViews(std::vector<std::vector<ResultView>> &views, const std::vector<AssigIds> &ids, int nSolvers);
    // solver thread: Reported.cu:105-158
    bool pop(int solver, int *&lits, int &count, int64_t &id);
    int64_t lastAssigAllReported(int solver) const { return lastAllReported_[solver]; }

private:
    struct DontImport {
        int64_t clauseId;
        int64_t assigId;
    };
    ClauseDb &db_;
    std::vector<std::vector<uint64_t>> &stats_;
    std::vector<std::unique_ptr<BatchQueue>> queues_;
    std::vector<std::set<int64_t>> notAgain_; // solver-thread private
    std::vector<ClauseBatch *> current_;      // solver-thread private
    std::vector<int64_t> lastSent_;
    std::vector<int64_t> lastAllReported_;
    std::vector<std::queue<DontImport>> dontImport_;
    std::vector<int> tmpLits_;
    bool referenceDupQuirk_;
    bool hostBumps_ = true;
    std::vector<HitRecord> grouped_;   // scratch: hits grouped by solver
    LazyPool ownPool_;
    LazyPool *pool_ = &ownPool_; // worker pool for large hit lists (the engine shares its own)
};

} // namespace gss
