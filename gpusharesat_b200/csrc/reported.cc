// reported.cc -- see reported.h
#include "reported.h"
#include "stats.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>

namespace gss {

ClauseBatch &BatchQueue::begin() {
    std::lock_guard<std::mutex> g(lock_);
    GSS_CHECK(published_ == live_.size());
    std::unique_ptr<ClauseBatch> b;
    if (!spare_.empty()) {
        b = std::move(spare_.back());
        spare_.pop_back();
    } else {
        b = std::make_unique<ClauseBatch>();
    }
    b->clear();
    live_.push_back(std::move(b));
    return *live_.back();
}

void BatchQueue::publish() {
    std::lock_guard<std::mutex> g(lock_);
    GSS_CHECK(published_ + 1 == live_.size());
    published_++;
}

bool BatchQueue::takeNext(ClauseBatch *&b) {
    std::lock_guard<std::mutex> g(lock_);
    if (taken_ == published_) return false;
    b = live_[taken_++].get();
    return true;
}

bool BatchQueue::oldest(ClauseBatch *&b) {
    std::lock_guard<std::mutex> g(lock_);
    if (taken_ == 0) return false; // nothing handed over that is not yet retired
    b = live_.front().get();
    return true;
}

void BatchQueue::retireOldest() {
    std::lock_guard<std::mutex> g(lock_);
    GSS_CHECK(taken_ > 0);
    live_.front()->views.clear(); // give the result buffers back now, not when the batch object is reused
    spare_.push_back(std::move(live_.front()));
    live_.pop_front();
    taken_--;
    published_--;
}

void Reported::setSolverCount(int n) {
    size_t old = queues_.size();
    if ((size_t)n <= old) return;
    queues_.resize(n);
    for (size_t s = old; s < (size_t)n; s++) queues_[s] = std::make_unique<BatchQueue>();
    notAgain_.resize(n);
    current_.resize(n, nullptr);
    lastSent_.resize(n, 0);
    lastAllReported_.resize(n, 0);
    dontImport_.resize(n);
}

void Reported::clauseWasAdded(int solver, int64_t clauseId) {
    // the exporter must not get its own clause back until an assignment that knows it was sent
    notAgain_[solver].insert(clauseId);
    dontImport_[solver].push(DontImport{clauseId, lastSent_[solver] + 1});
}

void Reported::fill(const std::vector<AssigIds> &ids, const HitRecord *hits, size_t nHits) {
    // every solver that had assignments in the run gets a batch, hits or not: the batch is
    // also the progress marker for getLastAssigAllReported (Reported.cu:166-174)
    std::vector<ClauseBatch *> perSolver(queues_.size(), nullptr);
    auto batchOf = [&](int s) -> ClauseBatch & {
        if (!perSolver[s]) perSolver[s] = &queues_[s]->begin();
        return *perSolver[s];
    };
    for (size_t s = 0; s < ids.size() && s < queues_.size(); s++)
        if (ids[s].count > 0) batchOf((int)s).ids = ids[s];
    for (size_t i = 0; i < nHits; i++) {
        const HitRecord &h = hits[i];
        if (i + 16 < nHits) db_.prefetchClause(hits[i + 16].len, hits[i + 16].idx);
        ClauseBatch &b = batchOf(h.solver);
        int32_t pos = (int32_t)b.lits.size();
        int64_t id = db_.appendClause(h.len, h.idx, b.lits);
        b.entries.push_back(ClauseBatch::Entry{id, pos});
        b.hadSomeReported |= h.mask;
    }
    for (size_t s = 0; s < queues_.size(); s++)
        if (perSolver[s]) queues_[s]->publish();
}

void Reported::fillBuckets(const std::vector<AssigIds> &ids, const HitRecord *hits, const std::vector<size_t> &start,
                           const std::function<void(const std::function<void(int)> &)> &forEach,
                           const std::function<void(int, int)> &bump) {
    size_t nSolvers = queues_.size();
    std::vector<ClauseBatch *> perSolver(nSolvers, nullptr);
    for (size_t s = 0; s < nSolvers; s++) {
        bool hasIds = s < ids.size() && ids[s].count > 0;
        bool hasHits = s + 1 < start.size() && start[s + 1] > start[s];
        if (!hasIds && !hasHits) continue;
        perSolver[s] = &queues_[s]->begin();
        if (hasIds) perSolver[s]->ids = ids[s];
    }
    forEach([&](int s) {
        if ((size_t)s >= nSolvers || !perSolver[s] || (size_t)s + 1 >= start.size()) return;
        ClauseBatch &b = *perSolver[s];
        size_t lo = start[s], hi = start[s + 1];
        b.entries.reserve(hi - lo);
        for (size_t i = lo; i < hi; i++) {
            const HitRecord &h = hits[i];
            if (i + 16 < hi) db_.prefetchClause(hits[i + 16].len, hits[i + 16].idx);
            int32_t pos = (int32_t)b.lits.size();
            int64_t id = db_.appendClause(h.len, h.idx, b.lits);
            b.entries.push_back(ClauseBatch::Entry{id, pos});
            b.hadSomeReported |= h.mask;
            bump(h.len, h.idx);
        }
    });
    for (size_t s = 0; s < nSolvers; s++)
        if (perSolver[s]) queues_[s]->publish();
}

void Reported::handOver(std::vector<HitRecord> &hits, const std::vector<AssigIds> &ids, int nSolvers) {
    // The kernels append hits in scheduling order.  Sorting them makes everything downstream
    // (activity bumps, batch order, hence which clause a solver sees first) reproducible;
    // the reference hands them over in whatever order the atomics produced.
    auto byClause = [](const HitRecord &a, const HitRecord &b) {
        if (a.len != b.len) return a.len < b.len;
        if (a.idx != b.idx) return a.idx < b.idx;
        return a.solver < b.solver;
    };
    if (hits.size() < kParallelHits) {
        std::sort(hits.begin(), hits.end(), byClause);
        for (size_t i = 0; i < hits.size(); i++) {
            if (!hostBumps_) break;
            if (i + 16 < hits.size()) db_.prefetchClause(hits[i + 16].len, hits[i + 16].idx);
            db_.bumpActivity(hits[i].len, hits[i].idx);
        }
        fill(ids, hits.data(), hits.size());
        return;
    }
    // Large hit list: the per-solver batches are independent.  Group the hits by solver (one
    // counting pass), then sort / bump / copy literals per solver on the worker pool.  Every
    // solver's batch ends up in the same (len, idx) order as on the serial path.
    static const bool prof = getenv("GSS_PROFILE_HANDOVER") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::micro>(b - a).count();
    };
    auto t0 = now();
    std::vector<size_t> start(nSolvers + 1, 0);
    for (const HitRecord &h : hits) start[h.solver + 1]++;
    for (int s = 0; s < nSolvers; s++) start[s + 1] += start[s];
    grouped_.resize(hits.size());
    {
        std::vector<size_t> cursor(start.begin(), start.end() - 1);
        for (const HitRecord &h : hits) grouped_[cursor[h.solver]++] = h;
    }
    hits.swap(grouped_);
    auto t1 = now();
    std::atomic<bool> rescale{false};
    fillBuckets(
        ids, hits.data(), start,
        [&](const std::function<void(int)> &perSolver) {
            pool_->get().parallelFor(nSolvers, [&](int s) {
                std::sort(hits.begin() + start[s], hits.begin() + start[s + 1], byClause);
                perSolver(s);
            });
        },
        [&](int len, int idx) {
            if (hostBumps_ && db_.bumpActivityAtomic(len, idx)) rescale.store(true);
        });
    db_.rescaleIfNeeded(rescale.load());
    if (prof) fprintf(stderr, "handOver: group %.0f us, sort+fill %.0f us (%zu hits)\n", us(t0, t1), us(t1, now()), hits.size());
}

void Reported::handOverSorted(const SortedHit *recs, size_t n, const int32_t *lits, int64_t totalLits,
                              const std::vector<AssigIds> &ids, int nSolvers) {
    static const bool prof = getenv("GSS_PROFILE_HANDOVER") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::micro>(b - a).count();
    };
    const auto t0 = now();
    // recs are ordered by (solver, length, index): find every solver's slice
    std::vector<size_t> start((size_t)nSolvers + 1, n);
    for (int s = 0; s < nSolvers; s++)
        start[s] = (size_t)(std::lower_bound(recs, recs + n, s, [](const SortedHit &h, int v) { return h.solver < v; }) - recs);
    start[nSolvers] = n;
    size_t nQueues = queues_.size();
    std::vector<ClauseBatch *> perSolver(nQueues, nullptr);
    for (size_t s = 0; s < nQueues && s < (size_t)nSolvers; s++) {
        bool hasIds = s < ids.size() && ids[s].count > 0;
        if (!hasIds && start[s + 1] == start[s]) continue;
        perSolver[s] = &queues_[s]->begin();
        if (hasIds) perSolver[s]->ids = ids[s];
    }
    std::atomic<bool> rescale{false};
    const auto t1 = now();
    pool_->get().parallelFor(nSolvers, [&](int s) {
        if ((size_t)s >= nQueues || !perSolver[s]) return;
        size_t lo = start[s], hi = start[s + 1];
        if (lo == hi) return;
        ClauseBatch &b = *perSolver[s];
        const int64_t base = recs[lo].litPos;
        const int64_t end = hi < n ? recs[hi].litPos : totalLits;
        b.lits.assign(lits + base, lits + end);
        b.entries.resize(hi - lo);
        uint32_t any = 0;
        for (size_t i = lo; i < hi; i++) {
            const SortedHit &h = recs[i];
            b.entries[i - lo] = ClauseBatch::Entry{h.id, (int32_t)(h.litPos - base)};
            any |= h.mask;
            if (hostBumps_) {
                if (i + 16 < hi) db_.prefetchMeta(recs[i + 16].len, recs[i + 16].idx);
                if (db_.bumpActivityAtomic(h.len, h.idx)) rescale.store(true);
            }
        }
        b.hadSomeReported |= any;
    });
    db_.rescaleIfNeeded(rescale.load());
    const auto t2 = now();
    for (size_t s = 0; s < nQueues; s++)
        if (perSolver[s]) queues_[s]->publish();
    if (prof) fprintf(stderr, "handOverSorted: slices + begin %.0f us, fill %.0f us, publish %.0f us (%zu hits)\n", us(t0, t1), us(t1, t2), us(t2, now()), n);
}

void Reported::handOverViews(std::vector<std::vector<ResultView>> &views, const std::vector<AssigIds> &ids, int nSolvers) {
    const size_t nQueues = queues_.size();
    for (size_t s = 0; s < nQueues && s < (size_t)nSolvers; s++) {
        const bool hasIds = s < ids.size() && ids[s].count > 0;
        bool hasHits = false;
        if (s < views.size())
            for (const ResultView &v : views[s]) hasHits = hasHits || v.n > 0;
        if (!hasIds && !hasHits) continue;
        ClauseBatch &b = queues_[s]->begin();
        if (hasIds) b.ids = ids[s];
        if (hasHits)
            for (ResultView &v : views[s])
                if (v.n > 0) b.views.push_back(std::move(v));
        queues_[s]->publish();
    }
}

bool Reported::pop(int s, int *&lits, int &count, int64_t &id) {
    while (true) {
        if (!current_[s]) queues_[s]->takeNext(current_[s]);
        ClauseBatch *cur = current_[s];
        if (!cur) return false;
        if (cur->pop(lits, count, id)) {
            // a clause may fire for several assignments that could not know about it yet;
            // hand it over once (Reported.cu:113-129)
            if (notAgain_[s].find(id) == notAgain_[s].end()) {
                notAgain_[s].insert(id);
                stats_[s][S_reportedClauses]++;
                if (count == 1) stats_[s][S_reportedClausesUnit]++;
                if (count == 2) stats_[s][S_reportedClausesBinary]++;
                return true;
            }
            // Already handed over: skip it and look at the next clause of the batch.  In the
            // reference control falls through to the end-of-batch block here, so one re-reported
            // clause silently discards the REST of its batch (Reported.cu:113-129) -- with a
            // reproducible hand-over order that would starve newly added clauses for good.
            // GPUSHARE_REFERENCE_DUP_QUIRK=1 restores the reference behaviour (parity tests).
            if (!referenceDupQuirk_) continue;
        }
        cur->assigWhichKnowsAboutThese = lastSent_[s] + 1;
        int64_t seenAllReportsUntil = cur->ids.start + cur->ids.count;
        // once every assignment that could not know about a batch's clauses has been fully
        // reported, those clauses may be reported again (the solver may have deleted them)
        ClauseBatch *old;
        while (queues_[s]->oldest(old) && old->assigWhichKnowsAboutThese <= seenAllReportsUntil) {
            old->forEachId([&](int64_t oldId) { notAgain_[s].erase(oldId); });
            queues_[s]->retireOldest();
        }
        auto &q = dontImport_[s];
        while (!q.empty() && q.front().assigId < seenAllReportsUntil) {
            notAgain_[s].erase(q.front().clauseId);
            q.pop();
        }
        lastAllReported_[s] = seenAllReportsUntil;
        current_[s] = nullptr;
    }
}

} // namespace gss
