// reported.cc -- see reported.h
#include "reported.h"
#include "stats.h"

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
            // NOTE: like the reference, a duplicate ends this batch (control reaches the
            // end-of-batch block below); the remaining clauses of the batch are not delivered.
        }
        cur->assigWhichKnowsAboutThese = lastSent_[s] + 1;
        int64_t seenAllReportsUntil = cur->ids.start + cur->ids.count;
        // once every assignment that could not know about a batch's clauses has been fully
        // reported, those clauses may be reported again (the solver may have deleted them)
        ClauseBatch *old;
        while (queues_[s]->oldest(old) && old->assigWhichKnowsAboutThese <= seenAllReportsUntil) {
            for (const auto &e : old->entries) notAgain_[s].erase(e.id);
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
