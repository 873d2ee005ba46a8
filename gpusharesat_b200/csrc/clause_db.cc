// clause_db.cc -- see clause_db.h.  Host logic; device memory through the CUDA runtime only.
#include "clause_db.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace gss {

static constexpr float kRescale = 1e19f; // reference RESCALE_CONST, Clauses.cuh:39

ClauseDb::ClauseDb(double activityDecay, const Logger &logger, size_t pinnedLimitBytes)
    : logger_(logger), pinnedLimit_(pinnedLimitBytes), actDecay_((float)activityDecay) {
    setMaxLen(kDefaultMaxClauseLen);
}

void ClauseDb::setMaxLen(int maxLen) {
    if (frozenMaxLen_) GSS_DIE("gss_set_max_clause_len must be called before the first clause is added");
    // the level-1 survivor record packs the clause length into 16 bits (Survivor::ptrLen)
    if (maxLen < 1 || maxLen > kMaxSupportedClauseLen)
        GSS_DIE("maximum clause length must be in [1, " + std::to_string(kMaxSupportedClauseLen) + "]");
    maxLen_ = maxLen;
    perLen_.clear();
    perLen_.resize(maxLen_ + 1);
    for (int s = 0; s <= maxLen_; s++) {
        perLen_[s] = std::make_unique<PerLen>();
        perLen_[s]->lits.setPinnedLimit(pinnedLimit_);
    }
}

void ClauseDb::setShard(int rank, int world) {
    if (frozenMaxLen_) GSS_DIE("the shard must be set before the first clause is added");
    GSS_CHECK(world >= 1 && rank >= 0 && rank < world);
    shardRank_ = rank;
    shardWorld_ = world;
}

int64_t ClauseDb::addClause(const int *lits, int n) {
    // Clauses.cu:350: oversize clauses are rejected before an id is taken.  (The reference
    // also accepts n == 0 into a never-scanned array; an empty clause carries no literal to
    // test, so it is rejected here as well.)
    if (n > maxLen_ || n < 1) return -1;
    std::lock_guard<std::mutex> g(pendingLock_);
    frozenMaxLen_ = true;
    if (pendingLens_.empty()) pendingFirstId_ = nextId_;
    pendingLits_.insert(pendingLits_.end(), lits, lits + n);
    pendingLens_.push_back(n);
    return nextId_++;
}

int64_t ClauseDb::addClausesBulk(const int64_t *offsets, const int *lits, int64_t nclauses) {
    for (int64_t c = 0; c < nclauses; c++) {
        int64_t n = offsets[c + 1] - offsets[c];
        if (n > maxLen_ || n < 1) return -1;
    }
    std::lock_guard<std::mutex> g(pendingLock_);
    frozenMaxLen_ = true;
    if (pendingLens_.empty()) pendingFirstId_ = nextId_;
    pendingLits_.insert(pendingLits_.end(), lits + offsets[0], lits + offsets[nclauses]);
    pendingLens_.reserve(pendingLens_.size() + nclauses);
    for (int64_t c = 0; c < nclauses; c++) pendingLens_.push_back((int)(offsets[c + 1] - offsets[c]));
    int64_t first = nextId_;
    nextId_ += nclauses;
    return first;
}

void ClauseDb::appendToMirror(const int *lits, int n, int64_t id) {
    PerLen &pl = *perLen_[n];
    int64_t idx = (int64_t)pl.meta.size();
    size_t need = wordsFor(n, idx + 1);
    if (need > pl.lits.size()) pl.lits.resize(need); // new tile, zero-filled
    for (int i = 0; i < n; i++) {
        pl.lits[wordPos(n, idx, i)] = lits[i];
        int v = litVar(lits[i]) + 1;
        if (v > maxVarPlusOne_) maxVarPlusOne_ = v;
    }
    pl.meta.push_back(ClauseMeta{id, actIncr_});
    stats_.clauses++;
    stats_.lengthSum += n;
    stats_.added++;
    if (actIncr_ > kRescale) rescaleActivity();
}

void ClauseDb::drainPending() {
    std::vector<int> lits, lens;
    int64_t firstId;
    {
        std::lock_guard<std::mutex> g(pendingLock_);
        if (pendingLens_.empty()) return;
        lits.swap(pendingLits_);
        lens.swap(pendingLens_);
        firstId = pendingFirstId_;
    }
    std::vector<char> wasEmpty(maxLen_ + 1, 0);
    for (int s = 1; s <= maxLen_; s++) wasEmpty[s] = perLen_[s]->meta.empty();
    size_t pos = 0;
    for (size_t c = 0; c < lens.size(); c++) {
        // Clauses.cu:334 + Clauses.cuh:192: the increment decays once per added clause
        actIncr_ /= actDecay_;
        appendToMirror(&lits[pos], lens[c], firstId + (int64_t)c);
        pos += lens[c];
    }
    // bulk load into an empty arena: nothing refers to these clause indices yet, order them
    for (int s = 1; s <= maxLen_; s++)
        if (wasEmpty[s] && perLen_[s]->meta.size() >= kSortMinClauses) sortArena(s);
}

void ClauseDb::sortArena(int len) {
    PerLen &pl = *perLen_[len];
    const int64_t n = (int64_t)pl.meta.size();
    if (n < 2) return;
    // counting sort by first literal (stable: equal literals keep arrival order)
    int32_t maxLit = 0;
    for (int64_t i = 0; i < n; i++) maxLit = std::max(maxLit, pl.lits[wordPos(len, i, 0)]);
    std::vector<int64_t> bucket((size_t)maxLit + 2, 0);
    for (int64_t i = 0; i < n; i++) bucket[(size_t)pl.lits[wordPos(len, i, 0)] + 1]++;
    for (size_t b = 1; b < bucket.size(); b++) bucket[b] += bucket[b - 1];
    std::vector<int32_t> oldLits(pl.lits.data(), pl.lits.data() + pl.lits.size());
    std::vector<ClauseMeta> oldMeta(pl.meta);
    for (int64_t i = 0; i < n; i++) {
        int64_t to = bucket[(size_t)oldLits[wordPos(len, i, 0)]]++;
        for (int k = 0; k < len; k++) pl.lits[wordPos(len, to, k)] = oldLits[wordPos(len, i, k)];
        pl.meta[to] = oldMeta[i];
    }
    pl.fullReupload = true;
    pl.dirtyFrom = 0;
}

void scaleActivitiesOnDevice(float *acts, int64_t n, float factor, cudaStream_t stream); // kernels.cu

void ClauseDb::applyPendingDeviceRescales(cudaStream_t stream) {
    for (; pendingDeviceRescales_ > 0; pendingDeviceRescales_--)
        for (int s = 1; s <= maxLen_; s++)
            if (perLen_[s]->actsOnDevice > 0)
                scaleActivitiesOnDevice(perLen_[s]->actsDev.data(), perLen_[s]->actsOnDevice, 1.0f / kRescale, stream);
}

bool ClauseDb::uploadDirty(cudaStream_t stream, int64_t *bytesCopied) {
    // rescales decided while draining (Clauses.cu:284-291) reach the device copies first
    applyPendingDeviceRescales(stream);
    for (int s = maxLen_; s >= 1; s--) {
        PerLen &pl = *perLen_[s];
        int64_t n = (int64_t)pl.meta.size();
        if (!pl.fullReupload && pl.dirtyFrom >= n) continue;
        const size_t tileWords = (size_t)kTileClauses * s;
        const int64_t tiles = (n + kTileClauses - 1) / kTileClauses;
        const int64_t firstTile = pl.fullReupload ? 0 : pl.dirtyFrom / kTileClauses;
        {
            // Sharding (multi-GPU) splits the WORK, not the storage: every device holds the whole
            // arena (176 MB at 10 M clauses is nothing next to 180 GB), so any rank can resolve any
            // hit (ids, literals) on its device; a rank only checks its share of the tiles.
            size_t total = (size_t)tiles * tileWords, from = (size_t)firstTile * tileWords;
            if (total > 0) {
                if (!pl.dev.tryReserve(total, pl.fullReupload ? 0 : from, stream)) return false;
                GSS_CUDA(cudaMemcpyAsync(pl.dev.data() + from, pl.lits.data() + from, (total - from) * sizeof(int32_t),
                                         cudaMemcpyHostToDevice, stream));
                if (bytesCopied) *bytesCopied += (int64_t)((total - from) * sizeof(int32_t));
            }
        }
        // clause ids of the same dirty range
        {
            int64_t from = pl.fullReupload ? 0 : pl.dirtyFrom;
            if (n > from) {
                if (!pl.idsDev.tryReserve((size_t)n, (size_t)from, stream)) return false;
                pl.idsStage.setPinnedLimit(pinnedLimit_);
                pl.idsStage.clear();
                pl.idsStage.resize((size_t)(n - from));
                for (int64_t i = from; i < n; i++) pl.idsStage[(size_t)(i - from)] = pl.meta[(size_t)i].id;
                GSS_CUDA(cudaMemcpyAsync(pl.idsDev.data() + from, pl.idsStage.data(), (size_t)(n - from) * sizeof(int64_t),
                                         cudaMemcpyHostToDevice, stream));
                if (bytesCopied) *bytesCopied += (n - from) * (int64_t)sizeof(int64_t);
                if (deviceActs_) {
                    if (!pl.actsDev.tryReserve((size_t)n, (size_t)from, stream)) return false;
                    pl.actsStage.setPinnedLimit(pinnedLimit_);
                    pl.actsStage.clear();
                    pl.actsStage.resize((size_t)(n - from));
                    for (int64_t i = from; i < n; i++) pl.actsStage[(size_t)(i - from)] = pl.meta[(size_t)i].activity;
                    GSS_CUDA(cudaMemcpyAsync(pl.actsDev.data() + from, pl.actsStage.data(), (size_t)(n - from) * sizeof(float),
                                             cudaMemcpyHostToDevice, stream));
                    pl.actsOnDevice = n;
                }
            }
        }
        pl.dirtyFrom = n;
        pl.fullReupload = false;
    }
    return true;
}

int ClauseDb::buildDirectory(std::vector<LenDir> &dir) const {
    dir.clear();
    int tiles = 0;
    for (int s = maxLen_; s >= 1; s--) { // longest first: the tail of the grid gets the cheap tiles
        const PerLen &pl = *perLen_[s];
        int64_t n = (int64_t)pl.meta.size();
        if (n == 0) continue;
        // a length none of whose tiles is checked here still gets its entry: hits of other ranks
        // are resolved (ids, literals, activity bumps) through this directory
        const int64_t allTiles = (n + kTileClauses - 1) / kTileClauses;
        tiles += (int)localTiles(allTiles);
        dir.push_back(LenDir{pl.dev.data(), s, (int32_t)n, tiles, (int32_t)shardFirstTile(allTiles), pl.idsDev.data(),
                             const_cast<float *>(pl.actsDev.data()), 0});
    }
    int64_t before = 0; // the directory is longest first: walk it backwards for the ascending prefix
    for (size_t k = dir.size(); k-- > 0;) {
        dir[k].ascStart = before;
        before += dir[k].count;
    }
    return tiles;
}

void ClauseDb::getClause(int len, int idx, std::vector<int> &lits, int64_t &id) const {
    const PerLen &pl = *perLen_[len];
    lits.resize(len);
    for (int i = 0; i < len; i++) lits[i] = pl.lits[wordPos(len, idx, i)];
    id = pl.meta[idx].id;
}

void ClauseDb::bumpActivity(int len, int idx) {
    float &a = perLen_[len]->meta[idx].activity;
    a += actIncr_;
    if (a > kRescale) rescaleActivity();
}

void ClauseDb::rescaleActivity() { // Clauses.cu:284-291
    for (int s = 0; s <= maxLen_; s++)
        for (auto &m : perLen_[s]->meta) m.activity /= kRescale;
    actIncr_ /= kRescale;
    if (deviceActs_) pendingDeviceRescales_++;
}

void ClauseDb::downloadActivities(cudaStream_t stream) {
    if (!deviceActs_) return;
    std::vector<float> tmp;
    for (int s = 1; s <= maxLen_; s++) {
        PerLen &pl = *perLen_[s];
        int64_t n = std::min<int64_t>(pl.actsOnDevice, (int64_t)pl.meta.size());
        if (n <= 0) continue;
        tmp.resize((size_t)n);
        GSS_CUDA(cudaMemcpyAsync(tmp.data(), pl.actsDev.data(), (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream));
        GSS_CUDA(cudaStreamSynchronize(stream));
        for (int64_t i = 0; i < n; i++) pl.meta[(size_t)i].activity = tmp[(size_t)i];
    }
}

float ClauseDb::approxNthAct(int64_t n) const {
    // Clauses.cu:492-525: 20000 log-scale buckets over the float range, rounded up.  The
    // facade passes the clause length as LBD (GpuClauseSharerImpl.cu:125) and the range is
    // [0, MAX_CL_SIZE), so clauses of exactly the maximum length are not counted.
    if (n == 0) return 0.0f;
    const int buckets = 20000;
    std::vector<int64_t> counts(buckets, 0);
    float lowestLog = std::log(std::numeric_limits<float>::min());
    float largestLog = std::log(std::numeric_limits<float>::max());
    float stepLog = (largestLog - lowestLog) / buckets;
    for (int s = 1; s < maxLen_; s++) {
        for (const auto &m : perLen_[s]->meta) {
            int b = (int)std::floor(((double)std::log(m.activity) - (double)lowestLog) / (double)stepLog);
            if (b < 0) b = 0;
            if (b >= buckets) b = buckets - 1;
            counts[b]++;
        }
    }
    int64_t seen = 0;
    for (int b = 0; b < buckets; b++) {
        seen += counts[b];
        if (seen >= n) return std::exp(lowestLog + (b + 1) * stepLog);
    }
    // more than half of the database has the maximum length: the reference aborts here
    // (bare `throw;`).  Removing every removable clause is the closest defined behaviour.
    return std::numeric_limits<float>::max();
}

void ClauseDb::syncActivitiesFromDevice(cudaStream_t stream) {
    // apply host-decided rescales to the device copies, then bring the activities home
    applyPendingDeviceRescales(stream);
    downloadActivities(stream);
}

void ClauseDb::copyActivitiesFrom(const ClauseDb &other) {
    GSS_CHECK(other.maxLen_ == maxLen_);
    for (int s = 1; s <= maxLen_; s++) {
        PerLen &pl = *perLen_[s];
        const PerLen &po = *other.perLen_[s];
        GSS_CHECK(pl.meta.size() == po.meta.size());
        for (size_t i = 0; i < pl.meta.size(); i++) pl.meta[i].activity = po.meta[i].activity;
    }
    pendingDeviceRescales_ = 0; // the copied values are final
}

void ClauseDb::reduceDb(cudaStream_t stream) {
    syncActivitiesFromDevice(stream);
    reduceAfterSync(stream);
}

void ClauseDb::reduceAfterSync(cudaStream_t stream) {
    reduceHost();
    for (int s = maxLen_; s >= 3; s--) {
        PerLen &pl = *perLen_[s];
        // give memory back when the arena is mostly empty (reference: CorrespArr.cu:103-113)
        if (pl.dev.capacity() > 1024 && wordsFor(s, (int64_t)pl.meta.size()) * 3 < pl.dev.capacity()) {
            GSS_CUDA(cudaStreamSynchronize(stream));
            pl.dev.free();
            pl.idsDev.free();
            pl.actsDev.free();
        }
    }
    int64_t dummy = 0;
    if (!uploadDirty(stream, &dummy)) GSS_DIE("out of device memory while re-uploading the reduced clause database");
    GSS_CUDA(cudaStreamSynchronize(stream));
    logger_.log(2, "c Done reducing gpu clause db, clause count is " + std::to_string(stats_.clauses) + "\n");
}

void ClauseDb::reduceHost() {
    addedAtLastReduce_ = stats_.added;
    reduceDbs_++;
    float act = approxNthAct(stats_.clauses / 2);
    logger_.log(2, "c Reducing gpu clause db, keeping clauses with act >= " + std::to_string(act) + "\n");
    // Clauses.cu:249-282 with minLimLbd = 0, maxLimLbd = MAX_CL_SIZE, lbd = clause length:
    // lengths 1 and 2 are never touched; a clause of length s >= 3 stays iff s < max and
    // activity >= act.
    for (int s = maxLen_; s >= 3; s--) {
        PerLen &pl = *perLen_[s];
        int64_t n = (int64_t)pl.meta.size(), to = 0;
        if (n == 0) continue;
        for (int64_t idx = 0; idx < n; idx++) {
            const ClauseMeta m = pl.meta[idx];
            if (s < maxLen_ && m.activity >= act) {
                if (to != idx) {
                    for (int i = 0; i < s; i++) pl.lits[wordPos(s, to, i)] = pl.lits[wordPos(s, idx, i)];
                    pl.meta[to] = m;
                }
                to++;
            }
        }
        stats_.clauses -= n - to;
        stats_.lengthSum -= (n - to) * s;
        pl.meta.resize(to);
        pl.lits.resize(wordsFor(s, to));
        pl.fullReupload = true;
        pl.dirtyFrom = 0;
    }
    // clause indices change anyway: put every arena back in first-literal order
    for (int s = 1; s <= maxLen_; s++)
        if ((int64_t)perLen_[s]->meta.size() >= kSortMinClauses) sortArena(s);
}

void ClauseDb::writeCnf(FILE *f, int varCount) const {
    // Clauses.cu:527-549 (the header goes to stdout in the reference; here to the file)
    fprintf(f, "p cnf %d %ld\n", varCount, (long)stats_.clauses);
    std::vector<int> lits;
    int64_t id;
    for (int s = 1; s <= maxLen_; s++) {
        int n = count(s);
        for (int idx = 0; idx < n; idx++) {
            getClause(s, idx, lits, id);
            for (int l : lits) fprintf(f, "%d ", (litVar(l) + 1) * (litSign(l) ? -1 : 1));
            fprintf(f, "0\n");
        }
    }
}

} // namespace gss
