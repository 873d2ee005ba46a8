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
    if (const char *e = getenv("GPUSHARE_RESORT_MIN_CLAUSES")) resortMinClauses_ = std::max(1ll, atoll(e));
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
        PerLen &pl = *perLen_[s];
        pl.lits.setPinnedLimit(pinnedLimit_);
        pl.ids.setPinnedLimit(pinnedLimit_);
        pl.acts.setPinnedLimit(pinnedLimit_);
        // the host mirror grows in place too (address space now, page-locked chunk by chunk: mem.h)
        pl.lits.setInPlace((size_t)8 << 30);
        pl.ids.setInPlace((size_t)2 << 30);
        pl.acts.setInPlace((size_t)1 << 30);
        // the arenas grow for as long as the solvers learn: in place (vmem.cc), never by realloc + copy
        pl.dev.setInPlace();
        pl.idsDev.setInPlace();
        pl.actsDev.setInPlace();
        pl.devAlt.setInPlace();
        pl.idsAlt.setInPlace();
        pl.actsAlt.setInPlace();
    }
    initPermScratch();
}

ClauseDb::~ClauseDb() {
    if (mirrorPending_ && mirrorEv_) cudaEventSynchronize(mirrorEv_);
    if (mirrorEv_) cudaEventDestroy(mirrorEv_);
    if (permuteDoneEv_) cudaEventDestroy(permuteDoneEv_);
    if (mirrorStream_) cudaStreamDestroy(mirrorStream_);
}

void ClauseDb::waitMirror() const {
    if (!mirrorPending_) return;
    HostProf hp("waitMirror (blocked)");
    GSS_CUDA(cudaEventSynchronize(mirrorEv_));
    mirrorPending_ = false;
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
    int64_t idx = pl.n;
    size_t need = wordsFor(n, idx + 1);
    // (appending does not touch what an asynchronous mirror refresh writes, reallocating would)
    if (mirrorPending_ && (need > pl.lits.capacity() || (size_t)idx + 1 > pl.ids.capacity() || (size_t)idx + 1 > pl.acts.capacity()))
        waitMirror();
    if (need > pl.lits.size()) pl.lits.resize(need); // new tile, zero-filled
    int32_t *dst = pl.lits.data() + wordPos(n, idx, 0);
    int maxLit = 0;
    for (int i = 0; i < n; i++) {
        dst[(size_t)i * kTileClauses] = lits[i];
        maxLit = std::max(maxLit, lits[i]);
    }
    if (litVar(maxLit) + 1 > maxVarPlusOne_) maxVarPlusOne_ = litVar(maxLit) + 1;
    pl.ids.push_back(id);
    pl.acts.push_back(actIncr_);
    pl.n++;
    stats_.clauses++;
    stats_.lengthSum += n;
    stats_.added++;
    if (actIncr_ > kRescale) rescaleActivity();
}

void ClauseDb::drainPending() {
    HostProf hp("drainPending");
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
    for (int s = 1; s <= maxLen_; s++) wasEmpty[s] = perLen_[s]->n == 0;
    size_t pos = 0;
    for (size_t c = 0; c < lens.size(); c++) {
        // Clauses.cu:334 + Clauses.cuh:192: the increment decays once per added clause
        actIncr_ /= actDecay_;
        appendToMirror(&lits[pos], lens[c], firstId + (int64_t)c);
        pos += lens[c];
    }
    // bulk load into an empty arena: nothing refers to these clause indices yet, order them
    for (int s = 1; s <= maxLen_; s++)
        if (wasEmpty[s] && perLen_[s]->n >= kSortMinClauses) sortArena(s);
}

void ClauseDb::sortArena(int len) {
    waitMirror();
    PerLen &pl = *perLen_[len];
    const int64_t n = pl.n;
    pl.sortedN = n;
    if (n < 2) return;
    // counting sort by first literal (stable: equal literals keep arrival order)
    int32_t maxLit = 0;
    for (int64_t i = 0; i < n; i++) maxLit = std::max(maxLit, pl.lits[wordPos(len, i, 0)]);
    std::vector<int64_t> bucket((size_t)maxLit + 2, 0);
    for (int64_t i = 0; i < n; i++) bucket[(size_t)pl.lits[wordPos(len, i, 0)] + 1]++;
    for (size_t b = 1; b < bucket.size(); b++) bucket[b] += bucket[b - 1];
    std::vector<int32_t> oldLits(pl.lits.data(), pl.lits.data() + pl.lits.size());
    std::vector<int64_t> oldIds(pl.ids.data(), pl.ids.data() + n);
    std::vector<float> oldActs(pl.acts.data(), pl.acts.data() + n);
    for (int64_t i = 0; i < n; i++) {
        int64_t to = bucket[(size_t)oldLits[wordPos(len, i, 0)]]++;
        for (int k = 0; k < len; k++) pl.lits[wordPos(len, to, k)] = oldLits[wordPos(len, i, k)];
        pl.ids[(size_t)to] = oldIds[(size_t)i];
        pl.acts[(size_t)to] = oldActs[(size_t)i];
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
    HostProf hp("uploadDirty");
    // rescales decided while draining (Clauses.cu:284-291) reach the device copies first
    applyPendingDeviceRescales(stream);
    for (int s = maxLen_; s >= 1; s--) {
        PerLen &pl = *perLen_[s];
        int64_t n = pl.n;
        if (!pl.fullReupload && pl.dirtyFrom >= n) continue;
        const size_t tileWords = (size_t)kTileClauses * s;
        const int64_t tiles = (n + kTileClauses - 1) / kTileClauses;
        const int64_t firstTile = pl.fullReupload ? 0 : pl.dirtyFrom / kTileClauses;
        {
            // Sharding (multi-GPU) splits the WORK, not the storage: every device holds the whole
            // arena (176 MB at 10 M clauses is nothing next to 180 GB), so any rank can resolve any
            // hit (ids, literals) on its device; a rank only checks its share of the tiles.
            size_t total = (size_t)tiles * tileWords, from = (size_t)firstTile * tileWords;
            // (a device buffer that outgrows its address range is re-mapped: not under a mirror refresh)
            if (mirrorPending_ && (total > pl.dev.capacity() || (size_t)n > pl.idsDev.capacity() || (size_t)n > pl.actsDev.capacity()))
                waitMirror();
            if (total > 0) {
                if (!pl.dev.tryReserve(total, pl.fullReupload ? 0 : from, stream)) return false;
                pl.lits.copyToDevice(pl.dev.data() + from, from, total - from, stream);
                if (bytesCopied) *bytesCopied += (int64_t)((total - from) * sizeof(int32_t));
            }
        }
        // clause ids of the same dirty range
        {
            int64_t from = pl.fullReupload ? 0 : pl.dirtyFrom;
            if (n > from) { // (the host arrays are the staging buffers: same element layout as on the device)
                if (!pl.idsDev.tryReserve((size_t)n, (size_t)from, stream)) return false;
                pl.ids.copyToDevice(pl.idsDev.data() + from, (size_t)from, (size_t)(n - from), stream);
                if (bytesCopied) *bytesCopied += (n - from) * (int64_t)sizeof(int64_t);
                if (deviceActs_) {
                    if (!pl.actsDev.tryReserve((size_t)n, (size_t)from, stream)) return false;
                    pl.acts.copyToDevice(pl.actsDev.data() + from, (size_t)from, (size_t)(n - from), stream);
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
        int64_t n = pl.n;
        if (n == 0) continue;
        // a length none of whose tiles is checked here still gets its entry: hits of other ranks
        // are resolved (ids, literals, activity bumps) through this directory
        const int64_t allTiles = (n + kTileClauses - 1) / kTileClauses;
        tiles += (int)localTiles(allTiles);
        dir.push_back(LenDir{pl.dev.data(), s, (int32_t)n, tiles, (int32_t)shardFirstTile(allTiles), pl.idsDev.data(),
                             const_cast<float *>(pl.actsDev.data()), 0});
    }
    // Canonical position (length ascending, index ascending) of a clause among THIS DEVICE's clauses: the
    // directory is longest first, walk it backwards for the ascending prefix.  (With the whole database as the
    // scale, a device that checks 1/N of every length would use 1/N of the record buckets: N times fuller
    // buckets, sorted in global memory -- 8 GPUs: check + emit 168 -> 490 us.)
    int64_t before = 0;
    for (size_t k = dir.size(); k-- > 0;) {
        const int64_t first = (int64_t)dir[k].firstTile * kTileClauses;
        dir[k].ascStart = before - first;
        before += localClausesOf(dir[k].count);
    }
    return tiles;
}

int64_t ClauseDb::localClausesOf(int64_t n) const {
    const int64_t allTiles = (n + kTileClauses - 1) / kTileClauses;
    const int64_t first = shardFirstTile(allTiles) * kTileClauses, mine = localTiles(allTiles) * kTileClauses;
    return std::max<int64_t>(0, std::min(n - first, mine));
}

int64_t ClauseDb::localClauses() const {
    int64_t total = 0;
    for (int s = 1; s <= maxLen_; s++)
        if (perLen_[s] && perLen_[s]->n) total += localClausesOf(perLen_[s]->n);
    return total;
}

void ClauseDb::getClause(int len, int idx, std::vector<int> &lits, int64_t &id) const {
    waitMirror();
    const PerLen &pl = *perLen_[len];
    lits.resize(len);
    for (int i = 0; i < len; i++) lits[i] = pl.lits[wordPos(len, idx, i)];
    id = pl.ids[(size_t)idx];
}

void ClauseDb::bumpActivity(int len, int idx) {
    float &a = perLen_[len]->acts[(size_t)idx];
    a += actIncr_;
    if (a > kRescale) rescaleActivity();
}

void ClauseDb::rescaleActivity() { // Clauses.cu:284-291
    waitMirror();
    for (int s = 0; s <= maxLen_; s++) {
        PerLen &pl = *perLen_[s];
        for (int64_t i = 0; i < pl.n; i++) pl.acts[(size_t)i] /= kRescale;
    }
    actIncr_ /= kRescale;
    if (deviceActs_) pendingDeviceRescales_++;
}

void ClauseDb::downloadActivities(cudaStream_t stream) {
    if (!deviceActs_) return;
    waitMirror();
    bool any = false;
    for (int s = 1; s <= maxLen_; s++) {
        PerLen &pl = *perLen_[s];
        int64_t n = std::min<int64_t>(pl.actsOnDevice, pl.n);
        if (n <= 0) continue;
        pl.acts.copyFromDevice(pl.actsDev.data(), 0, (size_t)n, stream);
        any = true;
    }
    if (any) GSS_CUDA(cudaStreamSynchronize(stream));
}

// Clauses.cu:492-525: 20000 log-scale buckets over the float range.
namespace {
struct ActScale {
    float lowestLog = std::log(std::numeric_limits<float>::min());
    float largestLog = std::log(std::numeric_limits<float>::max());
    float stepLog = (largestLog - lowestLog) / ClauseDb::kActBuckets;
};
const ActScale kActScale;
} // namespace

int ClauseDb::actBucket(float activity) {
    int b = (int)std::floor(((double)std::log(activity) - (double)kActScale.lowestLog) / (double)kActScale.stepLog);
    if (b < 0) b = 0;
    if (b >= kActBuckets) b = kActBuckets - 1;
    return b;
}

float ClauseDb::thresholdFromHistogram(const int64_t *counts, int64_t n) {
    if (n == 0) return 0.0f;
    int64_t seen = 0;
    for (int b = 0; b < kActBuckets; b++) {
        seen += counts[b];
        if (seen >= n) return std::exp(kActScale.lowestLog + (b + 1) * kActScale.stepLog); // rounded up
    }
    // more than half of the database has the maximum length: the reference aborts here
    // (bare `throw;`).  Removing every removable clause is the closest defined behaviour.
    return std::numeric_limits<float>::max();
}

// bounds[b] = bit pattern of the smallest non-negative float whose bucket is >= b (bounds[0] = 0), so that
// bucket(x) = number of b in [1, kActBuckets) with bounds[b] <= bits(x): the device histogram (reduce.cu)
// finds the bucket by binary search over this table and agrees with actBucket() bit for bit without
// depending on the device's logarithm.  (actBucket is monotone; 31 evaluations per bound, once per process.)
const std::vector<uint32_t> &ClauseDb::actBucketBounds() {
    static const std::vector<uint32_t> bounds = [] {
        std::vector<uint32_t> v((size_t)kActBuckets, 0u);
        auto bucketOfBits = [](uint32_t bits) {
            float x;
            memcpy(&x, &bits, 4);
            return actBucket(x);
        };
        for (int b = 1; b < kActBuckets; b++) {
            // start from the analytic boundary exp(lowest + b * step) and walk the few float neighbours to the
            // exact one (positive floats are ordered like their bit patterns; +inf always reaches the last bucket)
            const float guess = std::exp(kActScale.lowestLog + (float)b * kActScale.stepLog);
            uint32_t bits;
            memcpy(&bits, &guess, 4);
            bits = std::min(std::max(bits, v[(size_t)b - 1]), 0x7F800000u);
            while (bits > v[(size_t)b - 1] && bucketOfBits(bits - 1) >= b) bits--;
            while (bits < 0x7F800000u && bucketOfBits(bits) < b) bits++;
            v[(size_t)b] = bits;
        }
        return v;
    }();
    return bounds;
}

float ClauseDb::approxNthAct(int64_t n) const {
    // The facade passes the clause length as LBD (GpuClauseSharerImpl.cu:125) and the range is
    // [0, MAX_CL_SIZE), so clauses of exactly the maximum length are not counted.
    if (n == 0) return 0.0f;
    waitMirror();
    std::vector<int64_t> counts(kActBuckets, 0);
    for (int s = 1; s < maxLen_; s++) {
        const PerLen &pl = *perLen_[s];
        for (int64_t i = 0; i < pl.n; i++) counts[(size_t)actBucket(pl.acts[(size_t)i])]++;
    }
    return thresholdFromHistogram(counts.data(), n);
}

void ClauseDb::syncActivitiesFromDevice(cudaStream_t stream) {
    // apply host-decided rescales to the device copies, then bring the activities home
    applyPendingDeviceRescales(stream);
    downloadActivities(stream);
}

void ClauseDb::copyActivitiesFrom(const ClauseDb &other) {
    GSS_CHECK(other.maxLen_ == maxLen_);
    waitMirror();
    other.waitMirror();
    for (int s = 1; s <= maxLen_; s++) {
        PerLen &pl = *perLen_[s];
        const PerLen &po = *other.perLen_[s];
        GSS_CHECK(pl.n == po.n);
        if (pl.n) memcpy(pl.acts.data(), po.acts.data(), (size_t)pl.n * sizeof(float));
    }
    pendingDeviceRescales_ = 0; // the copied values are final
}

void ClauseDb::reduceDb(cudaStream_t stream) {
    if (deviceActs_ && deviceReduce_ && shardWorld_ == 1 && permuteOnDevice(stream, true)) return;
    syncActivitiesFromDevice(stream);
    reduceAfterSync(stream);
}

int64_t ClauseDb::unsortedClauses() const {
    int64_t u = 0;
    for (int s = 1; s <= maxLen_; s++) u += perLen_[s]->n - perLen_[s]->sortedN;
    return u;
}

bool ClauseDb::resortDue() const {
    if (!deviceActs_ || !deviceReduce_ || shardWorld_ != 1) return false;
    const int64_t u = unsortedClauses();
    return u >= resortMinClauses_ && u * kResortShare >= stats_.clauses;
}

void ClauseDb::resortOnDevice(cudaStream_t stream) {
    if (!permuteOnDevice(stream, false)) // no memory for the second set of arenas: the tails stay unsorted
        for (int s = 1; s <= maxLen_; s++) perLen_[s]->sortedN = perLen_[s]->n;
}

void ClauseDb::reduceAfterSync(cudaStream_t stream) {
    releaseSpare(stream); // (the host path re-uploads into the primary arenas; the spare set would only hold memory)
    reduceHost();
    for (int s = maxLen_; s >= 3; s--) {
        PerLen &pl = *perLen_[s];
        // give memory back when the arena is mostly empty (reference: CorrespArr.cu:103-113)
        if (pl.dev.capacity() > 1024 && wordsFor(s, pl.n) * 3 < pl.dev.capacity()) {
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
    waitMirror();
    addedAtLastReduce_ = stats_.added;
    reduceDbs_++;
    float act = approxNthAct(stats_.clauses / 2);
    logger_.log(2, "c Reducing gpu clause db, keeping clauses with act >= " + std::to_string(act) + "\n");
    // Clauses.cu:249-282 with minLimLbd = 0, maxLimLbd = MAX_CL_SIZE, lbd = clause length:
    // lengths 1 and 2 are never touched; a clause of length s >= 3 stays iff s < max and
    // activity >= act.
    for (int s = maxLen_; s >= 3; s--) {
        PerLen &pl = *perLen_[s];
        int64_t n = pl.n, to = 0;
        if (n == 0) continue;
        for (int64_t idx = 0; idx < n; idx++) {
            const float a = pl.acts[(size_t)idx];
            if (s < maxLen_ && a >= act) {
                if (to != idx) {
                    for (int i = 0; i < s; i++) pl.lits[wordPos(s, to, i)] = pl.lits[wordPos(s, idx, i)];
                    pl.ids[(size_t)to] = pl.ids[(size_t)idx];
                    pl.acts[(size_t)to] = a;
                }
                to++;
            }
        }
        stats_.clauses -= n - to;
        stats_.lengthSum -= (n - to) * s;
        pl.n = to;
        pl.ids.resize((size_t)to);
        pl.acts.resize((size_t)to);
        pl.lits.resize(wordsFor(s, to));
        pl.fullReupload = true;
        pl.dirtyFrom = 0;
    }
    // clause indices change anyway: put every arena back in first-literal order (a stable sort of the
    // stable compaction: the device-side reduce, reduce.cu, produces exactly the same order)
    for (int s = 1; s <= maxLen_; s++) sortArena(s);
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
