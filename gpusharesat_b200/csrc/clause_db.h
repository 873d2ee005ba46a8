// clause_db.h -- host + device clause database (replaces the reference's Clauses/ClauseUpdates,
// gpuShareLib/Clauses.{cuh,cu}, ClauseUpdates.{cuh,cu}).
//
// Layout (B200-first, not the reference's 32-interleave): one arena per clause length s, made of
// tiles of 128 clauses.  Inside a tile literal i of clause c sits at word i*128 + c, so one warp
// reads a whole literal row (512 B) with one 128-bit load per lane and every lane carries four
// clauses.  No padding literals: the clause count masks the tail of the last tile.  The host
// mirror has the same layout, so new clauses are uploaded with one plain async copy of the
// touched tiles per length -- no staging format and no scatter kernel (reference:
// ClauseUpdates + initClauses, GpuRunner.cu:64-66).
#pragma once
#include "common.h"
#include "mem.h"
#include <memory>
#include <mutex>
#include <vector>

namespace gss {


// One non-empty clause length, as the kernels see it.
struct LenDir {
    const int32_t *base; // device arena of this length
    int32_t len;
    int32_t count;   // clauses of this length
    int32_t tileEnd; // cumulative count of THIS DEVICE's tiles including this length (longest length first)
    int32_t firstTile; // first tile of this length that THIS DEVICE checks (multi-GPU: its contiguous share)
    const int64_t *ids; // device copy of the clause ids of this length, indexed by (global) clause index
    float *acts;        // device-resident clause activities (bumped by k_bump_activity), same indexing
    int64_t ascStart;   // ascStart + index = position of a clause of THIS DEVICE's share among the device's clauses in the canonical order
};

struct DbStats {
    int64_t clauses = 0, lengthSum = 0, added = 0;
};

class ClauseDb {
public:
    ClauseDb(double activityDecay, const Logger &logger, size_t pinnedLimitBytes);
    ~ClauseDb();

    void setMaxLen(int maxLen);
    int maxLen() const { return maxLen_; }
    // Multi-GPU: every rank keeps every clause (host mirror and device arenas) but checks only its
    // contiguous share of the tiles of every length (shardFirstTile / localTiles).  Before the first clause.
    void setShard(int rank, int world);
    int shardRank() const { return shardRank_; }
    int shardWorld() const { return shardWorld_; }

    // ---- any thread (locked): reference HostClauses::addClause, Clauses.cu:349-356 ----
    int64_t addClause(const int *lits, int n);
    int64_t addClausesBulk(const int64_t *offsets, const int *lits, int64_t nclauses);

    // ---- GPU thread only ----
    // move pending clauses into the host mirror (reference getUpdatesForDevice, Clauses.cu:318-346)
    void drainPending();
    // copy the tiles touched since the last upload; false when device memory ran out
    bool uploadDirty(cudaStream_t stream, int64_t *bytesCopied);
    // directory of non-empty lengths, longest first; returns the total tile count
    int buildDirectory(std::vector<LenDir> &dir) const;
    int64_t localClausesOf(int64_t n) const; // of a length with n clauses: how many lie in this device's tiles
    int64_t localClauses() const;            // clauses this device checks

    void getClause(int len, int idx, std::vector<int> &lits, int64_t &id) const;
    // append the literals to `out` (no temporary); returns the clause id
    int64_t appendClause(int len, int idx, std::vector<int> &out) const {
        waitMirror();
        const PerLen &pl = *perLen_[len];
        size_t at = out.size();
        out.resize(at + len);
        const int32_t *p = pl.lits.data() + wordPos(len, idx, 0);
        for (int i = 0; i < len; i++) out[at + i] = p[(size_t)i * kTileClauses];
        return pl.ids[(size_t)idx];
    }
    // hint the caches about a clause that is about to be read (hits arrive sorted by index)
    void prefetchClause(int len, int idx) const {
        const PerLen &pl = *perLen_[len];
        __builtin_prefetch(&pl.ids[(size_t)idx]);
        __builtin_prefetch(pl.lits.data() + wordPos(len, idx, 0));
        if (len > 1) __builtin_prefetch(pl.lits.data() + wordPos(len, idx, 1));
    }
    void prefetchMeta(int len, int idx) const { __builtin_prefetch(&perLen_[len]->ids[(size_t)idx]); }
    int64_t clauseId(int len, int idx) const { waitMirror(); return perLen_[len]->ids[(size_t)idx]; }
    float activity(int len, int idx) const { waitMirror(); return perLen_[len]->acts[(size_t)idx]; }
    void bumpActivity(int len, int idx); // Clauses.cu:231-237
    // same bump from several threads at once (large hit lists are processed per solver in
    // parallel); returns true when the activity passed the rescale limit -- the caller then calls
    // rescaleIfNeeded() once the parallel section is over
    bool bumpActivityAtomic(int len, int idx) {
        float *p = &perLen_[len]->acts[(size_t)idx];
        uint32_t *bits = reinterpret_cast<uint32_t *>(p);
        uint32_t old = __atomic_load_n(bits, __ATOMIC_RELAXED), want;
        float nv;
        do {
            float cur;
            memcpy(&cur, &old, 4);
            nv = cur + actIncr_;
            memcpy(&want, &nv, 4);
        } while (!__atomic_compare_exchange_n(bits, &old, want, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
        return nv > 1e19f;
    }
    void rescaleIfNeeded(bool needed) { if (needed) rescaleActivity(); }
    int count(int len) const { return len <= maxLen_ ? (int)perLen_[len]->n : 0; }

    // reference HostClauses::reduceDb (actOnly), Clauses.cu:426-465 / 249-282.  With device-resident
    // activities (the product path of a single-device sharer) the whole reduction runs on the device
    // (reduce.cu): threshold from a device histogram, keep flags, stable sort by first literal, one
    // permutation pass into fresh arenas -- the host mirror is refreshed by an asynchronous copy.
    // Otherwise (several devices behind one front-end, CPU test rig): host compaction + re-upload.
    void reduceDb(cudaStream_t stream);
    // Streamed clauses are appended behind the first-literal-sorted part of their arena; once the unsorted
    // tails have grown past a share of the database the arenas are put back in order on the device
    // (same permutation pass as reduceDb, nothing removed).  Clause indices change: the caller makes sure
    // no run is in flight and nothing refers to indices any more, exactly as for reduceDb.
    bool resortDue() const;
    void resortOnDevice(cudaStream_t stream);
    void setDeviceReduce(bool on) { deviceReduce_ = on; }
    int64_t unsortedClauses() const;
    // the two halves of reduceDb, and (several devices in one process: every device keeps the whole
    // database, the activities are authoritative on the first one) taking the activities from a twin
    void syncActivitiesFromDevice(cudaStream_t stream);
    void reduceAfterSync(cudaStream_t stream);
    void copyActivitiesFrom(const ClauseDb &other);
    void reduceHost(); // the host half: pick the threshold, compact the mirror (no device work)
    // Device-resident activities (product path): hits bump them on the GPU, the host copy is only
    // refreshed right before a reduceDb.
    void setDeviceActivities(bool on) { deviceActs_ = on; }
    bool deviceActivities() const { return deviceActs_; }
    float activityIncrement() const { return actIncr_; }
    void downloadActivities(cudaStream_t stream);
    // a bump on the device pushed an activity past the rescale limit (Clauses.cu:231-237)
    void rescaleAfterDeviceOverflow() { rescaleActivity(); }
    // rescales decided on the host (drain, device overflow) that the device copies have not seen yet
    void applyPendingDeviceRescales(cudaStream_t stream);
    // reference approxNthAct, Clauses.cu:492-525; its pieces are shared with the device-side reduce
    float approxNthAct(int64_t n) const;
    static constexpr int kActBuckets = 20000;
    static int actBucket(float activity);
    static float thresholdFromHistogram(const int64_t *counts, int64_t n);
    static const std::vector<uint32_t> &actBucketBounds();
    void writeCnf(FILE *f, int varCount) const; // Clauses.cu:527-549

    const DbStats &stats() const { return stats_; }
    int64_t addedAtLastReduceDb() const { return addedAtLastReduce_; }
    int64_t reduceDbCount() const { return reduceDbs_; }
    // largest variable index referenced by any clause + 1 (tables must cover it)
    int maxVarPlusOne() const { return maxVarPlusOne_; }

private:
    struct PerLen {
        int64_t n = 0;            // clauses of this length
        HostBuf<int32_t> lits;    // tiled host mirror
        HostBuf<int64_t> ids;     // GpuClauseId handed out by addClause, per clause
        HostBuf<float> acts;      // activity: bumped per hit, decayed per added clause (Clauses.cu:200-237)
        DevBuf<int32_t> dev;      // device arena (grows in place: vmem.cc)
        DevBuf<int64_t> idsDev;   // clause ids (the GPU emits them with the hits)
        DevBuf<float> actsDev;    // clause activities live on the device between two reduceDb calls
        DevBuf<int32_t> devAlt;   // the second set of arenas: a device-side reduce / re-sort permutes into it and swaps
        DevBuf<int64_t> idsAlt;
        DevBuf<float> actsAlt;
        int64_t actsOnDevice = 0; // clauses [0, actsOnDevice) have their authoritative activity on the device
        int64_t dirtyFrom = 0;    // first clause index not yet on the device
        int64_t sortedN = 0;      // clauses [0, sortedN) are in first-literal order
        bool fullReupload = false;
    };
    // device-side reduce / re-sort (reduce.cu); false: could not get the memory, nothing changed
    bool permuteOnDevice(cudaStream_t stream, bool dropByActivity);
    struct PermScratch; // key / value / histogram arrays of the pass, kept between passes (reduce.cu)
    struct PermScratchDeleter {
        void operator()(PermScratch *p) const;
    };
    std::unique_ptr<PermScratch, PermScratchDeleter> permScratch_;
    void initPermScratch();
    void releaseSpare(cudaStream_t stream);
    // The host mirror of the clauses [0, n) of every length is being refreshed by an asynchronous
    // device-to-host copy (after a device-side reduce / re-sort): wait for it before the host reads that
    // part of the mirror or reallocates any of its buffers.
    void waitMirror() const;
    mutable bool mirrorPending_ = false;
    cudaEvent_t mirrorEv_ = nullptr, permuteDoneEv_ = nullptr;
    cudaStream_t mirrorStream_ = nullptr;
    bool deviceReduce_ = true;
    static size_t wordsFor(int len, int64_t count) {
        return (size_t)((count + kTileClauses - 1) / kTileClauses) * kTileClauses * (size_t)len;
    }
    static size_t wordPos(int len, int64_t idx, int i) {
        return (size_t)(idx / kTileClauses) * kTileClauses * (size_t)len + (size_t)i * kTileClauses +
               (size_t)tileSlot((int)(idx % kTileClauses));
    }
    void appendToMirror(const int *lits, int n, int64_t id);
    // Reorder the clauses of one length by their first literal (ids and activities move along).
    // Consecutive clauses of a tile then gather from neighbouring entries of the level-1 table, so
    // the 128 first-literal gathers of a warp touch a handful of L2 sectors instead of 128.
    // Changes clause indices: only when no run refers to this arena (bulk load into an empty
    // arena, reduceDb).
    void sortArena(int len);
    void rescaleActivity();

    // Multi-GPU: rank r checks the contiguous share [tiles*r/world, tiles*(r+1)/world) of every length
    // array.  Every rank holds the whole arena, so the shares simply move as the arrays grow -- no data
    // does; contiguous shares keep a rank's reads sequential (with the first version's interleaved
    // split -- every world-th tile -- rank 0's sweep took 84 us at 8 ranks against 70 us on one GPU).
    int64_t shardFirstTile(int64_t globalTiles) const { return globalTiles * shardRank_ / shardWorld_; }
    int64_t localTiles(int64_t globalTiles) const {
        return globalTiles * (shardRank_ + 1) / shardWorld_ - globalTiles * shardRank_ / shardWorld_;
    }
    static constexpr int64_t kSortMinClauses = 1024;
    // re-sort on the device once the unsorted tails hold >= 64 k clauses and >= 1/8 of the database
    // (GPUSHARE_RESORT_MIN_CLAUSES overrides the first bound: tests)
    int64_t resortMinClauses_ = 65536;
    static constexpr int64_t kResortShare = 8;
    int maxLen_ = kDefaultMaxClauseLen;
    int shardRank_ = 0, shardWorld_ = 1;
    std::vector<std::unique_ptr<PerLen>> perLen_;
    const Logger &logger_;
    size_t pinnedLimit_;

    // pending clauses, shared with solver threads
    std::mutex pendingLock_;
    std::vector<int> pendingLits_;
    std::vector<int> pendingLens_;
    int64_t nextId_ = 0;      // guarded by pendingLock_
    int64_t pendingFirstId_ = 0;
    bool frozenMaxLen_ = false;

    float actIncr_ = 1.0f;
    float actDecay_;
    DbStats stats_;
    bool deviceActs_ = false;
    int pendingDeviceRescales_ = 0; // rescales decided on the host, not yet applied to the device copies
    int64_t addedAtLastReduce_ = 0;
    int64_t reduceDbs_ = 0;
    int maxVarPlusOne_ = 0;
};

} // namespace gss
