// kernels.cuh -- launch interface of the sm_100a kernels (kernels.cu).
//
// Device data (all 32-bit integer / bitwise work, no tensor cores -- this is not a contraction):
//   A1  level-1 table, per solver group: one {F, U} pair per LITERAL (8 B, one LDG.64, no sign
//       select).  Bit b of F: "the literal can be false in some slot summarised by aggregate bit
//       b"; U: "its variable can be undefined".  16 B per variable -> 16 MB for 1 M variables,
//       L2 resident.  (reference: MultiAgg, 12 B per variable + sign select, Assigs.cuh:44-48)
//   T2  level-2 table: {def, tru} per (variable, solver), row-major by variable so the 32 solver
//       words of a variable are one coalesced 256 B row (lane = solver).  (reference: one
//       MultiLBool array per solver, Assigs.cuh:105-111)
//   clause arenas: see clause_db.h (128-clause tiles, literal-row major, LDG.128 per lane).
#pragma once
#include "clause_db.h"
#include "common.h"

namespace gss {

struct DeviceTables {
    uint2 *a1 = nullptr; // [nGroups][2 * varCap]
    uint2 *t2 = nullptr; // [varCap][solverStride]
    int varCap = 0;
    int solverStride = 0;
    int nGroups = 0;
};

constexpr int kMaxGroups = 8; // solver groups of 32 (256 solver threads)
// Direct pipeline: a solver's hit records are appended to kRecBuckets sub-lists, bucket = the hit
// clause's position in the database's canonical order (length ascending, index ascending) scaled to
// 0..kRecBuckets-1.  Sorting every bucket on its own (a warp each) and reading the buckets in order
// gives the solver's hits in the canonical order; the appends of a solver spread over many counters.
constexpr int kRecBuckets = 256;
constexpr int kCtrStride = 4; // counters are 32 bytes apart: atomics on one L2 sector serialise

// device-side counters of one run
struct Counters {
    unsigned int nHits;
    unsigned int pad;
    unsigned long long exactTests;        // (clause, solver) pairs the exact pass looked at
    unsigned int nSurvivors[kMaxGroups];  // per solver group
};

struct LaunchDims {
    int blocks = 0;  // <= 0: derive from occupancy
    int threads = 256;
};

struct CheckArgs {
    const LenDir *dir; // device copy of the length directory
    int nDir;
    int totalTiles;               // tiles checked by THIS device
    int shardRank, shardWorld;    // (informational: the share of this device is in LenDir::firstTile / tileEnd)
    const SolverRunParams *params; // device, all solvers
    int groupBase;                 // first solver of this group
    int groupSolvers;              // solvers in this group (<= 32)
    uint32_t aggStart;             // aggregate bits in use by this group this run
    int aggStartOnDevice;          // != 0: the kernels derive aggStart from params (host never saw them)
    DeviceTables tables;
    Survivor *survivors;           // this group's survivor list
    unsigned int survCap;
    HitRecord *hits;
    unsigned int hitCap;
    Counters *counters;
    // multi-GPU over peer memory: the LAST k_exact launch of a batch also publishes the result (the
    // block that finishes last writes the slot header and the done flag: no extra kernel launch)
    long long *peerHdr = nullptr;     // this rank's 64-byte slot header in rank 0's gather window
    uint32_t *peerDone = nullptr;     // this rank's done flag in rank 0's window
    unsigned int *peerTicket = nullptr; // this device: blocks that have finished
    uint32_t peerSeq = 0;
    int peerGroups = 0;               // solver groups whose survivor counters count for the overflow flag
    // direct pipeline: k_exact appends its hits to PER-SOLVER record lists instead of one global hit
    // buffer (lane = solver: one 64-bit atomicAdd per lane and warp step reserves the slots and adds up
    // the literal count of that solver's result stream).  recKeys == nullptr: global hit buffer.
    unsigned long long *solverCtr = nullptr; // [solver][bucket] x kCtrStride: records (low 32 bits) | literals (high 32 bits)
    unsigned long long *recKeys = nullptr;   // [solver][bucket][recCap / kRecBuckets]  clause length << 32 | clause index
    uint32_t *recMasks = nullptr;            // same shape: slot mask of the hit
    unsigned int recCap = 0;                 // per solver, a power of two >= kRecBuckets
    long long totalClauses = 0;              // clauses in the database (bucket scaling)
};

void launchFillTables(const DeviceTables &t, int varFrom, cudaStream_t s, int64_t *launches);
void launchApplyUpdates(const VarUpdate *upd, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver,
                        int64_t avail, const DeviceTables &t, int numSMs, cudaStream_t s, int64_t *launches,
                        VarUpdate *keep = nullptr, SolverRunParams *keepParams = nullptr);
void launchCollapse(const VarUpdate *upd, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver,
                    int64_t avail, const DeviceTables &t, int numSMs, cudaStream_t s, int64_t *launches);
// production: aggregate filter + survivor compaction, then the exact pass on the survivors
void launchCheck(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches);
// level 2 alone, on the survivor list the last level-1 launch left behind (bench)
void launchExactOnly(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches);
// level 1 alone (bench: per-kernel timing of the dominant production kernel)
void launchFilterOnly(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches);
// bench: the level-1 kernel exists in several variants (kernels.cu: kFilterVariants)
int filterVariantCount();
const char *filterVariantName(int v);
void setFilterVariant(int v);
int exactVariantCount();
const char *exactVariantName(int v);
void setExactVariant(int v);
// bench-only dense mode: no filter, no early exit
void launchCheckDense(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches);
// the same against an L2-resident table: `sliced` = the group's T2 rows cut into slices of 8 solvers
// ([slice][var][8], 4 x varCap x 8 entries), one sweep of the clauses per slice
void launchSliceTables(const DeviceTables &t, int groupBase, int groupSolvers, uint2 *sliced, int numSMs, cudaStream_t s,
                       int64_t *launches);
void launchCheckDenseSliced(const CheckArgs &a, const uint2 *sliced, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches);

// Post-processing of a large hit list on the device (the host hand-over is the bottleneck of a run
// that returns 10^5 hits): radix-sort the hits by (solver, length, index) -- the reproducible
// hand-over order --, then emit for every hit its clause id and its literals as one contiguous
// stream.  The host then builds each solver's batch with sequential copies only.
struct PostBuffers {
    unsigned long long *keysIn, *keysOut; // n
    unsigned int *valsIn, *valsOut;       // n
    long long *litPos;                    // n + 1 (last = total literal count)
    SortedHit *sorted;                    // n
    int32_t *lits;                        // litCap
    long long litCap;
    void *temp;
    size_t tempBytes;
};
size_t postprocessTempBytes(unsigned int n);
// phase 1: sort + literal positions (litPos[n] = total); phase 2 (after the host has made room for the
// literals): emit records and literals
// solverBits / lenBits / idxBits: bits needed for the largest solver index, clause length and clause index
void launchPostSort(const HitRecord *hits, unsigned int n, const PostBuffers &b, int solverBits, int lenBits, int idxBits,
                    cudaStream_t s, int64_t *launches);
void launchPostEmit(const HitRecord *hits, unsigned int n, const LenDir *dir, int nDir, const PostBuffers &b,
                    cudaStream_t s, int64_t *launches);

// Activity bumps of a run's hits on the device (reference: one host-side bump per hit record,
// Clauses.cu:231-237 / GpuRunner.cu:375-378).  `recs` are HitRecord or SortedHit records (len / idx
// at byte offsets 8 / 12); *overflow is set when an activity passes the rescale limit.
void launchBumpActivity(const void *recs, int strideBytes, unsigned int n, const LenDir *dir, int nDir, float inc,
                        int *overflow, cudaStream_t s, int64_t *launches);

// multi-GPU: 64-byte result header {nHits, overflow flag} written on the device, so that every rank
// can tell from the gathered buffers whether some rank has to run again with larger buffers
void launchFinalize(const Counters *counters, unsigned int hitCap, unsigned int survCap, int groups, long long *dstHeader,
                    cudaStream_t s, int64_t *launches);

// peer-memory exchange (peer.cu): flags live in device memory that other GPUs map through CUDA IPC
constexpr int kMaxPeers = 16;
struct PeerFlagList {
    uint32_t *p[kMaxPeers];
    int n;
};
struct PeerPushList {
    uint4 *dst[kMaxPeers];        // every worker's payload area
    uint32_t *mailbox[kMaxPeers]; // every worker's mailbox
    int n;
};
// copy `bytes` of the batch to every worker and then store seq into every mailbox (one launch)
void launchPeerPush(const void *src, long long bytes, const PeerPushList &L, uint32_t seq, unsigned int *ticket, int numSMs,
                    cudaStream_t s, int64_t *launches);
// the same without a staging copy: the deltas are read from the solvers' page-locked buffers (src[s]) and stored into
// this rank's payload area (`local`) and every worker's window in one pass; `prefix` = [header][run parameters]
void launchPeerPushDirect(const VarUpdate *const *src, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver, void *local,
                          const void *prefix, long long prefixWords, const PeerPushList &L, uint32_t seq, unsigned int *ticket,
                          int numSMs, cudaStream_t s, int64_t *launches);
void launchPeerFinalize(const Counters *counters, unsigned int hitCap, unsigned int survCap, int groups, long long *hdr,
                        uint32_t *doneFlag, uint32_t seq, cudaStream_t s, int64_t *launches);
void launchPeerWaitAll(const PeerFlagList &flags, uint32_t value, unsigned long long timeoutNs, int *err, cudaStream_t s,
                       int64_t *launches);
void launchPeerWait(const uint32_t *flag, uint32_t value, unsigned long long timeoutNs, int *err, cudaStream_t s,
                    int64_t *launches);

// ---- direct pipeline (pipeline.cu): deltas read from, results written to, mapped pinned host memory ----
constexpr int kMaxSolvers = kMaxGroups * kMaxSolversPerGroup;
// Header of a result buffer in pinned host memory, written by k_emit (seq last, after a system fence).
struct RunHdr {
    uint32_t seq;     // the run this header belongs to (written last)
    uint32_t flags;   // 1 survivor list overflowed, 2 a solver's record list overflowed, 4 this buffer is too small
    int64_t nTotal;   // hit records of all solvers
    int64_t litTotal; // literals of all solvers
    uint64_t exactTests;
    uint32_t maxRec;  // largest per-solver record count (before clipping)
    uint32_t hasRecords; // multi-process exchange: the sorted record keys and masks lie next to the ids (written by the owning rank's host)
    uint32_t nSurvivors[kMaxGroups];
    struct PerSolver {
        int64_t entryBase; // first entry of this solver in ids[]; its positions start at pos[entryBase + solver]
        int64_t litBase;   // first literal of this solver in lits[]
        int32_t n;         // entries
        int32_t nLits;
    } solver[kMaxSolvers];
};
struct EmitSolver { // what k_emit_sort leaves for k_emit_write, the activity bumps and the parity hook
    long long entryBase, litBase;
    int32_t n;       // entries to write (-1: the solver's result does not fit, nothing is written)
    int32_t nLits;
    uint32_t nSorted; // records in the solver's sorted list
    uint32_t pad;
};
struct EmitArgs {
    const LenDir *dir;
    int nDir;
    int nSolvers;
    unsigned long long *solverCtr; // [solver][kRecBuckets] x kCtrStride
    unsigned long long *recKeys; // [solver][bucket][recCap / kRecBuckets]: the buckets k_exact filled (sorted in place when large)
    uint32_t *recMasks;
    unsigned long long *sortKeys; // [solver][recCap]: the solver's records in canonical order
    uint32_t *sortMasks;
    int32_t *recPos;             // [solver][recCap + 1] literal positions of the sorted list
    long long *bucketBase;       // [nSolvers * kRecBuckets][2]: entries / literals of the solver before every bucket's list, then [nSolvers][2] totals (k_emit_scan)
    EmitSolver *solverInfo;      // [solver]
    unsigned int recCap;         // power of two
    Counters *counters;
    unsigned int survCap;
    int groups;
    unsigned int *solverDone = nullptr; // [nSolvers] k_emit_fused: blocks of the solver whose buckets are sorted (nullptr: two kernels)
    unsigned int *ticket;        // [0] finished blocks of k_emit_write, [1] flag word, [2] fullest bucket, [3] finished blocks of k_emit_scan, then start order of k_emit_fused
    uint32_t seq;
    // result buffer in mapped pinned host memory
    RunHdr *hdr;
    int64_t *ids;   // entryCap
    int32_t *pos;   // entryCap + nSolvers (one more position than entries per solver)
    int32_t *lits;  // litCap
    long long entryCap, litCap;
    // optional (multi-process exchange): the sorted records themselves, next to the ids, for the
    // front-end's activity bumps (keys) and parity hook (masks); nullptr: stay on the device only
    unsigned long long *keysOut = nullptr; // entryCap
    uint32_t *masksOut = nullptr;          // entryCap
};
// per-solver sort by (length, index) = the reproducible hand-over order, clause ids, literal positions and
// the literal stream of every solver, written straight into the result buffer in host memory
void launchEmit(const EmitArgs &a, cudaStream_t s, int64_t *launches);
// deltas of every solver read from src[s] (mapped pinned host memory or device memory), a copy kept in `keep`
// side jobs of the run's first kernel: fetch the run header from page-locked host memory, zero the run's counters
struct ApplyExtra {
    const uint32_t *headHost = nullptr;
    uint32_t *headDev = nullptr;
    int headWords = 0;
    uint32_t *zeroA = nullptr; // Counters
    int zeroAWords = 0;
    uint32_t *zeroB = nullptr; // per-solver record counters
    int zeroBWords = 0;
};
void launchApplyDirect(const VarUpdate *const *src, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver,
                       const DeviceTables &t, VarUpdate *keep, int numSMs, cudaStream_t s, int64_t *launches,
                       const ApplyExtra &x = ApplyExtra());
// activity bumps from the sorted per-solver record lists of a finished run
void launchBumpFromRecs(const unsigned long long *recKeys, unsigned int recCap, const EmitSolver *solverInfo, int nSolvers,
                        unsigned int maxCount, const LenDir *dir, int nDir, float inc, int *overflow, cudaStream_t s, int64_t *launches);

// the same from a flat list of sorted record keys (length << 32 | index), e.g. another process's result
void launchBumpFromKeys(const unsigned long long *keys, long long n, const LenDir *dir, int nDir, float inc, int *overflow,
                        cudaStream_t s, int64_t *launches);

// register-only LOP3 micro-benchmark: thread-level LOP3 per second on this device
double measureLop3Peak(int numSMs, cudaStream_t s, int64_t *launches);

} // namespace gss
