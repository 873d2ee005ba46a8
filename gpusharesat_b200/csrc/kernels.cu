// kernels.cu -- hand-written sm_100a kernels of the clause check (see kernels.cuh for the data
// layout).  The recurrence is the reference's ReportComputer (gpuShareLib/GpuRunner.cu:34-57):
//     justOne = (allFalse & undef) | (justOne & false);   allFalse &= false;
// i.e. three LOP3 per (literal, 32-slot word); a clause is reported where allFalse|justOne != 0.
//
// Kernel inventory
//   k_fill_tables    initial table contents (everything undefined)
//   k_apply_updates  per-run assignment deltas -> T2 rows + the solver's aggregate bits in A1 (one
//                    64-bit atomicAnd + one atomicOr per {F,U} pair)
//   k_collapse       after a run: every slot := the solver's last slot (reference
//                    dSetAllAssigsToLast, Assigs.cu:127-141), deferred until the run is known not
//                    to need a repeat, so a run whose buffers overflowed can simply be re-launched
//   k_filter_t<>     level 1: one warp per 128-clause tile, lane = 4 clauses (LDG.128 literal
//                    rows, LDG.64 gathers from the L2-resident A1), ballot early exit per row,
//                    warp-aggregated survivor append; template over scheduling / cache policy
//                    (kFilterVariants, timed against each other by bench.py --filter-sweep)
//   k_exact_t<>      level 2: lane = solver (coalesced 256 B T2 rows), G survivors per warp step,
//                    software-pipelined, warp-aggregated hit append; multi-GPU: its last block
//                    publishes the rank's result (slot header + done flag in rank 0's memory)
//   k_check_dense    bench-only: every (literal, solver word) pair, no filter, no early exit
//   k_peer_*         multi-GPU exchange over peer memory (push of the batch + mailbox store, polling
//                    fallbacks / wait-for-all, stand-alone publish): see peer.cu
//   k_post_*         large hit lists: sort keys, literal positions, id + literal emission
#include "kernels.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace gss {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- cache-hinted loads ----
// literal rows are streamed once: evict-first so they do not push the assignment tables out of L2
__device__ __forceinline__ int ldStream32(const int32_t *p) { return __ldcs(p); }
__device__ __forceinline__ uint2 ldTable(const uint2 *p) { return __ldg(p); }

__device__ __forceinline__ void step(uint32_t &all, uint32_t &one, uint32_t f, uint32_t u) {
    one = (all & u) | (one & f); // 2 LOP3
    all &= f;                    // 1 LOP3
}

// One {F,U} pair of the level-1 table is 8 aligned bytes: both words are merged with ONE 64-bit
// atomicAnd and ONE atomicOr (bits is a subset of mask; other solvers own the other bits).
__device__ __forceinline__ void mergePair(uint2 *p, uint32_t mask, uint32_t bitsX, uint32_t bitsY) {
    const unsigned long long clear = (unsigned long long)(mask & ~bitsX) | ((unsigned long long)(mask & ~bitsY) << 32);
    const unsigned long long set = (unsigned long long)bitsX | ((unsigned long long)bitsY << 32);
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
    if (clear) atomicAnd(q, ~clear);
    if (set) atomicOr(q, set);
}

__device__ __forceinline__ void writeAggregates(const DeviceTables &t, int solver, int var, uint32_t mask, uint32_t T,
                                                uint32_t F, uint32_t U) {
    uint2 *a = t.a1 + (size_t)(solver / kMaxSolversPerGroup) * 2 * (size_t)t.varCap + 2 * (size_t)var;
    // positive literal (2v) is false where the variable is false; negated (2v+1) where it is true
    mergePair(&a[0], mask, F, U);
    mergePair(&a[1], mask, T, U);
}

__global__ void k_fill_tables(DeviceTables t, int varFrom) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nVars = (size_t)(t.varCap - varFrom);
    // A1: F = 0 (cannot be false), U = ~0 (can be undefined)  (reference MultiAgg{0,~0,0})
    for (int g = 0; g < t.nGroups; g++) {
        uint2 *a = t.a1 + (size_t)g * 2 * t.varCap + 2 * (size_t)varFrom;
        for (size_t i = i0; i < 2 * nVars; i += stride) a[i] = make_uint2(0u, ~0u);
    }
    // T2: undefined everywhere  (reference MultiLBool{0,0})
    uint2 *r = t.t2 + (size_t)varFrom * t.solverStride;
    for (size_t i = i0; i < nVars * t.solverStride; i += stride) r[i] = make_uint2(0u, 0u);
}

// blockIdx.y = solver.  (reference dUpdateAssigs, Assigs.cu:100-116)
// `avail` = number of update records that really are in `upd` (a multi-GPU receiver may have got
// a truncated payload: it must never read past what arrived)
// `keep` / `keepParams` != nullptr (multi-GPU workers): the records and run parameters are also
// copied into the run slot's own buffers, which the check kernels of this batch and the deferred
// k_collapse read later (the window the batch arrived in is overwritten by the next batch).
__global__ void __launch_bounds__(256) k_apply_updates(const VarUpdate *__restrict__ upd,
                                                       const SolverRunParams *__restrict__ params, DeviceTables t,
                                                       long long avail, VarUpdate *__restrict__ keep,
                                                       SolverRunParams *__restrict__ keepParams) {
    __shared__ uint32_t sAgg[kSlots], sSlot[kSlots];
    const int s = blockIdx.y;
    const SolverRunParams &p = params[s];
    if (keepParams && blockIdx.x == 0) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&p);
        uint32_t *dst = reinterpret_cast<uint32_t *>(keepParams + s);
        for (int i = threadIdx.x; i < (int)(sizeof(SolverRunParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const int updStart = p.updStart;
    const int n = (int)max(0ll, min((long long)p.updCount, avail - (long long)updStart)), nGroups = p.nGroups;
    if (n <= 0) return;
    if (threadIdx.x < kSlots) {
        sAgg[threadIdx.x] = p.groupAggBit[threadIdx.x];
        sSlot[threadIdx.x] = p.groupSlotMask[threadIdx.x];
    }
    const uint32_t used = p.usedAggBits;
    __syncthreads();
    const VarUpdate *u = upd + updStart;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const VarUpdate vu = u[i];
        if (keep) keep[updStart + i] = vu;
        t.t2[(size_t)vu.var * t.solverStride + s] = make_uint2(vu.def, vu.tru);
        if (used) {
            uint32_t bt = vu.tru & vu.def, bf = ~vu.tru & vu.def, bu = ~vu.def;
            uint32_t T = 0, F = 0, U = 0;
            for (int g = 0; g < nGroups; g++) {
                uint32_t m = sSlot[g], bit = sAgg[g];
                if (bt & m) T |= bit;
                if (bf & m) F |= bit;
                if (bu & m) U |= bit;
            }
            writeAggregates(t, s, vu.var, used, T, F, U);
        }
    }
}

// blockIdx.y = solver.  (reference dSetAllAssigsToLast, Assigs.cu:127-141)
__global__ void __launch_bounds__(256) k_collapse(const VarUpdate *__restrict__ upd,
                                                  const SolverRunParams *__restrict__ params, DeviceTables t,
                                                  long long avail) {
    const int s = blockIdx.y;
    const SolverRunParams &p = params[s];
    const int n = (int)max(0ll, min((long long)p.updCount, avail - (long long)p.updStart));
    const uint32_t last = p.lastMask, allAgg = p.allAggBits;
    const VarUpdate *u = upd + p.updStart;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        VarUpdate vu = u[i];
        uint32_t def = (vu.def & last) ? ~0u : 0u, tru = (vu.tru & last) ? ~0u : 0u;
        t.t2[(size_t)vu.var * t.solverStride + s] = make_uint2(def, tru);
        if (allAgg) writeAggregates(t, s, vu.var, allAgg, (def & tru) ? allAgg : 0u, (def & ~tru) ? allAgg : 0u,
                                    def ? 0u : allAgg);
    }
}

// Per-warp staging of output records in shared memory.  A global atomicAdd whose result is needed
// (to know where to write) is a ~0.7 us round trip; paying it per tile / per survivor put it on the
// critical path of every warp.  Records are staged per warp and flushed with ONE atomic and one
// coalesced burst when the stage fills up (and at the end of the kernel).
constexpr int kStageCap = 64;   // records per warp (>= 32)
constexpr int kMaxWarpsPerBlock = 8;

template <typename T> struct WarpStage {
    T *buf;  // this warp's kStageCap slots in shared memory
    int n;   // staged records (warp-uniform)
    __device__ __forceinline__ void flush(T *out, unsigned int *counter, unsigned int cap, int lane) {
        if (n == 0) return;
        __syncwarp();
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned int)n);
        base = __shfl_sync(FULL, base, 0);
        for (int i = lane; i < n; i += 32)
            if (base + i < cap) out[base + i] = buf[i];
        __syncwarp();
        n = 0;
    }
    // every lane contributes up to four records (r0..r3, quarter after quarter): four ballots, ONE capacity check
    __device__ __forceinline__ void push4(bool h0, bool h1, bool h2, bool h3, const T &r0, const T &r1, const T &r2, const T &r3, T *out,
                                          unsigned int *counter, unsigned int cap, int lane) {
        const unsigned m0 = __ballot_sync(FULL, h0), m1 = __ballot_sync(FULL, h1), m2 = __ballot_sync(FULL, h2), m3 = __ballot_sync(FULL, h3);
        const int c0 = __popc(m0), c1 = __popc(m1), c2 = __popc(m2), c3 = __popc(m3);
        const int cnt = c0 + c1 + c2 + c3;
        if (cnt == 0) return;
        if (cnt > kStageCap) { // (more than the stage holds: the plain path, one quarter at a time)
            push(h0, r0, out, counter, cap, lane);
            push(h1, r1, out, counter, cap, lane);
            push(h2, r2, out, counter, cap, lane);
            push(h3, r3, out, counter, cap, lane);
            return;
        }
        if (n + cnt > kStageCap) flush(out, counter, cap, lane);
        const unsigned below = (1u << lane) - 1;
        if (h0) buf[n + __popc(m0 & below)] = r0;
        if (h1) buf[n + c0 + __popc(m1 & below)] = r1;
        if (h2) buf[n + c0 + c1 + __popc(m2 & below)] = r2;
        if (h3) buf[n + c0 + c1 + c2 + __popc(m3 & below)] = r3;
        n += cnt;
    }
    // every lane contributes at most one record
    __device__ __forceinline__ void push(bool has, const T &rec, T *out, unsigned int *counter, unsigned int cap, int lane) {
        const unsigned m = __ballot_sync(FULL, has);
        if (!m) return;
        const int cnt = __popc(m);
        if (n + cnt > kStageCap) flush(out, counter, cap, lane);
        if (has) buf[n + __popc(m & ((1u << lane) - 1))] = rec;
        n += cnt;
    }
};

// aggregate bits in use by the solvers of this group; taken from the launch arguments, or -- when
// the host launched without ever seeing the run parameters (multi-GPU receivers) -- from params
__device__ __forceinline__ uint32_t groupAggStart(const CheckArgs &a, int lane) {
    if (!a.aggStartOnDevice) return a.aggStart;
    uint32_t v = lane < a.groupSolvers ? a.params[a.groupBase + lane].usedAggBits : 0u;
    return __reduce_or_sync(FULL, v);
}

// directory lookup: first entry whose cumulative tile count exceeds `tile`
__device__ __forceinline__ int findDir(const int *sTileEnd, int nDir, int tile) {
    int lo = 0, hi = nDir - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (tile < sTileEnd[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// Level 1.  One warp per tile; lane l owns clauses 4l..4l+3 of the tile.  Per literal row: one
// LDG.128 (512 B per warp, fully coalesced), four LDG.64 gathers, twelve LOP3, one vote.
// (reference pass 1 of dFindClauses, GpuRunner.cu:148-172: one clause per thread, 4 B loads,
// 12 B gathers with a sign select, no early exit across the warp)
// ---------------------------------------------------------------------------------------------
// Variants of the same kernel (template parameters), so that the choices can be measured against
// each other on the device in one run (bench.py --filter-sweep; GSS_FILTER_VARIANT selects one):
//   PREFETCH  request rows 0 and 1 of the warp's NEXT tile before working on the current one
//   CONTIG    0: a warp takes every nWarps-th tile; 1: a contiguous run of tiles per warp; 2: a
//             contiguous run per block, the block's warps interleaved inside it
//   ROWMODE   literal rows: 0 = ld.global.cs (evict-first), 1 = ld.global.L1::no_allocate
//   GMODE     level-1 gathers: 0 = ld.global.nc (L1 + L2), 1 = ld.global.cg (L2 only), 2 = 0 + lean addressing / staging
template <int ROWMODE> __device__ __forceinline__ int4 ldRow(const int32_t *p) {
    if constexpr (ROWMODE == 0) {
        return __ldcs(reinterpret_cast<const int4 *>(p));
    } else {
        int4 v;
        asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(p));
        return v;
    }
}
template <int GMODE> __device__ __forceinline__ uint2 ldGather(const uint2 *p) {
    if constexpr (GMODE == 1) return __ldcg(p);
    else return __ldg(p);
}

template <bool PREFETCH, int CONTIG, int ROWMODE, int GMODE, int MINBLOCKS>
__global__ void __launch_bounds__(256, MINBLOCKS) k_filter_t(CheckArgs a) {
    extern __shared__ int sTileEnd[];
    for (int i = threadIdx.x; i < a.nDir; i += blockDim.x) sTileEnd[i] = a.dir[i].tileEnd;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const int warp = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
    const int nWarps = gridDim.x * warpsPerBlock;
    const uint2 *a1 = a.tables.a1 + (size_t)(a.groupBase / kMaxSolversPerGroup) * 2 * (size_t)a.tables.varCap;
    // GMODE 2 ("lean"): the group's table base is made opaque to the compiler, so that it stays ONE 64-bit register
    // pair and a gather address is one IMAD.WIDE.U32 (left to itself the compiler re-adds the group offset to every
    // literal: 4-5 instructions per gather, 16-20 of the ~64 of a row step); survivors are staged with one push4
    if constexpr (GMODE == 2) asm volatile("" : "+l"(a1));
    const uint32_t start = groupAggStart(a, lane);
    if (start == 0) return; // no frozen slot in this group
    __shared__ Survivor sStage[kMaxWarpsPerBlock][kStageCap];
    WarpStage<Survivor> stage{sStage[threadIdx.x >> 5], 0};
    unsigned int *survCounter = &a.counters->nSurvivors[a.groupBase / kMaxSolversPerGroup];

    // where a tile lives: every device holds the whole arena, this rank checks a contiguous share of
    // every length's tiles, starting at LenDir::firstTile
    struct Tile {
        const int32_t *row; // this lane's column of the tile's first literal row
        int len, c0, nValid;
        int lenEnd; // first (device-local) tile index of the next length
    };
    auto locate = [&](int tile) -> Tile {
        const int k = findDir(sTileEnd, a.nDir, tile);
        const LenDir d = a.dir[k];
        const int gTile = (tile - (k ? sTileEnd[k - 1] : 0)) + d.firstTile;
        // words 4*lane .. 4*lane+3 of a row hold the clauses lane, lane+32, lane+64, lane+96 (tileSlot)
        const int c0 = gTile * kTileClauses + lane; // global clause index of this lane's first clause
        return Tile{d.base + (size_t)gTile * kTileClauses * d.len + lane * 4, d.len, c0, d.count - c0, sTileEnd[k]};
    };
    // the tile after `t` (device-local index tile + 1): the same length array continues, or look it up
    auto advance = [&](const Tile &t, int tile) -> Tile {
        if (tile + 1 >= t.lenEnd) return locate(tile + 1);
        const int dc = kTileClauses;
        return Tile{t.row + (size_t)dc * t.len, t.len, t.c0 + dc, t.nValid - dc, t.lenEnd};
    };

    int first, last, stepT;
    if constexpr (CONTIG == 1) {
        const int per = (a.totalTiles + nWarps - 1) / nWarps;
        first = warp * per;
        last = min(a.totalTiles, first + per);
        stepT = 1;
    } else if constexpr (CONTIG == 2) {
        // a block takes a contiguous run of tiles and its warps interleave inside it: at any moment
        // the warps of an SM read neighbouring tiles (one DRAM page, neighbouring level-1 entries)
        const int perBlock = ((a.totalTiles + (int)gridDim.x - 1) / (int)gridDim.x + warpsPerBlock - 1) / warpsPerBlock * warpsPerBlock;
        const int b0 = blockIdx.x * perBlock;
        first = b0 + (threadIdx.x >> 5);
        last = min(a.totalTiles, b0 + perBlock);
        stepT = warpsPerBlock;
    } else {
        first = warp;
        last = a.totalTiles;
        stepT = nWarps;
    }
    if (first >= last) return;

    // one tile, whose first row (and, when PREFETCH, second row) has been requested already
    auto process = [&](const Tile &cur, const int4 &r0, const int4 &r1) {
        const int len = cur.len, nValid = cur.nValid; // clause c0 + 32q exists iff nValid > 32q
        const int32_t *row = cur.row;

        uint32_t all0 = nValid > 0 ? start : 0u, all1 = nValid > 32 ? start : 0u;
        uint32_t all2 = nValid > 64 ? start : 0u, all3 = nValid > 96 ? start : 0u;
        uint32_t one0 = 0, one1 = 0, one2 = 0, one3 = 0;

        int4 lits = r0;
        uint32_t alive = 0;
        for (int i = 0; i < len; i++) {
            int4 next = lits;
            if constexpr (PREFETCH) {
                next = r1;
                if (i >= 1 && i + 1 < len) next = ldRow<ROWMODE>(row + (size_t)(i + 1) * kTileClauses);
            } else {
                if (i + 1 < len) next = ldRow<ROWMODE>(row + (size_t)(i + 1) * kTileClauses); // overlaps the gathers
            }
            // a dead clause stays dead: gather only for the live ones (every gather costs a 32 B
            // L2 sector, and after the first literal ~97 % of the clauses are dead)
            const uint2 dead = make_uint2(0u, 0u);
            uint2 g0, g1, g2, g3;
            if constexpr (GMODE == 2) { // (a literal is never negative: unsigned index = no sign extension)
                g0 = (all0 | one0) ? ldGather<GMODE>(a1 + (uint32_t)lits.x) : dead;
                g1 = (all1 | one1) ? ldGather<GMODE>(a1 + (uint32_t)lits.y) : dead;
                g2 = (all2 | one2) ? ldGather<GMODE>(a1 + (uint32_t)lits.z) : dead;
                g3 = (all3 | one3) ? ldGather<GMODE>(a1 + (uint32_t)lits.w) : dead;
            } else {
                g0 = (all0 | one0) ? ldGather<GMODE>(a1 + lits.x) : dead;
                g1 = (all1 | one1) ? ldGather<GMODE>(a1 + lits.y) : dead;
                g2 = (all2 | one2) ? ldGather<GMODE>(a1 + lits.z) : dead;
                g3 = (all3 | one3) ? ldGather<GMODE>(a1 + lits.w) : dead;
            }
            step(all0, one0, g0.x, g0.y);
            step(all1, one1, g1.x, g1.y);
            step(all2, one2, g2.x, g2.y);
            step(all3, one3, g3.x, g3.y);
            alive = (all0 | one0) | (all1 | one1) | (all2 | one2) | (all3 | one3);
            if (!__any_sync(FULL, alive)) break; // all 128 clauses are dead: skip the remaining rows
            lits = next;
        }
        if (__any_sync(FULL, alive)) {
            // survivors go to the warp's stage (one global atomic per ~64 survivors)
            const uint64_t rowTag = (uint64_t)(uintptr_t)row | ((uint64_t)len << 48);
            const int c0 = cur.c0;
            const uint32_t m0 = all0 | one0, m1 = all1 | one1, m2 = all2 | one2, m3 = all3 | one3;
            if constexpr (GMODE == 2) {
                stage.push4(m0 != 0, m1 != 0, m2 != 0, m3 != 0, Survivor{rowTag + 0 * sizeof(int32_t), c0 + 0, m0},
                            Survivor{rowTag + 1 * sizeof(int32_t), c0 + 32, m1}, Survivor{rowTag + 2 * sizeof(int32_t), c0 + 64, m2},
                            Survivor{rowTag + 3 * sizeof(int32_t), c0 + 96, m3}, a.survivors, survCounter, a.survCap, lane);
            } else {
                stage.push(m0 != 0, Survivor{rowTag + 0 * sizeof(int32_t), c0 + 0, m0}, a.survivors, survCounter, a.survCap, lane);
                stage.push(m1 != 0, Survivor{rowTag + 1 * sizeof(int32_t), c0 + 32, m1}, a.survivors, survCounter, a.survCap, lane);
                stage.push(m2 != 0, Survivor{rowTag + 2 * sizeof(int32_t), c0 + 64, m2}, a.survivors, survCounter, a.survCap, lane);
                stage.push(m3 != 0, Survivor{rowTag + 3 * sizeof(int32_t), c0 + 96, m3}, a.survivors, survCounter, a.survCap, lane);
            }
        }
    };
    auto following = [&](const Tile &t, int tile) -> Tile { // the tile this warp takes after `tile`
        if constexpr (CONTIG == 1) return advance(t, tile);
        else return locate(tile + stepT);
    };

    if constexpr (PREFETCH) {
        // Software pipeline ACROSS tiles, two register sets in ping-pong (no copies between them, a
        // copy would wait for the load): while tile A is worked on, rows 0 and 1 of tile B -- needed
        // almost always, a tile of 128 clauses rarely dies on its first row -- are in flight, and
        // vice versa.  A warp never waits for a DRAM round trip with nothing else to do.
        Tile tA = locate(first), tB = tA;
        int4 a0 = ldRow<ROWMODE>(tA.row), a1v = tA.len > 1 ? ldRow<ROWMODE>(tA.row + kTileClauses) : a0;
        int4 b0 = a0, b1 = a0;
        for (int tile = first; tile < last; tile += 2 * stepT) {
            const bool hasB = tile + stepT < last;
            if (hasB) {
                tB = following(tA, tile);
                b0 = ldRow<ROWMODE>(tB.row);
                b1 = tB.len > 1 ? ldRow<ROWMODE>(tB.row + kTileClauses) : b0;
            }
            process(tA, a0, a1v);
            if (!hasB) break;
            if (tile + 2 * stepT < last) {
                tA = following(tB, tile + stepT);
                a0 = ldRow<ROWMODE>(tA.row);
                a1v = tA.len > 1 ? ldRow<ROWMODE>(tA.row + kTileClauses) : a0;
            }
            process(tB, b0, b1);
        }
    } else {
        Tile cur = locate(first);
        int4 r0 = ldRow<ROWMODE>(cur.row);
        for (int tile = first; tile < last; tile += stepT) {
            process(cur, r0, r0);
            if (tile + stepT < last) {
                cur = following(cur, tile);
                r0 = ldRow<ROWMODE>(cur.row);
            }
        }
    }
    stage.flush(a.survivors, survCounter, a.survCap, lane);
}

// ---------------------------------------------------------------------------------------------
// Level 1, literal rows staged by TMA.  Same work split and arithmetic as k_filter_t (warp per tile,
// lane = 4 clauses, contiguous run of tiles per warp), but rows 0 and 1 of a tile -- all a tile needs
// 98 % of the time -- no longer travel through registers: one elected lane issues ONE 1-D bulk copy
// (cp.async.bulk, 512 or 1024 contiguous bytes, L2 evict-first) per tile into the warp's private ring
// in shared memory, kRing tiles ahead, completion on an mbarrier.  A warp therefore has the rows of
// the next kRing tiles in flight at no register cost, and the per-tile dependent chain shrinks from
// "row load -> gather -> vote -> gather -> vote" to "gather -> vote -> gather -> vote".  Rows >= 2
// (12 % of the tiles of longer clauses) are fetched on demand with LDG.128 as before.
// ---------------------------------------------------------------------------------------------
constexpr int kRing = 4;
constexpr int kRingSlotBytes = 2 * kTileClauses * (int)sizeof(int32_t); // two literal rows

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulkLoad(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

template <int MINBLOCKS> __global__ void __launch_bounds__(256, MINBLOCKS) k_filter_tma_t(CheckArgs a) {
    extern __shared__ __align__(128) unsigned char sDynRaw[];
    // [ring: warps x kRing x 1 KB][mbarriers: warps x kRing x 8 B][cumulative tile counts: nDir ints]
    const int warpsPerBlock = blockDim.x >> 5;
    unsigned char *ringBase = sDynRaw;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sDynRaw + (size_t)warpsPerBlock * kRing * kRingSlotBytes);
    int *sTileEnd = reinterpret_cast<int *>(bars + warpsPerBlock * kRing);
    for (int i = threadIdx.x; i < a.nDir; i += blockDim.x) sTileEnd[i] = a.dir[i].tileEnd;
    if ((int)threadIdx.x < warpsPerBlock * kRing) mbarInit(smemAddr(bars + threadIdx.x), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * warpsPerBlock + wib;
    const int nWarps = gridDim.x * warpsPerBlock;
    const uint2 *__restrict__ a1 = a.tables.a1 + (size_t)(a.groupBase / kMaxSolversPerGroup) * 2 * (size_t)a.tables.varCap;
    const uint32_t start = groupAggStart(a, lane);
    if (start == 0) return; // no frozen slot in this group
    __shared__ Survivor sStage[kMaxWarpsPerBlock][kStageCap];
    WarpStage<Survivor> stage{sStage[wib], 0};
    unsigned int *survCounter = &a.counters->nSurvivors[a.groupBase / kMaxSolversPerGroup];
    const unsigned char *myRing = ringBase + (size_t)wib * kRing * kRingSlotBytes;
    const uint32_t myRingAddr = smemAddr(myRing), myBars = smemAddr(bars + wib * kRing);
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    struct Cursor { // where a tile lives (every device holds the whole arena; this rank checks a contiguous share)
        const int32_t *base; // the tile's first literal row
        int len, c0;         // clause length; global index of the tile's first clause
        int count;           // clauses of this length
        int lenEnd;          // first (device-local) tile index of the next length
    };
    auto locate = [&](int tile) -> Cursor {
        const int k = findDir(sTileEnd, a.nDir, tile);
        const LenDir d = a.dir[k];
        const int gTile = (tile - (k ? sTileEnd[k - 1] : 0)) + d.firstTile;
        return Cursor{d.base + (size_t)gTile * kTileClauses * d.len, d.len, gTile * kTileClauses, d.count, sTileEnd[k]};
    };
    auto advance = [&](const Cursor &t, int tile) -> Cursor {
        if (tile + 1 >= t.lenEnd) return locate(tile + 1);
        return Cursor{t.base + (size_t)kTileClauses * t.len, t.len, t.c0 + kTileClauses, t.count, t.lenEnd};
    };
    const int per = (a.totalTiles + nWarps - 1) / nWarps;
    const int first = warp * per, last = min(a.totalTiles, first + per);
    if (first >= last) return;

    auto issue = [&](const Cursor &t, int slot) { // rows 0 (and 1) of tile t -> ring slot
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)min(t.len, 2) * kTileClauses * (uint32_t)sizeof(int32_t);
            const uint32_t bar = myBars + slot * 8;
            mbarExpectTx(bar, bytes);
            bulkLoad(myRingAddr + slot * kRingSlotBytes, t.base, bytes, bar, policy);
        }
    };
    Cursor cur = locate(first), pf = cur;
    int pfTile = first;
    for (int k = 0; k < kRing && pfTile < last; k++) {
        issue(pf, k);
        if (pfTile + 1 < last) pf = advance(pf, pfTile);
        pfTile++;
    }
    for (int tile = first; tile < last; tile++) {
        const int n = tile - first, slot = n % kRing;
        mbarWait(myBars + slot * 8, (uint32_t)(n / kRing) & 1u);
        const int len = cur.len;
        const int c0 = cur.c0 + lane, nValid = cur.count - c0; // clause c0 + 32q exists iff nValid > 32q
        const int4 *sl = reinterpret_cast<const int4 *>(myRing + (size_t)slot * kRingSlotBytes) + lane;
        uint32_t all0 = nValid > 0 ? start : 0u, all1 = nValid > 32 ? start : 0u;
        uint32_t all2 = nValid > 64 ? start : 0u, all3 = nValid > 96 ? start : 0u;
        uint32_t one0 = 0, one1 = 0, one2 = 0, one3 = 0;
        int4 lits = sl[0];
        uint32_t alive = 0;
        const int32_t *row = cur.base + lane * 4;
        for (int i = 0; i < len; i++) {
            int4 next = lits;
            if (i + 1 < len) next = i == 0 ? sl[kTileClauses / 4] : ldRow<0>(row + (size_t)(i + 1) * kTileClauses);
            const uint2 dead = make_uint2(0u, 0u);
            uint2 g0 = (all0 | one0) ? __ldg(a1 + lits.x) : dead;
            uint2 g1 = (all1 | one1) ? __ldg(a1 + lits.y) : dead;
            uint2 g2 = (all2 | one2) ? __ldg(a1 + lits.z) : dead;
            uint2 g3 = (all3 | one3) ? __ldg(a1 + lits.w) : dead;
            step(all0, one0, g0.x, g0.y);
            step(all1, one1, g1.x, g1.y);
            step(all2, one2, g2.x, g2.y);
            step(all3, one3, g3.x, g3.y);
            alive = (all0 | one0) | (all1 | one1) | (all2 | one2) | (all3 | one3);
            if (!__any_sync(FULL, alive)) break;
            lits = next;
        }
        // the slot is free again once every lane has read its words (generic-proxy reads ordered before the
        // async-proxy write of the refill): refill it kRing tiles ahead
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (pfTile < last) {
            issue(pf, slot);
            if (pfTile + 1 < last) pf = advance(pf, pfTile);
            pfTile++;
        }
        if (__any_sync(FULL, alive)) {
            const uint64_t rowTag = (uint64_t)(uintptr_t)row | ((uint64_t)len << 48);
            const uint32_t m0 = all0 | one0, m1 = all1 | one1, m2 = all2 | one2, m3 = all3 | one3;
            stage.push(m0 != 0, Survivor{rowTag + 0 * sizeof(int32_t), c0 + 0, m0}, a.survivors, survCounter, a.survCap, lane);
            stage.push(m1 != 0, Survivor{rowTag + 1 * sizeof(int32_t), c0 + 32, m1}, a.survivors, survCounter, a.survCap, lane);
            stage.push(m2 != 0, Survivor{rowTag + 2 * sizeof(int32_t), c0 + 64, m2}, a.survivors, survCounter, a.survCap, lane);
            stage.push(m3 != 0, Survivor{rowTag + 3 * sizeof(int32_t), c0 + 96, m3}, a.survivors, survCounter, a.survCap, lane);
        }
        if (tile + 1 < last) cur = advance(cur, tile);
    }
    stage.flush(a.survivors, survCounter, a.survCap, lane);
}

struct FilterVariant {
    void (*kernel)(CheckArgs);
    int threads;
    const char *name;
    int ringBytesPerWarp = 0; // TMA variants: dynamic shared memory per warp (ring + mbarriers)
};
const FilterVariant kFilterVariants[] = {
    {k_filter_t<false, 0, 0, 0, 5>, 256, "strided tiles, 5 x 256 threads/SM"},
    {k_filter_t<false, 1, 0, 0, 5>, 256, "contiguous tiles per warp, 5 x 256"},
    {k_filter_t<false, 2, 0, 0, 5>, 256, "contiguous tiles per block, warps interleaved, 5 x 256"},
    {k_filter_t<false, 1, 0, 0, 4>, 256, "contiguous per warp, 4 x 256"},
    {k_filter_t<false, 1, 0, 0, 6>, 256, "contiguous per warp, 6 x 256 (40 registers, spills)"},
    {k_filter_t<false, 1, 1, 0, 5>, 256, "contiguous per warp, rows L1::no_allocate, 5 x 256"},
    {k_filter_t<false, 1, 0, 1, 5>, 256, "contiguous per warp, gathers ld.cg (L2 only), 5 x 256"},
    {k_filter_t<false, 2, 0, 0, 5>, 128, "contiguous per block, 10 x 128"},
    {k_filter_t<true, 1, 0, 0, 4>, 256, "contiguous per warp + ping-pong prefetch, 4 x 256"},
    {k_filter_t<true, 1, 0, 0, 5>, 256, "contiguous per warp + ping-pong prefetch, 5 x 256"},
    {k_filter_tma_t<5>, 256, "rows 0-1 by TMA bulk copies into a per-warp ring (4 tiles ahead), 5 x 256", kRing *(kRingSlotBytes + 8)},
    {k_filter_tma_t<6>, 256, "TMA ring, 6 x 256", kRing *(kRingSlotBytes + 8)},
    {k_filter_tma_t<4>, 256, "TMA ring, 4 x 256", kRing *(kRingSlotBytes + 8)},
    {k_filter_t<false, 1, 0, 2, 5>, 256, "contiguous per warp, lean (one-instruction gather addresses, one push4), 5 x 256"},
    {k_filter_t<false, 1, 0, 2, 4>, 256, "contiguous per warp, lean, 4 x 256"},
    {k_filter_t<true, 1, 0, 2, 4>, 256, "contiguous per warp + ping-pong prefetch, lean, 4 x 256"},
};
constexpr int kNumFilterVariants = (int)(sizeof(kFilterVariants) / sizeof(kFilterVariants[0]));
int gFilterVariant = -1; // -1: not chosen yet (GSS_FILTER_VARIANT or the default)
constexpr int kDefaultFilterVariant = 1;

// append one hit per lane with a non-zero mask; one atomic per warp
__device__ __forceinline__ void reportHits(const CheckArgs &a, uint32_t m, int solver, int len, int idx, int lane) {
    const unsigned hitLanes = __ballot_sync(FULL, m != 0);
    if (!hitLanes) return;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(&a.counters->nHits, (unsigned int)__popc(hitLanes));
    base = __shfl_sync(FULL, base, 0);
    if (m) {
        unsigned int pos = base + __popc(hitLanes & ((1u << lane) - 1));
        if (pos < a.hitCap) a.hits[pos] = HitRecord{m, solver, len, idx};
    }
}

// ---------------------------------------------------------------------------------------------
// Level 2.  One warp per survivor, lane = solver of the group.  Per literal: one uniform 4 B
// load and one coalesced T2 row (8 B per lane).  (reference dCheckOneClauseAllSolvers /
// dCheckOneClauseOneSolver, GpuRunner.cu:68-131: serial per thread, one solver after the other)
// ---------------------------------------------------------------------------------------------
// Direct pipeline: append one hit to its solver's record list, in the bucket of the clause's position
// in the canonical order (kernels.cuh: kRecBuckets).  One 64-bit atomic reserves the slot and adds the
// clause's literals to the bucket's literal count.  sLen / sAsc: the directory's lengths (descending)
// and ascending clause prefix in shared memory.
// The directory (one entry per non-empty clause length, longest first) is searched by length from shared memory;
// a database with more distinct lengths than the cache holds (GPUSHARE_MAX_CLAUSE_LEN above kDirCache) reads the
// entries past it from the directory itself.  (Up to round 2 the searches simply stopped at 128 entries: with 129
// distinct lengths -- the import-latency harness: 2-30 and 101-200 -- every hit on the SHORTEST length was resolved
// against the next one's arena.)
constexpr int kDirCache = 128;
__device__ __forceinline__ int dirLenAt(const LenDir *dir, const int *sLen, int i) { return i < kDirCache ? sLen[i] : dir[i].len; }
__device__ __forceinline__ int dirIndexOfLen(const LenDir *dir, const int *sLen, int nDir, int len) { // descending length
    int lo = 0, hi = nDir - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (dirLenAt(dir, sLen, mid) <= len) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ void appendRec(const CheckArgs &a, const int *sLen, const long long *sAsc, int nDir, int solver, int len,
                                          int idx, uint32_t mask) {
    const int lo = dirIndexOfLen(a.dir, sLen, nDir, len);
    const unsigned long long g = (unsigned long long)((lo < kDirCache ? sAsc[lo] : a.dir[lo].ascStart) + idx);
    const unsigned int bucket = (unsigned int)min((unsigned long long)(kRecBuckets - 1), g * kRecBuckets / (unsigned long long)a.totalClauses);
    const unsigned int bucketCap = a.recCap / kRecBuckets;
    const unsigned int slot = (unsigned int)atomicAdd(a.solverCtr + ((size_t)solver * kRecBuckets + bucket) * kCtrStride,
                                                      1ull | ((unsigned long long)(unsigned int)len << 32));
    if (slot < bucketCap) {
        const size_t at = (size_t)solver * a.recCap + (size_t)bucket * bucketCap + slot;
        a.recKeys[at] = ((unsigned long long)(unsigned int)len << 32) | (unsigned int)idx;
        a.recMasks[at] = mask;
    }
}

__device__ __forceinline__ void flushRecs(const CheckArgs &a, WarpStage<HitRecord> &stage, const int *sLen, const long long *sAsc,
                                          int nDir, int lane) {
    if (stage.n == 0) return;
    __syncwarp();
    for (int i = lane; i < stage.n; i += 32) {
        const HitRecord r = stage.buf[i];
        appendRec(a, sLen, sAsc, nDir, r.solver, r.len, r.idx, r.mask);
    }
    __syncwarp();
    stage.n = 0;
}

// G = survivors a warp checks together (their row gathers are independent)
template <int G, int MINBLOCKS> __global__ void __launch_bounds__(256, MINBLOCKS) k_exact_t(CheckArgs a) {
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const unsigned int warp = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
    const unsigned int nWarps = gridDim.x * warpsPerBlock;
    const bool active = lane < a.groupSolvers;
    const int solver = a.groupBase + (active ? lane : 0);
    const uint32_t myStart = active ? a.params[solver].startVals : 0u;
    const uint32_t myAgg = active ? a.params[solver].allAggBits : 0u;
    const uint2 *__restrict__ t2 = a.tables.t2 + solver;
    const size_t stride = (size_t)a.tables.solverStride;

    unsigned int n = a.counters->nSurvivors[a.groupBase / kMaxSolversPerGroup];
    if (n > a.survCap) n = a.survCap;
    const unsigned int nGroups = (n + G - 1) / G;
    unsigned long long tests = 0;
    const uint2 ident = make_uint2(0u, 0u);
    __shared__ HitRecord sStage[kMaxWarpsPerBlock][kStageCap];
    WarpStage<HitRecord> stage{sStage[threadIdx.x >> 5], 0};
    __shared__ int sDirLen[kDirCache];
    __shared__ long long sDirAsc[kDirCache];
    const int nDirS = a.nDir; // (entries past the cache are read from the directory)
    if (a.recKeys) {
        for (int i = threadIdx.x; i < min(nDirS, kDirCache); i += blockDim.x) { sDirLen[i] = a.dir[i].len; sDirAsc[i] = a.dir[i].ascStart; }
        __syncthreads();
    }

    // A warp takes G consecutive survivors at a time; lane g (< G) holds record g of the group.  The
    // chain per survivor is record -> literals -> T2 rows, each a dependent DRAM / L2 round trip:
    // the records of group i+2 and the literals of group i+1 are in flight while group i is checked,
    // and the rows of all G survivors of a group are gathered together (2 literals x G rows in flight).
    auto ldGroup = [&](unsigned int grp) -> uint4 {
        const unsigned int i = grp * G + lane;
        return (grp < nGroups && lane < G && i < n) ? __ldg(reinterpret_cast<const uint4 *>(a.survivors + i))
                                                    : make_uint4(0u, 0u, 0u, 0u);
    };
    // literal `base + lane` of survivor g of the group whose records are in sv (0 past the end)
    auto ldLits = [&](const uint4 &sv, int g, int base) -> int {
        const uint32_t x = __shfl_sync(FULL, sv.x, g), y = __shfl_sync(FULL, sv.y, g);
        const int32_t *p = reinterpret_cast<const int32_t *>(((uint64_t)(y & 0xFFFFu) << 32) | x);
        return base + lane < (int)(y >> 16) ? __ldg(p + (size_t)(base + lane) * kTileClauses) : 0;
    };
    uint4 sv0 = ldGroup(warp), sv1 = ldGroup(warp + nWarps);
    int lit0[G];
#pragma unroll
    for (int g = 0; g < G; g++) lit0[g] = ldLits(sv0, g, 0);
    for (unsigned int grp = warp; grp < nGroups; grp += nWarps) {
        const uint4 sv2 = ldGroup(grp + 2u * nWarps);
        int lit1[G];
#pragma unroll
        for (int g = 0; g < G; g++) lit1[g] = ldLits(sv1, g, 0);
        uint32_t all[G], one[G];
        int len[G], maxLen = 0;
#pragma unroll
        for (int g = 0; g < G; g++) {
            len[g] = (int)(__shfl_sync(FULL, sv0.y, g) >> 16); // 0: no such survivor
            const uint32_t agg = __shfl_sync(FULL, sv0.w, g);
            all[g] = (len[g] > 0 && (agg & myAgg)) ? myStart : 0u;
            one[g] = 0u;
            tests += __popc(__ballot_sync(FULL, all[g] != 0));
            maxLen = max(maxLen, len[g]);
        }
        bool dead = false;
        for (int base = 0; base < maxLen && !dead; base += 32) {
            int myLit[G];
#pragma unroll
            for (int g = 0; g < G; g++) myLit[g] = base == 0 ? lit0[g] : ldLits(sv0, g, base);
            const int nl = min(32, maxLen - base);
            for (int j = 0; j < nl; j += 2) {
                int l0[G], l1[G];
                uint2 e0[G], e1[G];
                // lanes whose solver is already dead for a survivor (or was never selected by the
                // filter) do not load its rows at all
#pragma unroll
                for (int g = 0; g < G; g++) {
                    l0[g] = __shfl_sync(FULL, myLit[g], j);
                    l1[g] = __shfl_sync(FULL, myLit[g], (j + 1) & 31);
                    const bool live = (all[g] | one[g]) != 0;
                    e0[g] = (live && base + j < len[g]) ? ldTable(t2 + (size_t)(l0[g] >> 1) * stride) : ident;
                    e1[g] = (live && base + j + 1 < len[g]) ? ldTable(t2 + (size_t)(l1[g] >> 1) * stride) : ident;
                }
                uint32_t any = 0;
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (base + j < len[g]) step(all[g], one[g], e0[g].x & ((l0[g] & 1) ? e0[g].y : ~e0[g].y), ~e0[g].x);
                    if (base + j + 1 < len[g]) step(all[g], one[g], e1[g].x & ((l1[g] & 1) ? e1[g].y : ~e1[g].y), ~e1[g].x);
                    any |= all[g] | one[g];
                }
                if (!__any_sync(FULL, any)) { dead = true; break; }
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int idx = (int)__shfl_sync(FULL, sv0.z, g);
            const bool has = (all[g] | one[g]) != 0;
            const HitRecord rec{all[g] | one[g], solver, len[g], idx};
            if (a.recKeys) {
                // staged per warp like the global hit buffer; the flush appends them to the per-solver
                // buckets with every lane's atomic in flight at once (an append per hit inside this loop
                // put four dependent atomic round trips into every warp step: 19 -> 32 us)
                const unsigned m = __ballot_sync(FULL, has);
                if (m) {
                    const int cnt = __popc(m);
                    if (stage.n + cnt > kStageCap) flushRecs(a, stage, sDirLen, sDirAsc, nDirS, lane);
                    if (has) stage.buf[stage.n + __popc(m & ((1u << lane) - 1))] = rec;
                    stage.n += cnt;
                }
            } else {
                stage.push(has, rec, a.hits, &a.counters->nHits, a.hitCap, lane);
            }
        }
        sv0 = sv1;
        sv1 = sv2;
#pragma unroll
        for (int g = 0; g < G; g++) lit0[g] = lit1[g];
    }
    if (a.recKeys) flushRecs(a, stage, sDirLen, sDirAsc, nDirS, lane);
    else stage.flush(a.hits, &a.counters->nHits, a.hitCap, lane);
    if (lane == 0 && tests) atomicAdd(&a.counters->exactTests, tests);
    if (a.peerDone) {
        // Multi-GPU: the hits went straight into this rank's slot of rank 0's gather window (peer
        // stores).  The block that finishes last publishes the result: slot header, then the done
        // flag (2*seq: complete; 2*seq-1: a buffer overflowed, this rank is going to run again).
        // (barrier, then ONE fence by the thread that takes the ticket: the fence is cumulative over
        // what the barrier ordered before it -- the pattern of a grid-wide barrier)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system(); // this block's hits and counter updates are visible system-wide
            const unsigned int ticket = atomicAdd(a.peerTicket, 1u);
            if (ticket == gridDim.x - 1) {
                *a.peerTicket = 0u;
                __threadfence();
                const volatile Counters *c = a.counters;
                const unsigned int nHits = c->nHits;
                long long flags = nHits > a.hitCap ? 2 : 0;
                for (int g = 0; g < a.peerGroups && g < kMaxGroups; g++)
                    if (c->nSurvivors[g] > a.survCap) flags |= 1;
                a.peerHdr[0] = (long long)nHits;
                a.peerHdr[1] = flags;
                a.peerHdr[2] = (long long)c->exactTests;
                a.peerHdr[3] = (long long)a.peerSeq;
                for (int i = 4; i < 8; i++) a.peerHdr[i] = 0;
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t *>(a.peerDone) = flags ? 2u * a.peerSeq - 1u : 2u * a.peerSeq;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dense mode (bench only).  One warp per group of kDenseGroup clauses, lane = solver.  The
// clauses of a group advance together one literal at a time: all their T2 row gathers are issued
// back to back (kDenseGroup independent 256 B rows in flight per warp) before any result is
// consumed -- this kernel is bound by the gather stream, so memory-level parallelism is everything.
// ---------------------------------------------------------------------------------------------
constexpr int kDenseGroup = 16;

__global__ void __launch_bounds__(128) k_check_dense(CheckArgs a) {
    extern __shared__ int sTileEnd[];
    for (int i = threadIdx.x; i < a.nDir; i += blockDim.x) sTileEnd[i] = a.dir[i].tileEnd;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const int warp = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
    const int nWarps = gridDim.x * warpsPerBlock;
    const bool active = lane < a.groupSolvers;
    const int solver = a.groupBase + (active ? lane : 0);
    const uint32_t myStart = active ? a.params[solver].startVals : 0u;
    const uint2 *__restrict__ t2 = a.tables.t2 + solver;
    const size_t stride = (size_t)a.tables.solverStride;
    constexpr int kGroupsPerTile = kTileClauses / kDenseGroup;

    const long long nWork = (long long)a.totalTiles * kGroupsPerTile;
    for (long long w = warp; w < nWork; w += nWarps) {
        const int tile = (int)(w / kGroupsPerTile), q = (int)(w % kGroupsPerTile);
        const int k = findDir(sTileEnd, a.nDir, tile);
        const LenDir d = a.dir[k];
        const int tileInLen = tile - (k ? sTileEnd[k - 1] : 0);
        const int len = d.len;
        const int gTile = tileInLen + d.firstTile;
        const int c0 = gTile * kTileClauses + q * kDenseGroup;
        const int nValid = d.count - c0;
        if (nValid <= 0) continue;
        // lane l < kDenseGroup holds literal i of clause c0 + l
        const int32_t *col = d.base + (size_t)gTile * kTileClauses * len + tileSlot(q * kDenseGroup + (lane & (kDenseGroup - 1)));

        uint32_t all[kDenseGroup], one[kDenseGroup];
#pragma unroll
        for (int c = 0; c < kDenseGroup; c++) { all[c] = c < nValid ? myStart : 0u; one[c] = 0u; }
        int word = ldStream32(col);
        for (int i = 0; i < len; i++) {
            int nextWord = word;
            if (i + 1 < len) nextWord = ldStream32(col + (size_t)(i + 1) * kTileClauses);
            int lit[kDenseGroup];
            uint2 e[kDenseGroup];
#pragma unroll
            for (int c = 0; c < kDenseGroup; c++) {
                lit[c] = __shfl_sync(FULL, word, c);
                e[c] = ldTable(t2 + (size_t)(lit[c] >> 1) * stride);
            }
#pragma unroll
            for (int c = 0; c < kDenseGroup; c++)
                step(all[c], one[c], e[c].x & ((lit[c] & 1) ? e[c].y : ~e[c].y), ~e[c].x);
            word = nextWord;
        }
#pragma unroll
        for (int c = 0; c < kDenseGroup; c++) reportHits(a, all[c] | one[c], solver, len, c0 + c, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// Dense mode, L2-resident variant.  k_check_dense gathers 256-byte rows of a V x 32-solver table: 256 MB
// at 1 M variables, twice the L2, so every row comes from HBM (ncu round 1: 6.67 GB of DRAM reads per
// sweep against 0.43 GB algorithmic).  Here the table is cut into SLICES of 8 solvers -- slice g is a
// dense [var][8] array of {def,tru} pairs, 64 MB at 1 M variables -- and the clauses are swept once per
// slice (the literal rows are streamed evict-first, 176 MB per sweep), so the gathers of a sweep hit an
// L2-resident table and DRAM only sees the literal stream plus one read of the table: ~1 GB.  What is
// left is the L2 -> SM gather stream itself: 8 B per (literal, word) = 11.3 GB per pass over the database,
// whatever the layout (measured ceiling of 64-byte row gathers from a 64 MB table: ~10 TB/s, profiles/).
// A warp takes HALF a tile (64 clauses); lane = (c4 = lane / 8: one of four clauses, s = lane % 8: solver),
// so one gather instruction fetches four 64-byte rows; 16 gathers (= 64 clauses) are in flight per literal row.
// ---------------------------------------------------------------------------------------------
constexpr int kSliceSolvers = 8;

__global__ void __launch_bounds__(256) k_slice_tables(DeviceTables t, int groupBase, int groupSolvers, uint2 *__restrict__ sliced) {
    const int lane = threadIdx.x & 31; // solver of the group
    const size_t nVars = (size_t)t.varCap;
    for (size_t v = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); v < nVars; v += (size_t)gridDim.x * (blockDim.x >> 5)) {
        uint2 e = make_uint2(0u, 0u);
        if (lane < groupSolvers) e = t.t2[v * t.solverStride + groupBase + lane];
        sliced[((size_t)(lane / kSliceSolvers) * nVars + v) * kSliceSolvers + (lane % kSliceSolvers)] = e;
    }
}

__global__ void __launch_bounds__(128) k_check_dense_sliced(CheckArgs a, const uint2 *__restrict__ slice, int sliceBase) {
    extern __shared__ int sTileEnd[];
    for (int i = threadIdx.x; i < a.nDir; i += blockDim.x) sTileEnd[i] = a.dir[i].tileEnd;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int sIdx = lane & (kSliceSolvers - 1), c4 = lane >> 3;
    const int warpsPerBlock = blockDim.x >> 5;
    const int warp = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
    const int nWarps = gridDim.x * warpsPerBlock;
    const bool active = sliceBase + sIdx < a.groupSolvers;
    const int solver = a.groupBase + (active ? sliceBase + sIdx : 0);
    const uint32_t myStart = active ? a.params[solver].startVals : 0u;
    const uint2 *__restrict__ t2 = slice + sIdx;

    const long long nWork = (long long)a.totalTiles * 2;
    for (long long w = warp; w < nWork; w += nWarps) {
        const int tile = (int)(w >> 1), half = (int)(w & 1);
        const int k = findDir(sTileEnd, a.nDir, tile);
        const LenDir d = a.dir[k];
        const int len = d.len;
        const int gTile = tile - (k ? sTileEnd[k - 1] : 0) + d.firstTile;
        const int nValid = d.count - gTile * kTileClauses; // clauses of this tile
        if (nValid <= 64 * half) continue;
        // lane l holds, of every literal row, the words of clauses 32 * (2 half) + l and 32 * (2 half + 1) + l
        const int32_t *row = d.base + (size_t)gTile * kTileClauses * len + 4 * lane + 2 * half;

        uint32_t all[16], one[16];
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const int j = 32 * (2 * half + (c >> 3)) + 4 * (c & 7) + c4;
            all[c] = j < nValid ? myStart : 0u;
            one[c] = 0u;
        }
        int2 word = __ldcs(reinterpret_cast<const int2 *>(row));
        for (int i = 0; i < len; i++) {
            int2 nextWord = word;
            if (i + 1 < len) nextWord = __ldcs(reinterpret_cast<const int2 *>(row + (size_t)(i + 1) * kTileClauses));
            int lit[16];
            uint2 e[16];
#pragma unroll
            for (int c = 0; c < 16; c++) {
                lit[c] = __shfl_sync(FULL, (c >> 3) ? word.y : word.x, 4 * (c & 7) + c4);
                e[c] = ldTable(t2 + (size_t)(lit[c] >> 1) * kSliceSolvers);
            }
#pragma unroll
            for (int c = 0; c < 16; c++) step(all[c], one[c], e[c].x & ((lit[c] & 1) ? e[c].y : ~e[c].y), ~e[c].x);
            word = nextWord;
        }
#pragma unroll
        for (int c = 0; c < 16; c++)
            reportHits(a, all[c] | one[c], solver, len, gTile * kTileClauses + 32 * (2 * half + (c >> 3)) + 4 * (c & 7) + c4, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// Direct pipeline.  k_apply_direct = k_apply_updates with the deltas of solver s read from src[s]:
// the solver thread's own delta buffer in mapped pinned host memory (no staging copy on the host, no
// separate H2D: the transfer IS the kernel's load stream) or device memory.  Records are 12 bytes:
// a block pulls 256 of them as 768 coalesced words through shared memory, so PCIe sees full lines.
// A copy of every record goes to `keep` (the deferred collapse and re-timed launches read it).
// ---------------------------------------------------------------------------------------------
// Side jobs (ApplyExtra): this is the first kernel of a run, so it also (a) brings the run header -- directory,
// run parameters, delta pointers; ~10 KB -- from page-locked host memory into the slot's device copy, which
// the kernels that follow read (its own blocks read their parameters from the host copy), and (b) zeroes the
// run's counters.  That is one host-to-device copy and two memsets fewer to issue per run (~30 us of host time
// on the GPU boxes of this pool, where every asynchronous call costs 5-10 us).
__global__ void __launch_bounds__(256) k_apply_direct(const VarUpdate *const *__restrict__ src,
                                                      const SolverRunParams *__restrict__ params, DeviceTables t,
                                                      VarUpdate *__restrict__ keep, ApplyExtra x) {
    __shared__ uint32_t sAgg[kSlots], sSlot[kSlots];
    __shared__ __align__(16) uint32_t sWords[3 * 256];
    const int s = blockIdx.y;
    if (x.headWords | x.zeroAWords | x.zeroBWords) {
        const int nBlocks = gridDim.x * gridDim.y, me = blockIdx.y * gridDim.x + blockIdx.x;
        for (int i = me * blockDim.x + threadIdx.x; i < x.headWords; i += nBlocks * blockDim.x) x.headDev[i] = x.headHost[i];
        for (int i = me * blockDim.x + threadIdx.x; i < x.zeroBWords; i += nBlocks * blockDim.x) x.zeroB[i] = 0u;
        if (me == 0)
            for (int i = threadIdx.x; i < x.zeroAWords; i += blockDim.x) x.zeroA[i] = 0u;
    }
    const SolverRunParams &p = params[s];
    const int n = p.updCount, nGroups = p.nGroups, updStart = p.updStart;
    if (n <= 0) return;
    if (threadIdx.x < kSlots) {
        sAgg[threadIdx.x] = p.groupAggBit[threadIdx.x];
        sSlot[threadIdx.x] = p.groupSlotMask[threadIdx.x];
    }
    const uint32_t used = p.usedAggBits;
    const uint32_t *__restrict__ w = reinterpret_cast<const uint32_t *>(src[s]);
    uint32_t *__restrict__ kw = reinterpret_cast<uint32_t *>(keep + updStart);
    for (int c0 = blockIdx.x * 256; c0 < n; c0 += gridDim.x * 256) {
        const int cnt = min(256, n - c0), nw = 3 * cnt;
        __syncthreads();
        // (a chunk starts 3072 * k bytes into the solver's buffer: 16-byte loads whenever the buffers are
        // aligned -- they are, page-locked allocations -- so that PCIe sees 512-byte requests per warp)
        const uint32_t *cw = w + (size_t)3 * c0;
        uint32_t *ck = kw + (size_t)3 * c0;
        const int nv = ((uintptr_t)cw & 15) == 0 ? nw >> 2 : 0;
        for (int k = threadIdx.x; k < nv; k += 256)
            *reinterpret_cast<uint4 *>(sWords + 4 * k) = *reinterpret_cast<const uint4 *>(cw + 4 * k);
        for (int k = 4 * nv + threadIdx.x; k < nw; k += 256) sWords[k] = cw[k];
        __syncthreads();
        for (int k = threadIdx.x; k < nw; k += 256) ck[k] = sWords[k]; // (the kept copy starts wherever the solver's share starts)
        if ((int)threadIdx.x < cnt) {
            VarUpdate vu;
            vu.var = (int32_t)sWords[3 * threadIdx.x];
            vu.def = sWords[3 * threadIdx.x + 1];
            vu.tru = sWords[3 * threadIdx.x + 2];
            t.t2[(size_t)vu.var * t.solverStride + s] = make_uint2(vu.def, vu.tru);
            if (used) {
                uint32_t bt = vu.tru & vu.def, bf = ~vu.tru & vu.def, bu = ~vu.def;
                uint32_t T = 0, F = 0, U = 0;
                for (int g = 0; g < nGroups; g++) {
                    uint32_t m = sSlot[g], bit = sAgg[g];
                    if (bt & m) T |= bit;
                    if (bf & m) F |= bit;
                    if (bu & m) U |= bit;
                }
                writeAggregates(t, s, vu.var, used, T, F, U);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_emit_sort + k_emit_write.  k_exact appended every solver's hits to kRecShards sub-lists (so that
// a solver's appends are spread over several counters).  k_emit_sort, one block per solver: pulls the
// sub-lists together, sorts the records by (length, index) -- the reproducible hand-over order -- with
// a bitonic network (shared memory; lists that do not fit are compacted and sorted in place in global
// memory by the same code), turns the lengths into literal positions and leaves the sorted list on
// the device.  k_emit_write, many blocks per solver (PCIe needs many SMs writing): clause ids,
// positions and the literal stream of every solver go straight into the run's result buffer in
// mapped page-locked host memory, coalesced.  The host builds the solver's ClauseBatch as a view
// over that memory: no device-side global sort, no D2H copy of unknown size, no host-side literal
// copies (reference: Reporter.cuh:103-126 + Reported.cu:160-204).  The block that finishes last
// writes the header and, after a system fence, the run's sequence number.
// ---------------------------------------------------------------------------------------------
constexpr int kWriteChunk = 256;    // entries per block step of k_emit_write


constexpr int kSortWarps = 8;        // buckets per block of k_emit_sort (a warp each)
constexpr int kSortSmemRecs = 256;   // records per bucket sorted in shared memory

// k_emit_scan, one block per solver, a thread per bucket: where every bucket's list starts in its solver's sorted list
// and literal stream (exclusive prefix over the solver's bucket counters); the block that finishes last adds up the
// solvers (where a solver starts in the run's streams, does it fit the result buffer: EmitSolver) and the overflow
// flags.  (Round-2 profile: with every block of k_emit_sort adding up "all the counters before mine" that kernel read
// 4.3 M sectors and took 34 us; a single-block scan over all the counters took 26 us -- three dependent global round
// trips by one block.)
__global__ void __launch_bounds__(kRecBuckets) k_emit_scan(EmitArgs a) {
    __shared__ long long sE[kRecBuckets / 32], sL[kRecBuckets / 32];
    __shared__ unsigned int sMax[kRecBuckets / 32], sOver[kRecBuckets / 32];
    __shared__ bool sLast;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned int bucketCap = a.recCap / kRecBuckets;
    long long *const solverTot = a.bucketBase + 2 * (size_t)a.nSolvers * kRecBuckets; // [solver][2]
    const size_t me = (size_t)s * kRecBuckets + tid;
    const unsigned long long ct = a.solverCtr[me * kCtrStride];
    const unsigned int nRaw = (unsigned int)ct;
    const long long e = min(nRaw, bucketCap), l = (long long)(ct >> 32);
    auto blockExclusive = [&](long long ve, long long vl, long long &outE, long long &outL, long long &totE, long long &totL) {
        long long ie = ve, il = vl;
        for (int o = 1; o < 32; o <<= 1) {
            const long long xe = __shfl_up_sync(FULL, ie, o), xl = __shfl_up_sync(FULL, il, o);
            if (lane >= o) { ie += xe; il += xl; }
        }
        __syncthreads();
        if (lane == 31) { sE[wid] = ie; sL[wid] = il; }
        __syncthreads();
        long long be = 0, bl = 0;
        totE = totL = 0;
        for (int w = 0; w < kRecBuckets / 32; w++) {
            if (w < wid) { be += sE[w]; bl += sL[w]; }
            totE += sE[w];
            totL += sL[w];
        }
        outE = be + ie - ve;
        outL = bl + il - vl;
    };
    long long exE, exL, totE, totL;
    blockExclusive(e, l, exE, exL, totE, totL);
    a.bucketBase[2 * me] = exE;
    a.bucketBase[2 * me + 1] = exL;
    const unsigned int mx = __reduce_max_sync(FULL, nRaw), over = __reduce_or_sync(FULL, nRaw > bucketCap ? 1u : 0u);
    if (lane == 0) { sMax[wid] = mx; sOver[wid] = over; }
    __syncthreads();
    if (tid == 0) {
        unsigned int m = 0, o = 0;
        for (int w = 0; w < kRecBuckets / 32; w++) { m = max(m, sMax[w]); o |= sOver[w]; }
        if (o) atomicOr(a.ticket + 1, 2u);
        atomicMax(a.ticket + 2, m); // (x kRecBuckets = what recCap would have had to be)
        solverTot[2 * s] = totE;
        solverTot[2 * s + 1] = totL;
        a.recPos[(size_t)s * (a.recCap + 1) + totE] = (int32_t)totL; // the position behind the solver's last entry
        __threadfence();
        sLast = atomicAdd(a.ticket + 3, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    // the last block: the solvers one after the other in the run's entry and literal streams
    long long carryE = 0, carryL = 0;
    for (int base = 0; base < a.nSolvers; base += kRecBuckets) {
        const int sv = base + tid;
        const long long nS = sv < a.nSolvers ? __ldcg(solverTot + 2 * sv) : 0, litsS = sv < a.nSolvers ? __ldcg(solverTot + 2 * sv + 1) : 0;
        long long eS, lS, tE, tL;
        blockExclusive(nS, litsS, eS, lS, tE, tL);
        eS += carryE;
        lS += carryL;
        if (sv < a.nSolvers) {
            EmitSolver &es = a.solverInfo[sv];
            es.entryBase = eS;
            es.litBase = lS;
            es.nLits = (int32_t)litsS;
            es.nSorted = (uint32_t)nS;
            es.n = (eS + nS <= a.entryCap && lS + litsS <= a.litCap) ? (int32_t)nS : -1;
            if (es.n < 0) atomicOr(a.ticket + 1, 4u);
        }
        carryE += tE;
        carryL += tL;
    }
    if (tid == 0) a.ticket[3] = 0u;
}

// block bx of solver s sorts the solver's buckets bx * kSortWarps .. + kSortWarps - 1, a warp each
__device__ __forceinline__ void emitSortBuckets(const EmitArgs &a, int s, int bx, unsigned long long (*sK)[kSortSmemRecs],
                                                uint32_t (*sM)[kSortSmemRecs]) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int bucket = bx * kSortWarps + wid;
    const unsigned int bucketCap = a.recCap / kRecBuckets;
    auto ctrOf = [&](int solver, int b) { return a.solverCtr[((size_t)solver * kRecBuckets + b) * kCtrStride]; };
    // where this bucket starts in the solver's sorted list and in its literal stream (k_emit_scan)
    const size_t me = (size_t)s * kRecBuckets + bucket;
    const long long eBucket = a.bucketBase[2 * me], lBucket = a.bucketBase[2 * me + 1];
    const unsigned long long mine = ctrOf(s, bucket);
    const unsigned int nRaw = (unsigned int)mine, n = min(nRaw, bucketCap);

    // sort this bucket by (length, index): bitonic network, one warp; in shared memory when it fits,
    // else in place in the bucket's own storage (P <= bucketCap, a power of two)
    unsigned int P = 1;
    while (P < n) P <<= 1;
    unsigned long long *const gK = a.recKeys + (size_t)s * a.recCap + (size_t)bucket * bucketCap;
    uint32_t *const gM = a.recMasks + (size_t)s * a.recCap + (size_t)bucket * bucketCap;
    unsigned long long *K;
    uint32_t *M;
    if (P <= (unsigned int)kSortSmemRecs) {
        K = sK[wid];
        M = sM[wid];
        for (unsigned int i = lane; i < P; i += 32) {
            K[i] = i < n ? gK[i] : ~0ull;
            M[i] = i < n ? gM[i] : 0u;
        }
    } else {
        K = gK;
        M = gM;
        for (unsigned int i = n + lane; i < P; i += 32) { K[i] = ~0ull; M[i] = 0u; }
    }
    __syncwarp();
    for (unsigned int k = 2; k <= P; k <<= 1)
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            for (unsigned int i = lane; i < P; i += 32) {
                const unsigned int x = i ^ j;
                if (x > i) {
                    const unsigned long long ki = K[i], kx = K[x];
                    if ((ki > kx) == ((i & k) == 0)) {
                        K[i] = kx; K[x] = ki;
                        const uint32_t mi = M[i]; M[i] = M[x]; M[x] = mi;
                    }
                }
            }
            __syncwarp();
        }
    // the bucket's place in the solver's sorted list: records, masks, literal positions
    unsigned long long *const oK = a.sortKeys + (size_t)s * a.recCap + eBucket;
    uint32_t *const oM = a.sortMasks + (size_t)s * a.recCap + eBucket;
    int32_t *const oPos = a.recPos + (size_t)s * (a.recCap + 1) + eBucket;
    long long run = lBucket;
    for (unsigned int i0 = 0; i0 < n; i0 += 32) {
        const unsigned int i = i0 + lane;
        const unsigned long long key = i < n ? K[i] : 0ull;
        const long long len = i < n ? (long long)(key >> 32) : 0;
        long long incl = len;
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (i < n) {
            oK[i] = key;
            oM[i] = M[i];
            oPos[i] = (int32_t)(run + incl - len);
        }
        run += __shfl_sync(FULL, incl, 31);
    }
}

// grid = (kRecBuckets / kSortWarps, solvers)
__global__ void __launch_bounds__(kSortWarps * 32) k_emit_sort(EmitArgs a) {
    __shared__ unsigned long long sK[kSortWarps][kSortSmemRecs];
    __shared__ uint32_t sM[kSortWarps][kSortSmemRecs];
    emitSortBuckets(a, blockIdx.y, blockIdx.x, sK, sM);
}

// block bx of the nbx blocks of solver s: the blocks of a solver take chunks of kWriteChunk entries in turn; nBlocks = all
// the blocks of the launch (the one that finishes last writes the header)
__device__ __forceinline__ void emitWriteSolver(const EmitArgs &a, int s, int bx, int nbx, unsigned int nBlocks, int *sLen, int32_t *sPos) {
    const int tid = threadIdx.x;
    const int nDir = a.nDir;
    for (int i = tid; i < min(nDir, kDirCache); i += blockDim.x) sLen[i] = a.dir[i].len;
    const EmitSolver es = a.solverInfo[s];
    const unsigned long long *__restrict__ K = a.sortKeys + (size_t)s * a.recCap;
    const int32_t *__restrict__ gPos = a.recPos + (size_t)s * (a.recCap + 1);
    const int n = es.n; // -1: this solver does not fit
    for (int c0 = bx * kWriteChunk; c0 < n; c0 += nbx * kWriteChunk) {
        const int cnt = min(kWriteChunk, n - c0);
        __syncthreads();
        for (int i = tid; i <= cnt; i += blockDim.x) sPos[i] = __ldcg(gPos + c0 + i); // (ld.cg: k_emit_fused reads what other blocks just wrote)
        __syncthreads();
        if (tid < cnt) {
            const unsigned long long key = __ldcg(K + c0 + tid);
            a.ids[es.entryBase + c0 + tid] = a.dir[dirIndexOfLen(a.dir, sLen, nDir, (int)(key >> 32))].ids[(unsigned int)key];
            if (a.keysOut) a.keysOut[es.entryBase + c0 + tid] = key;
            if (a.masksOut) a.masksOut[es.entryBase + c0 + tid] = __ldcg(a.sortMasks + (size_t)s * a.recCap + c0 + tid);
        }
        for (int i = tid; i < cnt + (c0 + cnt == n ? 1 : 0); i += blockDim.x) a.pos[es.entryBase + s + c0 + i] = sPos[i];
        // Literal stream of these entries.  PCIe wants full, aligned lines: the body goes out as 16-byte
        // vectors aligned in the HOST buffer (thread = four consecutive literals), the unaligned head and
        // tail of the chunk as single words.
        const int q0 = sPos[0], q1 = sPos[cnt];
        auto litAt = [&](int q) -> int32_t {
            int lo = 0, hi = cnt; // last i with sPos[i] <= q
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sPos[mid] <= q) lo = mid; else hi = mid;
            }
            const unsigned long long key = __ldcg(K + c0 + lo);
            const int len = (int)(key >> 32), idx = (int)(unsigned int)key, j = q - sPos[lo];
            const int32_t *src = a.dir[dirIndexOfLen(a.dir, sLen, nDir, len)].base + (size_t)(idx / kTileClauses) * kTileClauses * len +
                                 tileSlot(idx % kTileClauses);
            return __ldg(src + (size_t)j * kTileClauses);
        };
        const long long A0 = es.litBase + q0, A1 = es.litBase + q1; // absolute positions in the buffer's literal stream
        const long long B0 = min(A1, (A0 + 3) & ~3ll), B1 = max(B0, A1 & ~3ll);
        for (long long A = A0 + tid; A < B0; A += blockDim.x) a.lits[A] = litAt((int)(A - es.litBase));
        for (long long A = B0 + 4ll * tid; A < B1; A += 4ll * blockDim.x) {
            const int q = (int)(A - es.litBase);
            int4 v;
            v.x = litAt(q);
            v.y = litAt(q + 1);
            v.z = litAt(q + 2);
            v.w = litAt(q + 3);
            *reinterpret_cast<int4 *>(a.lits + A) = v;
        }
        for (long long A = B1 + tid; A < A1; A += blockDim.x) a.lits[A] = litAt((int)(A - es.litBase));
    }
    if (n == 0 && bx == 0 && tid == 0) a.pos[es.entryBase + s] = 0;
    if (bx == 0 && tid == 0) {
        a.hdr->solver[s].entryBase = es.entryBase;
        a.hdr->solver[s].litBase = es.litBase;
        a.hdr->solver[s].n = max(es.n, 0);
        a.hdr->solver[s].nLits = es.nLits;
    }
    // Completion is the kernel's: the host waits for the event behind it, which makes every store to
    // host memory visible.  (A system-scope fence per block -- needed only if the host polled the header
    // while the kernel runs -- drains the block's posted PCIe writes: 1184 of them cost ~200 us.)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(a.ticket, 1u);
        if (ticket == nBlocks - 1) {
            __threadfence();
            const volatile Counters *cn = a.counters;
            unsigned int flags = *reinterpret_cast<volatile unsigned int *>(a.ticket + 1);
            for (int g = 0; g < a.groups && g < kMaxGroups; g++) {
                a.hdr->nSurvivors[g] = cn->nSurvivors[g];
                if (cn->nSurvivors[g] > a.survCap) flags |= 1u;
            }
            long long tot = 0, lt = 0;
            for (int t = 0; t < a.nSolvers; t++) {
                tot += a.solverInfo[t].nSorted;
                lt += a.solverInfo[t].nLits;
            }
            a.hdr->nTotal = tot;
            a.hdr->litTotal = lt;
            a.hdr->exactTests = cn->exactTests;
            a.hdr->maxRec = *reinterpret_cast<volatile unsigned int *>(a.ticket + 2); // largest bucket
            a.hdr->flags = flags;
            a.ticket[0] = 0u;
            a.ticket[1] = 0u;
            a.ticket[2] = 0u;
            a.ticket[3] = 0u; // (k_emit_fused: the start-order counter)
            if (a.solverDone)
                for (int t = 0; t < a.nSolvers; t++) a.solverDone[t] = 0u;
            a.hdr->seq = a.seq;
        }
    }
}

// grid = (blocks per solver, solvers)
__global__ void __launch_bounds__(256) k_emit_write(EmitArgs a) {
    __shared__ int sLen[kDirCache];
    __shared__ int32_t sPos[kWriteChunk + 1];
    emitWriteSolver(a, blockIdx.y, blockIdx.x, gridDim.x, gridDim.x * gridDim.y, sLen, sPos);
}

// k_emit_fused = k_emit_sort + k_emit_write in one launch, so that the PCIe writes of the first solvers run while the
// last buckets are still being sorted (sort ~20 us, write ~70 us back to back).  kRecBuckets / kSortWarps = 32 blocks per
// solver: a block sorts its 8 buckets, then waits for the solver's other blocks and takes its share of the solver's
// entries.  Waiting on other blocks is safe because a block's place is its START order (a ticket), not its block index:
// the blocks of a solver are then the 32 that started one after the other, at most one solver is ever started in part,
// and every other resident block belongs to a solver whose blocks have all started and therefore finish.
__global__ void __launch_bounds__(256) k_emit_fused(EmitArgs a) {
    __shared__ unsigned long long sK[kSortWarps][kSortSmemRecs];
    __shared__ uint32_t sM[kSortWarps][kSortSmemRecs];
    __shared__ int sLen[kDirCache];
    __shared__ int32_t sPos[kWriteChunk + 1];
    __shared__ unsigned int sPlace;
    constexpr int kPerSolver = kRecBuckets / kSortWarps;
    if (threadIdx.x == 0) sPlace = atomicAdd(a.ticket + 3, 1u);
    __syncthreads();
    const int s = (int)(sPlace / kPerSolver), bx = (int)(sPlace % kPerSolver);
    emitSortBuckets(a, s, bx, sK, sM);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence(); // this block's part of the sorted list is visible before it says so
        atomicAdd(a.solverDone + s, 1u);
        while (*reinterpret_cast<volatile unsigned int *>(a.solverDone + s) < (unsigned int)kPerSolver) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
    emitWriteSolver(a, s, bx, kPerSolver, gridDim.x, sLen, sPos);
}

// activity bumps straight from the sorted per-solver record lists (blockIdx.y = solver)
__global__ void k_bump_recs(const unsigned long long *__restrict__ recKeys, unsigned int recCap,
                            const EmitSolver *__restrict__ solverInfo, const LenDir *__restrict__ dir, int nDir, float inc,
                            int *overflow) {
    const int s = blockIdx.y;
    const unsigned int n = solverInfo[s].nSorted;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = recKeys[(size_t)s * recCap + i];
        const int len = (int)(key >> 32), idx = (int)(unsigned int)key;
        int lo = 0, hi = nDir - 1; // descending length
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (dir[mid].len <= len) hi = mid; else lo = mid + 1;
        }
        float old = atomicAdd(dir[lo].acts + idx, inc);
        if (old + inc > 1e19f) *overflow = 1;
    }
}

// 8 independent LOP3 chains per thread, 0xE8 = majority so nothing folds away
__global__ void __launch_bounds__(256) k_lop3_peak(uint32_t *out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3u + 1u, a2 = a0 * 5u + 2u, a3 = a0 * 7u + 3u;
    uint32_t a4 = a0 * 11u + 4u, a5 = a0 * 13u + 5u, a6 = a0 * 17u + 6u, a7 = a0 * 19u + 7u;
    uint32_t b = blockIdx.x * 0x9E3779B9u + 12345u, c = ~b * 0x85EBCA6Bu;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a0) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a1) : "r"(c), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a2) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a3) : "r"(c), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a4) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a5) : "r"(c), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a6) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a7) : "r"(c), "r"(b));
        }
    }
    uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x12345678u) out[0] = r; // never true in practice; keeps the chains live
}

__global__ void k_finalize(const Counters *c, unsigned int hitCap, unsigned int survCap, int groups, long long *dst) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bool over = c->nHits > hitCap;
    for (int g = 0; g < groups && g < kMaxGroups; g++) over = over || c->nSurvivors[g] > survCap;
    dst[0] = (long long)c->nHits;
    dst[1] = over ? 1 : 0;
    for (int i = 2; i < 8; i++) dst[i] = 0;
}

// ---- peer-memory exchange (multi-GPU, peer.cu) ----
// root: push the batch into every worker's window with plain (posted) peer stores and then store the
// sequence number into every mailbox -- one launch; the block that finishes last signals.  (A copy-
// engine transfer costs ~17 us of fixed latency per peer for these 2 MB payloads; SM stores do not.)
__global__ void __launch_bounds__(256) k_peer_push(const uint4 *__restrict__ src, long long bytes, PeerPushList L, uint32_t seq,
                                                   unsigned int *ticket) {
    const long long n16 = (bytes + 15) / 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int r = 0; r < L.n; r++) L.dst[r][i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system(); // this block's stores have been performed at the peers (cumulative over the barrier)
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence_system();
            for (int r = 0; r < L.n; r++) *reinterpret_cast<volatile uint32_t *>(L.mailbox[r]) = seq;
        }
    }
}

// root, direct variant: the deltas do not pass through a staging buffer and a host-to-device copy first.  Row s
// of the grid reads solver s's records where the solver thread left them (src[s]: page-locked host memory, as
// in k_apply_direct) and stores them -- ONE pass over PCIe -- into this rank's payload area and into every
// worker's window; row nSolvers forwards the prefix ([header][run parameters], uploaded with the run header).
// The block that finishes last signals the mailboxes.
__global__ void __launch_bounds__(256) k_peer_push_direct(const VarUpdate *const *__restrict__ src,
                                                          const SolverRunParams *__restrict__ params, int nSolvers,
                                                          uint32_t *__restrict__ local, const uint32_t *__restrict__ prefix,
                                                          long long prefixWords, PeerPushList L, uint32_t seq, unsigned int *ticket) {
    const long long stride = (long long)gridDim.x * blockDim.x, first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)blockIdx.y == nSolvers) {
        for (long long i = first; i < prefixWords; i += stride) {
            const uint32_t v = prefix[i];
            local[i] = v;
            for (int r = 0; r < L.n; r++) reinterpret_cast<uint32_t *>(L.dst[r])[i] = v;
        }
    } else {
        const SolverRunParams &p = params[blockIdx.y];
        const long long nw = 3ll * p.updCount, base = prefixWords + 3ll * p.updStart;
        const uint32_t *__restrict__ w = reinterpret_cast<const uint32_t *>(src[blockIdx.y]);
        const bool inPlace = w == local + base; // (a delta buffer that was not page-locked went up as an ordinary copy)
        for (long long k = first; k < nw; k += stride) {
            const uint32_t v = w[k];
            if (!inPlace) local[base + k] = v;
            for (int r = 0; r < L.n; r++) reinterpret_cast<uint32_t *>(L.dst[r])[base + k] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x * gridDim.y - 1) {
            *ticket = 0u;
            __threadfence_system();
            for (int r = 0; r < L.n; r++) *reinterpret_cast<volatile uint32_t *>(L.mailbox[r]) = seq;
        }
    }
}

// every rank, when no k_exact launch could publish the result itself (see k_exact): result header into its slot of the root's gather window, then
// the done flag.  The hits themselves were appended to the slot by k_exact (peer stores), which has
// completed.  flag = 2*seq when the result is complete, 2*seq-1 when a survivor / hit buffer
// overflowed and this rank is going to run again.
__global__ void k_peer_finalize(const Counters *c, unsigned int hitCap, unsigned int survCap, int groups, long long *hdr,
                                uint32_t *doneFlag, uint32_t seq) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long flags = c->nHits > hitCap ? 2 : 0;
    for (int g = 0; g < groups && g < kMaxGroups; g++)
        if (c->nSurvivors[g] > survCap) flags |= 1;
    hdr[0] = (long long)c->nHits;
    hdr[1] = flags;
    hdr[2] = (long long)c->exactTests;
    hdr[3] = (long long)seq;
    for (int i = 4; i < 8; i++) hdr[i] = 0;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(doneFlag) = flags ? 2u * seq - 1u : 2u * seq;
}

// fallback for drivers without stream memory operations: poll a flag in this device's memory
// (bounded: *err is set after ~timeoutNs instead of hanging the device)
__global__ void k_peer_wait(const uint32_t *flag, uint32_t value, unsigned long long timeoutNs, int *err) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const uint32_t v = *reinterpret_cast<const volatile uint32_t *>(flag);
        if ((int32_t)(v - value) >= 0) return;
        __nanosleep(200);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeoutNs) { *err = 1; return; }
    }
}

// rank 0: wait for the done flags of ALL workers with one launch (thread r polls flag r).  Seven
// stream memory operations in a row cost ~2.5 us each even when the flags are already set.
__global__ void k_peer_wait_all(PeerFlagList flags, uint32_t value, unsigned long long timeoutNs, int *err) {
    const int r = threadIdx.x;
    if (r >= flags.n) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const uint32_t v = *reinterpret_cast<const volatile uint32_t *>(flags.p[r]);
        if ((int32_t)(v - value) >= 0) return;
        __nanosleep(100);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeoutNs) { *err = 1; return; }
    }
}

// ---- post-processing of large hit lists ----
// key = solver | length | index packed into the fewest bits the database needs: the radix sort makes
// one pass per 8 key bits
__global__ void k_post_keys(const HitRecord *__restrict__ hits, unsigned int n, unsigned long long *keys, unsigned int *vals,
                            int lenBits, int idxBits) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const HitRecord h = hits[i];
    keys[i] = ((unsigned long long)(unsigned int)h.solver << (lenBits + idxBits)) |
              ((unsigned long long)(unsigned int)h.len << idxBits) | (unsigned long long)(unsigned int)h.idx;
    vals[i] = i;
}

__global__ void k_post_lens(const HitRecord *__restrict__ hits, const unsigned int *__restrict__ order, unsigned int n,
                            long long *lens) {
    unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) lens[j] = hits[order[j]].len;
    if (j == n) lens[j] = 0;
}

// one warp per hit (in sorted order): record + literals
__global__ void __launch_bounds__(256) k_post_emit(const HitRecord *__restrict__ hits, const unsigned int *__restrict__ order,
                                                   unsigned int n, const LenDir *__restrict__ dir, int nDir,
                                                   const long long *__restrict__ litPos, SortedHit *__restrict__ out,
                                                   int32_t *__restrict__ lits, long long litCap) {
    const int lane = threadIdx.x & 31;
    const unsigned int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nWarps) {
        const HitRecord h = hits[order[j]];
        // the directory is sorted by descending length
        int lo = 0, hi = nDir - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (dir[mid].len <= h.len) hi = mid; else lo = mid + 1;
        }
        const LenDir d = dir[lo];
        const long long pos = litPos[j];
        const int tile = h.idx / kTileClauses;
        const int32_t *src = d.base + (size_t)tile * kTileClauses * h.len + tileSlot(h.idx % kTileClauses);
        for (int i = lane; i < h.len; i += 32)
            if (pos + i < litCap) lits[pos + i] = __ldg(src + (size_t)i * kTileClauses);
        if (lane == 0) out[j] = SortedHit{h.mask, h.solver, h.len, h.idx, d.ids[h.idx], pos};
    }
}

__global__ void k_bump_activity(const unsigned char *__restrict__ recs, int stride, unsigned int n,
                                const LenDir *__restrict__ dir, int nDir, float inc, int *overflow) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int len = *reinterpret_cast<const int *>(recs + (size_t)i * stride + 8);
    const int idx = *reinterpret_cast<const int *>(recs + (size_t)i * stride + 12);
    int lo = 0, hi = nDir - 1; // descending length
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (dir[mid].len <= len) hi = mid; else lo = mid + 1;
    }
    float old = atomicAdd(dir[lo].acts + idx, inc);
    if (old + inc > 1e19f) *overflow = 1;
}

__global__ void k_scale_acts(float *acts, long long n, float factor) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acts[i] *= factor;
}

int resolveBlocks(const void *kernel, int threads, size_t smem, int numSMs, int requested, long long work) {
    // (the occupancy query is a driver call: asked once per kernel / shape, not once per launch)
    struct Seen {
        const void *kernel;
        int threads;
        size_t smem;
        int perSM;
    };
    static thread_local Seen seen[16];
    static thread_local int nSeen = 0;
    int perSM = 0;
    for (int i = 0; i < nSeen; i++)
        if (seen[i].kernel == kernel && seen[i].threads == threads && seen[i].smem == smem) perSM = seen[i].perSM;
    if (perSM == 0) {
        perSM = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, threads, smem);
        if (perSM < 1) perSM = 1;
        if (nSeen < 16) seen[nSeen++] = Seen{kernel, threads, smem, perSM};
    }
    long long blocks = requested > 0 ? requested : (long long)perSM * numSMs;
    if (work >= 0 && blocks > work) blocks = work;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

inline void checkLaunch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) GSS_DIE(std::string("kernel launch failed: ") + what + ": " + cudaGetErrorString(e));
}

} // namespace

double measureLop3Peak(int numSMs, cudaStream_t s, int64_t *launches) {
    uint32_t *d = nullptr;
    GSS_CUDA(cudaMalloc(&d, 4));
    const int iters = 2048, blocks = numSMs * 8, threads = 256;
    cudaEvent_t e0, e1;
    GSS_CUDA(cudaEventCreate(&e0));
    GSS_CUDA(cudaEventCreate(&e1));
    k_lop3_peak<<<blocks, threads, 0, s>>>(d, 64); // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        GSS_CUDA(cudaEventRecord(e0, s));
        k_lop3_peak<<<blocks, threads, 0, s>>>(d, iters);
        GSS_CUDA(cudaEventRecord(e1, s));
        GSS_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        GSS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0;
        best = std::max(best, ops / (ms * 1e-3));
        *launches += 1;
    }
    checkLaunch("k_lop3_peak");
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

size_t postprocessTempBytes(unsigned int n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (unsigned int *)nullptr, (unsigned int *)nullptr, (int)n, 0, 56);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (long long *)nullptr, (long long *)nullptr, (int)n + 1);
    return std::max(a, b) + 256;
}

void launchPostSort(const HitRecord *hits, unsigned int n, const PostBuffers &b, int solverBits, int lenBits, int idxBits,
                    cudaStream_t s, int64_t *launches) {
    const unsigned int blocks = (n + 1 + 255) / 256;
    k_post_keys<<<blocks, 256, 0, s>>>(hits, n, b.keysIn, b.valsIn, lenBits, idxBits);
    size_t tb = b.tempBytes;
    cub::DeviceRadixSort::SortPairs(b.temp, tb, b.keysIn, b.keysOut, b.valsIn, b.valsOut, (int)n, 0,
                                    std::min(64, solverBits + lenBits + idxBits), s);
    // literal positions: exclusive sum of the lengths in sorted order (litPos doubles as the input)
    k_post_lens<<<blocks, 256, 0, s>>>(hits, b.valsOut, n, b.litPos);
    tb = b.tempBytes;
    cub::DeviceScan::ExclusiveSum(b.temp, tb, b.litPos, b.litPos, (int)n + 1, s);
    checkLaunch("post sort");
    *launches += 4;
}

void launchPostEmit(const HitRecord *hits, unsigned int n, const LenDir *dir, int nDir, const PostBuffers &b,
                    cudaStream_t s, int64_t *launches) {
    unsigned int blocks = std::min<unsigned int>((n + 7) / 8, 148 * 8);
    k_post_emit<<<blocks, 256, 0, s>>>(hits, b.valsOut, n, dir, nDir, b.litPos, b.sorted, b.lits, b.litCap);
    checkLaunch("k_post_emit");
    ++*launches;
}

void launchBumpActivity(const void *recs, int strideBytes, unsigned int n, const LenDir *dir, int nDir, float inc,
                        int *overflow, cudaStream_t s, int64_t *launches) {
    if (n == 0) return;
    k_bump_activity<<<(n + 255) / 256, 256, 0, s>>>((const unsigned char *)recs, strideBytes, n, dir, nDir, inc, overflow);
    checkLaunch("k_bump_activity");
    ++*launches;
}

void scaleActivitiesOnDevice(float *acts, int64_t n, float factor, cudaStream_t stream) {
    if (n <= 0) return;
    k_scale_acts<<<(unsigned int)((n + 255) / 256), 256, 0, stream>>>(acts, (long long)n, factor);
    checkLaunch("k_scale_acts");
}

void launchFinalize(const Counters *counters, unsigned int hitCap, unsigned int survCap, int groups, long long *dstHeader,
                    cudaStream_t s, int64_t *launches) {
    k_finalize<<<1, 32, 0, s>>>(counters, hitCap, survCap, groups, dstHeader);
    checkLaunch("k_finalize");
    ++*launches;
}

void launchPeerPush(const void *src, long long bytes, const PeerPushList &L, uint32_t seq, unsigned int *ticket, int numSMs,
                    cudaStream_t s, int64_t *launches) {
    if (L.n == 0) return;
    k_peer_push<<<numSMs * 2, 256, 0, s>>>(static_cast<const uint4 *>(src), bytes, L, seq, ticket);
    checkLaunch("k_peer_push");
    ++*launches;
}

void launchPeerPushDirect(const VarUpdate *const *src, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver, void *local,
                          const void *prefix, long long prefixWords, const PeerPushList &L, uint32_t seq, unsigned int *ticket,
                          int numSMs, cudaStream_t s, int64_t *launches) {
    // enough blocks per solver to keep PCIe full (every block has 3 KB in flight), few enough to stay in one wave
    const int perSolver = std::max(1, std::min((3 * maxUpdPerSolver + 767) / 768, std::max(1, numSMs * 8 / (nSolvers + 1))));
    k_peer_push_direct<<<dim3((unsigned int)perSolver, (unsigned int)nSolvers + 1), 256, 0, s>>>(
        src, params, nSolvers, static_cast<uint32_t *>(local), static_cast<const uint32_t *>(prefix), prefixWords, L, seq, ticket);
    checkLaunch("k_peer_push_direct");
    ++*launches;
}

void launchPeerFinalize(const Counters *counters, unsigned int hitCap, unsigned int survCap, int groups, long long *hdr,
                        uint32_t *doneFlag, uint32_t seq, cudaStream_t s, int64_t *launches) {
    k_peer_finalize<<<1, 32, 0, s>>>(counters, hitCap, survCap, groups, hdr, doneFlag, seq);
    checkLaunch("k_peer_finalize");
    ++*launches;
}

void launchPeerWaitAll(const PeerFlagList &flags, uint32_t value, unsigned long long timeoutNs, int *err, cudaStream_t s,
                       int64_t *launches) {
    if (flags.n == 0) return;
    k_peer_wait_all<<<1, 32, 0, s>>>(flags, value, timeoutNs, err);
    checkLaunch("k_peer_wait_all");
    ++*launches;
}

void launchPeerWait(const uint32_t *flag, uint32_t value, unsigned long long timeoutNs, int *err, cudaStream_t s,
                    int64_t *launches) {
    k_peer_wait<<<1, 32, 0, s>>>(flag, value, timeoutNs, err);
    checkLaunch("k_peer_wait");
    ++*launches;
}

void launchFillTables(const DeviceTables &t, int varFrom, cudaStream_t s, int64_t *launches) {
    if (t.varCap <= varFrom) return;
    k_fill_tables<<<592, 256, 0, s>>>(t, varFrom);
    checkLaunch("k_fill_tables");
    ++*launches;
}

static dim3 updateGrid(int nSolvers, int maxUpdPerSolver, int numSMs) {
    int bx = (maxUpdPerSolver + 255) / 256;
    int cap = std::max(1, numSMs * 8 / std::max(1, nSolvers));
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    return dim3(bx, nSolvers, 1);
}

void launchApplyUpdates(const VarUpdate *upd, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver,
                        int64_t avail, const DeviceTables &t, int numSMs, cudaStream_t s, int64_t *launches, VarUpdate *keep,
                        SolverRunParams *keepParams) {
    if (nSolvers == 0 || maxUpdPerSolver == 0) return;
    k_apply_updates<<<updateGrid(nSolvers, maxUpdPerSolver, numSMs), 256, 0, s>>>(upd, params, t, (long long)avail, keep,
                                                                                 keepParams);
    checkLaunch("k_apply_updates");
    ++*launches;
}

void launchCollapse(const VarUpdate *upd, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver,
                    int64_t avail, const DeviceTables &t, int numSMs, cudaStream_t s, int64_t *launches) {
    if (nSolvers == 0 || maxUpdPerSolver == 0) return;
    k_collapse<<<updateGrid(nSolvers, maxUpdPerSolver, numSMs), 256, 0, s>>>(upd, params, t, (long long)avail);
    checkLaunch("k_collapse");
    ++*launches;
}

void launchApplyDirect(const VarUpdate *const *src, const SolverRunParams *params, int nSolvers, int maxUpdPerSolver,
                       const DeviceTables &t, VarUpdate *keep, int numSMs, cudaStream_t s, int64_t *launches, const ApplyExtra &x) {
    const bool extra = (x.headWords | x.zeroAWords | x.zeroBWords) != 0;
    if (nSolvers == 0 || (maxUpdPerSolver == 0 && !extra)) return;
    k_apply_direct<<<updateGrid(nSolvers, std::max(1, maxUpdPerSolver), numSMs), 256, 0, s>>>(src, params, t, keep, x);
    checkLaunch("k_apply_direct");
    ++*launches;
}

void launchEmit(const EmitArgs &a, cudaStream_t s, int64_t *launches) {
    if (a.nSolvers <= 0) return;
    k_emit_scan<<<a.nSolvers, kRecBuckets, 0, s>>>(a);
    static const bool split = getenv("GSS_EMIT_SPLIT") != nullptr; // the two-kernel form, for comparison
    if (a.solverDone && !split) {
        k_emit_fused<<<(kRecBuckets / kSortWarps) * a.nSolvers, 256, 0, s>>>(a);
        checkLaunch("k_emit");
        *launches += 2;
        return;
    }
    k_emit_sort<<<dim3(kRecBuckets / kSortWarps, a.nSolvers, 1), kSortWarps * 32, 0, s>>>(a);
    // enough writers for PCIe: the blocks of a solver take chunks of its entries in turn
    const unsigned int perSolver = std::max(1u, std::min(64u, 1184u / (unsigned int)a.nSolvers));
    k_emit_write<<<dim3(perSolver, a.nSolvers, 1), 256, 0, s>>>(a);
    checkLaunch("k_emit");
    *launches += 3;
}

__global__ void k_bump_keys(const unsigned long long *__restrict__ keys, long long n, const LenDir *__restrict__ dir, int nDir,
                            float inc, int *overflow) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        const int len = (int)(key >> 32), idx = (int)(unsigned int)key;
        int lo = 0, hi = nDir - 1; // descending length
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (dir[mid].len <= len) hi = mid; else lo = mid + 1;
        }
        float old = atomicAdd(dir[lo].acts + idx, inc);
        if (old + inc > 1e19f) *overflow = 1;
    }
}

void launchBumpFromKeys(const unsigned long long *keys, long long n, const LenDir *dir, int nDir, float inc, int *overflow,
                        cudaStream_t s, int64_t *launches) {
    if (n <= 0) return;
    k_bump_keys<<<(unsigned int)std::min<long long>((n + 255) / 256, 1184), 256, 0, s>>>(keys, n, dir, nDir, inc, overflow);
    checkLaunch("k_bump_keys");
    ++*launches;
}

void launchBumpFromRecs(const unsigned long long *recKeys, unsigned int recCap, const EmitSolver *solverInfo, int nSolvers,
                        unsigned int maxCount, const LenDir *dir, int nDir, float inc, int *overflow, cudaStream_t s,
                        int64_t *launches) {
    if (nSolvers <= 0 || maxCount == 0) return;
    dim3 grid(std::max(1u, std::min((maxCount + 255u) / 256u, 64u)), nSolvers, 1);
    k_bump_recs<<<grid, 256, 0, s>>>(recKeys, recCap, solverInfo, dir, nDir, inc, overflow);
    checkLaunch("k_bump_recs");
    ++*launches;
}

int filterVariantCount() { return kNumFilterVariants; }
const char *filterVariantName(int v) { return v >= 0 && v < kNumFilterVariants ? kFilterVariants[v].name : ""; }
void setFilterVariant(int v) { gFilterVariant = v >= 0 && v < kNumFilterVariants ? v : kDefaultFilterVariant; }

void launchFilterOnly(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches) {
    if (a.totalTiles == 0) return;
    if (gFilterVariant < 0) {
        const char *e = getenv("GSS_FILTER_VARIANT");
        setFilterVariant(e ? atoi(e) : kDefaultFilterVariant);
    }
    const FilterVariant &fv = kFilterVariants[gFilterVariant];
    int threads = std::min(dims.threads, fv.threads);
    int warpsPerBlock = threads / 32;
    size_t smem = (size_t)a.nDir * sizeof(int) + (size_t)fv.ringBytesPerWarp * warpsPerBlock;
    if (fv.ringBytesPerWarp) { // the ring needs the large shared-memory carve-out (L1 hit rate of this kernel: 3 %)
        static bool configured[64][32] = {};
        int dev = 0;
        GSS_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && gFilterVariant < 32 && !configured[dev][gFilterVariant]) {
            GSS_CUDA(cudaFuncSetAttribute((const void *)fv.kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            GSS_CUDA(cudaFuncSetAttribute((const void *)fv.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            configured[dev][gFilterVariant] = true;
        }
    }
    int blocks = resolveBlocks((const void *)fv.kernel, threads, smem, numSMs, dims.blocks,
                               ((long long)a.totalTiles + warpsPerBlock - 1) / warpsPerBlock);
    fv.kernel<<<blocks, threads, smem, s>>>(a);
    checkLaunch("k_filter");
    ++*launches;
}

struct ExactVariant {
    void (*kernel)(CheckArgs);
    const char *name;
};
const ExactVariant kExactVariants[] = {
    {k_exact_t<4, 3>, "4 survivors per warp step, 3 x 256 threads/SM"},
    {k_exact_t<4, 4>, "4 survivors, 4 x 256"},
    {k_exact_t<2, 5>, "2 survivors, 5 x 256"},
    {k_exact_t<8, 2>, "8 survivors, 2 x 256"},
    {k_exact_t<2, 4>, "2 survivors, 4 x 256"},
    {k_exact_t<1, 6>, "1 survivor, 6 x 256"},
};
constexpr int kNumExactVariants = (int)(sizeof(kExactVariants) / sizeof(kExactVariants[0]));
int gExactVariant = -1;
constexpr int kDefaultExactVariant = 0;
int exactVariantCount() { return kNumExactVariants; }
const char *exactVariantName(int v) { return v >= 0 && v < kNumExactVariants ? kExactVariants[v].name : ""; }
void setExactVariant(int v) { gExactVariant = v >= 0 && v < kNumExactVariants ? v : kDefaultExactVariant; }

void launchExactOnly(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches) {
    if (a.totalTiles == 0) return;
    if (gExactVariant < 0) {
        const char *e = getenv("GSS_EXACT_VARIANT");
        setExactVariant(e ? atoi(e) : kDefaultExactVariant);
    }
    void (*k_exact)(CheckArgs) = kExactVariants[gExactVariant].kernel;
    // the survivor count is only known on the device: a fixed grid strides over it
    int blocks2 = resolveBlocks((const void *)k_exact, dims.threads, 0, numSMs, dims.blocks, -1);
    k_exact<<<blocks2, dims.threads, 0, s>>>(a);
    checkLaunch("k_exact");
    ++*launches;
}

void launchCheck(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches) {
    if (a.totalTiles == 0) return;
    launchFilterOnly(a, dims, numSMs, s, launches);
    launchExactOnly(a, dims, numSMs, s, launches);
}

void launchSliceTables(const DeviceTables &t, int groupBase, int groupSolvers, uint2 *sliced, int numSMs, cudaStream_t s,
                       int64_t *launches) {
    k_slice_tables<<<numSMs * 8, 256, 0, s>>>(t, groupBase, groupSolvers, sliced);
    checkLaunch("k_slice_tables");
    ++*launches;
}

void launchCheckDenseSliced(const CheckArgs &a, const uint2 *sliced, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches) {
    if (a.totalTiles == 0) return;
    const int threads = 128;
    const size_t smem = (size_t)a.nDir * sizeof(int);
    const int blocks = resolveBlocks((const void *)k_check_dense_sliced, threads, smem, numSMs, dims.blocks,
                                     ((long long)a.totalTiles * 2 + 3) / 4);
    for (int base = 0; base < a.groupSolvers; base += kSliceSolvers) { // one sweep of the clauses per slice of 8 solvers
        k_check_dense_sliced<<<blocks, threads, smem, s>>>(a, sliced + (size_t)(base / kSliceSolvers) * a.tables.varCap * kSliceSolvers, base);
        checkLaunch("k_check_dense_sliced");
        ++*launches;
    }
}

void launchCheckDense(const CheckArgs &a, LaunchDims dims, int numSMs, cudaStream_t s, int64_t *launches) {
    if (a.totalTiles == 0) return;
    int threads = 128;
    size_t smem = (size_t)a.nDir * sizeof(int);
    int blocks = resolveBlocks((const void *)k_check_dense, threads, smem, numSMs, dims.blocks,
                               ((long long)a.totalTiles * (kTileClauses / kDenseGroup) + 3) / 4);
    k_check_dense<<<blocks, threads, smem, s>>>(a);
    checkLaunch("k_check_dense");
    ++*launches;
}

} // namespace gss
