// reduce.cu -- reduceDb and the re-sort of streamed clauses ON THE DEVICE.
//
// Reference: HostClauses::reduceDb (Clauses.cu:426-465) picks an approximate median activity on the
// host (approxNthAct, :492-525), then PerSizeKeeper::removeClauses (:249-282) compacts the host copy
// clause by clause and the device copy is brought up to date through the update machinery.  Round 1 of
// this library compacted and re-sorted the host mirror and re-uploaded every arena (75 ms at 8 M
// clauses).  Here nothing but two small arrays crosses PCIe on the critical path:
//   1. k_act_hist     histogram of the activities over the reference's 20000 log-scale buckets (the
//                     bucket of a value is found by binary search over a table of bucket bounds built
//                     on the host with the reference's own formula, so host and device agree bit for
//                     bit) -> 160 KB to the host -> threshold (same arithmetic as approxNthAct)
//   2. k_reduce_keys  per clause: sort key = first literal, or "past the end" for a clause that goes
//                     (length >= 3, activity below the threshold or length == maximum), count the rest
//   3. a stable radix sort of (key, old index) per length (cub::DeviceRadixSort: library code, off the
//                     check path): the survivors come out in first-literal order, ties in their old
//                     order -- exactly what stable host compaction + stable counting sort produce
//   4. k_reduce_permute   literals (tile layout), ids and activities move to their new index in a second
//                     set of arenas; the sets are swapped, the old one is given back
//   5. the host mirror is refreshed by a device-to-host copy on a side stream (a reader of the mirror
//      waits for it: ClauseDb::waitMirror); only the partial last tile of every length is copied
//      synchronously, because new clauses are appended into it.
// The same pass with "nothing goes" puts the arenas back in first-literal order after clauses have been
// streamed in behind the sorted part (ClauseDb::resortOnDevice).
#include "clause_db.h"
#include <cub/device/device_radix_sort.cuh>

namespace gss {

namespace {

struct PermLen { // one clause length of the pass
    const int32_t *src;
    int32_t *dst;
    const int64_t *idsSrc;
    int64_t *idsDst;
    const float *actsSrc;
    float *actsDst;
    int32_t len, n;
    int32_t mode;   // 0: keep every clause, 1: keep activity >= threshold, 2: drop every clause
    int32_t inHist; // counted by approxNthAct (every length but the maximum)
    long long off;  // this length's slice of the key / value scratch arrays
};

__device__ __forceinline__ size_t wordPosDev(int len, int idx, int i) {
    return (size_t)(idx / kTileClauses) * kTileClauses * (size_t)len + (size_t)i * kTileClauses + (size_t)tileSlot(idx % kTileClauses);
}

// blockIdx.y = length.  bucket(x) = #{b >= 1 : bounds[b] <= bits(x)}
__global__ void __launch_bounds__(256) k_act_hist(const PermLen *__restrict__ L, const uint32_t *__restrict__ bounds, int nBuckets,
                                                  unsigned long long *__restrict__ hist) {
    const PermLen l = L[blockIdx.y];
    if (!l.inHist) return;
    for (int i0 = blockIdx.x * blockDim.x; i0 < l.n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool valid = i < l.n;
        int b = -1;
        if (valid) {
            const uint32_t bits = __float_as_uint(l.actsSrc[i]);
            int lo = 0, hi = nBuckets - 1; // largest b with bounds[b] <= bits (bounds[0] = 0)
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (bounds[mid] <= bits) lo = mid; else hi = mid - 1;
            }
            b = (bits & 0x80000000u) ? 0 : lo; // (negative: log is NaN on the host too; never happens)
        }
        // activities cluster in a few buckets: one atomic per distinct bucket of the warp
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, b);
        if (valid && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(hist + b, (unsigned long long)__popc(peers));
    }
}

__global__ void __launch_bounds__(256) k_reduce_keys(const PermLen *__restrict__ L, float threshold, uint32_t dropKey,
                                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int *__restrict__ kept) {
    const PermLen l = L[blockIdx.y];
    int mine = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < l.n; i += gridDim.x * blockDim.x) {
        const bool keep = l.mode == 0 || (l.mode == 1 && l.actsSrc[i] >= threshold);
        keys[l.off + i] = keep ? (uint32_t)l.src[wordPosDev(l.len, i, 0)] : dropKey;
        vals[l.off + i] = (uint32_t)i;
        mine += keep ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept + blockIdx.y, mine);
}

__global__ void __launch_bounds__(256) k_reduce_permute(const PermLen *__restrict__ L, const uint32_t *__restrict__ order,
                                                        const int *__restrict__ kept) {
    const PermLen l = L[blockIdx.y];
    const int n = kept[blockIdx.y];
    if (l.mode == 2) return;
    // (the unused slots of the last tile are zeroed: literal 0 is a valid table index for kernels that do not mask)
    const int nPad = (n + kTileClauses - 1) / kTileClauses * kTileClauses;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nPad; j += gridDim.x * blockDim.x) {
        int32_t *d = l.dst + wordPosDev(l.len, j, 0);
        if (j >= n) {
            for (int i = 0; i < l.len; i++) d[(size_t)i * kTileClauses] = 0;
            continue;
        }
        const int old = (int)order[l.off + j];
        const int32_t *s = l.src + wordPosDev(l.len, old, 0);
        for (int i = 0; i < l.len; i++) d[(size_t)i * kTileClauses] = s[(size_t)i * kTileClauses];
        l.idsDst[j] = l.idsSrc[old];
        l.actsDst[j] = l.actsSrc[old];
    }
}

} // namespace

struct ClauseDb::PermScratch {
    DevBuf<uint8_t> lens, cubTmp;
    DevBuf<uint32_t> keysA, keysB, valsA, valsB, bounds;
    DevBuf<unsigned long long> hist;
    DevBuf<int> kept;
};

void ClauseDb::PermScratchDeleter::operator()(PermScratch *p) const { delete p; }

void ClauseDb::initPermScratch() { permScratch_.reset(new PermScratch()); }

// give the spare arenas and the scratch arrays back (host-path reduce, out of device memory)
void ClauseDb::releaseSpare(cudaStream_t stream) {
    GSS_CUDA(cudaStreamSynchronize(stream));
    for (int s = 0; s <= maxLen_; s++) {
        perLen_[s]->devAlt.free();
        perLen_[s]->idsAlt.free();
        perLen_[s]->actsAlt.free();
    }
    initPermScratch();
}

bool ClauseDb::permuteOnDevice(cudaStream_t stream, bool dropByActivity) {
    HostProf hpAll(dropByActivity ? "reduce on device" : "re-sort on device");
    waitMirror();
    applyPendingDeviceRescales(stream);
    {
        int64_t dummy = 0; // clauses that have not reached the device yet go up first (with their activities)
        if (!uploadDirty(stream, &dummy)) return false;
    }
    std::vector<PermLen> lens;
    std::vector<int> lenOf;
    long long total = 0;
    int32_t maxLit = 1;
    for (int s = 1; s <= maxLen_; s++) {
        PerLen &pl = *perLen_[s];
        if (pl.n == 0) continue;
        GSS_CHECK(pl.actsOnDevice == pl.n);
        PermLen l;
        memset(&l, 0, sizeof(l));
        l.src = pl.dev.data();
        l.idsSrc = pl.idsDev.data();
        l.actsSrc = pl.actsDev.data();
        l.len = s;
        l.n = (int32_t)pl.n;
        // Clauses.cu:249-282 with minLimLbd = 0, maxLimLbd = MAX_CL_SIZE, lbd = clause length: lengths 1 and 2
        // are never touched; a clause of length s >= 3 stays iff s < max and activity >= threshold
        l.mode = !dropByActivity || s < 3 ? 0 : (s < maxLen_ ? 1 : 2);
        l.inHist = s < maxLen_;
        l.off = total;
        total += pl.n;
        lens.push_back(l);
        lenOf.push_back(s);
    }
    if (lens.empty()) {
        if (dropByActivity) {
            addedAtLastReduce_ = stats_.added;
            reduceDbs_++;
        }
        return true;
    }
    maxLit = 2 * std::max(1, maxVarPlusOne_) + 1;
    int litBits = 1;
    while ((1ll << litBits) <= (long long)maxLit) litBits++;
    const uint32_t dropKey = 1u << litBits;
    const int nL = (int)lens.size();

    // Memory first: when any of it is missing nothing has changed yet (the caller takes the host path / gives up).
    // The second set of arenas and the scratch arrays are kept between passes (they only ever grow, in place):
    // in steady state a pass allocates and frees nothing -- cudaMalloc / cudaFree synchronise the whole device
    // and cost tens of milliseconds for arrays of this size (measured: 58 + ~200 ms per pass at 10 M clauses).
    auto hpAlloc = std::make_unique<HostProf>("  permute: allocate");
    for (int k = 0; k < nL; k++) {
        PerLen &pl = *perLen_[lenOf[k]];
        if (lens[k].mode == 2) continue; // nothing of this length survives
        if (!pl.devAlt.tryReserve(wordsFor(lens[k].len, lens[k].n), 0, stream) ||
            !pl.idsAlt.tryReserve((size_t)lens[k].n, 0, stream) || !pl.actsAlt.tryReserve((size_t)lens[k].n, 0, stream))
            return false;
        lens[k].dst = pl.devAlt.data();
        lens[k].idsDst = pl.idsAlt.data();
        lens[k].actsDst = pl.actsAlt.data();
    }
    size_t cubBytes = 0;
    int32_t maxN = 0;
    for (const PermLen &l : lens) maxN = std::max(maxN, l.n);
    cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, maxN, 0, litBits + 1, stream);
    PermScratch &sc = *permScratch_;
    const bool needBounds = dropByActivity && sc.bounds.capacity() == 0;
    if (!sc.lens.tryReserve((size_t)nL * sizeof(PermLen), 0, stream) || !sc.keysA.tryReserve((size_t)total, 0, stream) ||
        !sc.keysB.tryReserve((size_t)total, 0, stream) || !sc.valsA.tryReserve((size_t)total, 0, stream) ||
        !sc.valsB.tryReserve((size_t)total, 0, stream) || !sc.kept.tryReserve((size_t)nL, 0, stream) ||
        !sc.cubTmp.tryReserve(cubBytes, 0, stream) ||
        (dropByActivity && (!sc.bounds.tryReserve(kActBuckets, 0, stream, true) || !sc.hist.tryReserve(kActBuckets, 0, stream, true))))
        return false;
    struct { // (the names the rest of the pass uses)
        PermLen *p;
    } lensDev{reinterpret_cast<PermLen *>(sc.lens.data())};
    struct { uint32_t *p; } keysA{sc.keysA.data()}, keysB{sc.keysB.data()}, valsA{sc.valsA.data()}, valsB{sc.valsB.data()}, boundsDev{sc.bounds.data()};
    struct { unsigned long long *p; } histDev{sc.hist.data()};
    struct { int *p; } keptDev{sc.kept.data()};
    struct { uint8_t *p; } cubTmp{sc.cubTmp.data()};
    hpAlloc.reset();
    auto hpKernels = std::make_unique<HostProf>("  permute: kernels + sync");
    GSS_CUDA(cudaMemcpyAsync(lensDev.p, lens.data(), (size_t)nL * sizeof(PermLen), cudaMemcpyHostToDevice, stream));
    GSS_CUDA(cudaMemsetAsync(keptDev.p, 0, (size_t)nL * sizeof(int), stream));
    const dim3 grid((unsigned int)std::min<long long>((maxN + 255) / 256, 2368), (unsigned int)nL);

    float threshold = 0.0f;
    if (dropByActivity) {
        addedAtLastReduce_ = stats_.added;
        reduceDbs_++;
        if (needBounds) { // (once per database: the table never changes)
            const std::vector<uint32_t> &bounds = actBucketBounds();
            GSS_CUDA(cudaMemcpyAsync(boundsDev.p, bounds.data(), (size_t)kActBuckets * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        }
        GSS_CUDA(cudaMemsetAsync(histDev.p, 0, (size_t)kActBuckets * sizeof(unsigned long long), stream));
        k_act_hist<<<grid, 256, 0, stream>>>(lensDev.p, boundsDev.p, kActBuckets, histDev.p);
        std::vector<int64_t> hist((size_t)kActBuckets);
        GSS_CUDA(cudaMemcpyAsync(hist.data(), histDev.p, (size_t)kActBuckets * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        GSS_CUDA(cudaStreamSynchronize(stream));
        threshold = thresholdFromHistogram(hist.data(), stats_.clauses / 2);
        logger_.log(2, "c Reducing gpu clause db, keeping clauses with act >= " + std::to_string(threshold) + "\n");
    }
    k_reduce_keys<<<grid, 256, 0, stream>>>(lensDev.p, threshold, dropKey, keysA.p, valsA.p, keptDev.p);
    for (const PermLen &l : lens) {
        if (l.mode == 2) continue;
        size_t bytes = cubBytes;
        cub::DeviceRadixSort::SortPairs(cubTmp.p, bytes, keysA.p + l.off, keysB.p + l.off, valsA.p + l.off, valsB.p + l.off, l.n, 0,
                                        litBits + 1, stream);
    }
    k_reduce_permute<<<grid, 256, 0, stream>>>(lensDev.p, valsB.p, keptDev.p);
    std::vector<int> kept((size_t)nL);
    GSS_CUDA(cudaMemcpyAsync(kept.data(), keptDev.p, (size_t)nL * sizeof(int), cudaMemcpyDeviceToHost, stream));
    GSS_CUDA(cudaStreamSynchronize(stream));
    GSS_CUDA(cudaGetLastError());

    hpKernels.reset();
    HostProf hpSwap("  permute: swap + mirror + free");
    // swap the arena sets, shrink the host mirror, start its refresh
    if (!mirrorStream_) {
        GSS_CUDA(cudaStreamCreateWithFlags(&mirrorStream_, cudaStreamNonBlocking));
        GSS_CUDA(cudaEventCreateWithFlags(&mirrorEv_, cudaEventDisableTiming));
        GSS_CUDA(cudaEventCreateWithFlags(&permuteDoneEv_, cudaEventDisableTiming));
    }
    for (int k = 0; k < nL; k++) {
        const int s = lenOf[k];
        PerLen &pl = *perLen_[s];
        const int64_t to = lens[k].mode == 2 ? 0 : kept[k];
        GSS_CHECK(to <= pl.n && (lens[k].mode != 0 || to == pl.n));
        stats_.clauses -= pl.n - to;
        stats_.lengthSum -= (pl.n - to) * s;
        pl.dev.swap(pl.devAlt); // (the old set is the spare set of the next pass)
        pl.idsDev.swap(pl.idsAlt);
        pl.actsDev.swap(pl.actsAlt);
        pl.n = to;
        pl.sortedN = to;
        pl.actsOnDevice = to;
        pl.dirtyFrom = to;
        pl.fullReupload = false;
        pl.lits.resize(wordsFor(s, to));
        pl.ids.resize((size_t)to);
        pl.acts.resize((size_t)to);
        if (to == 0) continue;
        // the last, partial tile now (new clauses are appended into it), the full tiles on the side stream
        const size_t tileWords = (size_t)kTileClauses * s;
        const size_t fullWords = (size_t)(to / kTileClauses) * tileWords;
        if (to % kTileClauses) pl.lits.copyFromDevice(pl.dev.data() + fullWords, fullWords, tileWords, stream);
    }
    GSS_CUDA(cudaEventRecord(permuteDoneEv_, stream));
    GSS_CUDA(cudaStreamWaitEvent(mirrorStream_, permuteDoneEv_, 0));
    for (int k = 0; k < nL; k++) {
        const int s = lenOf[k];
        PerLen &pl = *perLen_[s];
        if (pl.n == 0) continue;
        const size_t fullWords = (size_t)(pl.n / kTileClauses) * kTileClauses * (size_t)s;
        if (fullWords) pl.lits.copyFromDevice(pl.dev.data(), 0, fullWords, mirrorStream_);
        pl.ids.copyFromDevice(pl.idsDev.data(), 0, (size_t)pl.n, mirrorStream_);
        pl.acts.copyFromDevice(pl.actsDev.data(), 0, (size_t)pl.n, mirrorStream_);
    }
    GSS_CUDA(cudaEventRecord(mirrorEv_, mirrorStream_));
    mirrorPending_ = true;
    GSS_CUDA(cudaStreamSynchronize(stream)); // (the partial tiles)
    if (dropByActivity)
        logger_.log(2, "c Done reducing gpu clause db, clause count is " + std::to_string(stats_.clauses) + "\n");
    return true;
}

} // namespace gss
