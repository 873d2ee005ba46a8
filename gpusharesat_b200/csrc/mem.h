// mem.h -- the small arena/staging layer that replaces the reference's CorrespArr/ContigCopy
// substrate (gpuShareLib/CorrespArr.cuh, ContigCopy.cuh).  Not a port: two plain growable
// buffers, one pinned-host, one device, both stream-ordered.
#pragma once
#include "common.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace gss {

// address space for a host buffer that grows in place (vmem.cc: anonymous mapping, untouched pages cost nothing)
void *hostReserve(size_t bytes); // nullptr: not available
void hostUnreserve(void *p, size_t bytes);

// Growable host buffer, page-locked when the allocation is below the pinned budget
// (reference option maxPageLockedMemory, GpuClauseSharer.h:44-47).
// Two growth strategies.  Default: a larger block, copy, free (small, short-lived buffers).  setInPlace(): the
// host mirror of a clause arena grows for as long as the solvers learn, and re-allocating page-locked memory
// is expensive (allocating 100 MB of it takes tens of milliseconds, and the old contents have to be copied:
// measured as 0.3 - 1.5 s gpuRun() spikes while 10 M clauses streamed in).  Such a buffer reserves address
// space once (anonymous mapping, no memory behind it until touched) and page-locks it CHUNK BY CHUNK
// (cudaHostRegister) as it grows: the data never moves, only new memory is ever pinned.  Copies between such a
// buffer and the device are split at chunk boundaries (copyToDevice / copyFromDevice), so every piece lies
// inside one registered range.
template <typename T> class HostBuf {
    T *p_ = nullptr;
    size_t size_ = 0, cap_ = 0;
    bool pinned_ = false;
    size_t pinnedLimitBytes_ = (size_t)1 << 40;
    // in-place mode
    bool inPlace_ = false;
    size_t reservedBytes_ = 0, usableBytes_ = 0, registeredBytes_ = 0;
    // chunks: [0, 64 KB), then doubling up to 8 MB, then 8 MB each (a database has one mirror per clause
    // length, most of them tiny: the first chunk must be small; large mirrors must not need many registrations)
    static constexpr size_t kFirstChunk = (size_t)64 << 10, kChunkBytes = (size_t)8 << 20;
    static size_t chunkEnd(size_t off) { // end of the chunk that contains byte `off`
        if (off < kFirstChunk) return kFirstChunk;
        if (off >= kChunkBytes) return (off / kChunkBytes + 1) * kChunkBytes;
        size_t e = kFirstChunk;
        while (e <= off) e *= 2;
        return e;
    }

    static T *alloc(size_t n, size_t limit, bool &pinned) {
        if (n == 0) return nullptr;
        void *q = nullptr;
        pinned = false;
        if (n * sizeof(T) <= limit) {
            if (cudaHostAlloc(&q, n * sizeof(T), cudaHostAllocDefault) == cudaSuccess) pinned = true;
            else { cudaGetLastError(); q = nullptr; }
        }
        if (!q) {
            q = malloc(n * sizeof(T));
            if (!q) GSS_DIE("out of host memory");
        }
        return (T *)q;
    }
    static void release(T *q, bool pinned) {
        if (!q) return;
        if (pinned) cudaFreeHost(q); else free(q);
    }
    void releaseAll() {
        if (inPlace_) {
            uint8_t *base = reinterpret_cast<uint8_t *>(p_);
            for (size_t off = 0; off < registeredBytes_; off = chunkEnd(off)) cudaHostUnregister(base + off);
            if (p_) hostUnreserve(p_, reservedBytes_);
            inPlace_ = false;
            reservedBytes_ = usableBytes_ = registeredBytes_ = 0;
        } else {
            release(p_, pinned_);
        }
        p_ = nullptr;
        cap_ = 0;
    }
    // make [0, bytes) usable (and page-locked while the budget lasts); false: beyond the reserved range
    bool growInPlace(size_t bytes) {
        if (bytes > reservedBytes_) return false;
        uint8_t *base = reinterpret_cast<uint8_t *>(p_);
        while (usableBytes_ < bytes) {
            const size_t len = std::min(chunkEnd(usableBytes_), reservedBytes_) - usableBytes_;
            if (registeredBytes_ == usableBytes_ && usableBytes_ + len <= pinnedLimitBytes_) {
                if (cudaHostRegister(base + usableBytes_, len, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess)
                    registeredBytes_ += len;
                else
                    cudaGetLastError(); // (no device, or out of lockable memory: ordinary memory from here on)
            }
            usableBytes_ += len;
        }
        pinned_ = registeredBytes_ == usableBytes_;
        cap_ = usableBytes_ / sizeof(T);
        return true;
    }
    template <typename F> void forEachPiece(size_t firstElem, size_t nElems, F f) const {
        size_t off = firstElem * sizeof(T), end = off + nElems * sizeof(T);
        while (off < end) {
            const size_t stop = inPlace_ ? std::min(end, chunkEnd(off)) : end;
            f(off, stop - off);
            off = stop;
        }
    }

public:
    HostBuf() = default;
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
    ~HostBuf() { releaseAll(); }
    void setPinnedLimit(size_t bytes) { pinnedLimitBytes_ = bytes; }
    // before the first element: reserve maxBytes of address space and grow inside it (falls back to the default
    // strategy when the address space cannot be reserved, or later when the buffer outgrows it)
    void setInPlace(size_t maxBytes) {
        if (p_ || inPlace_) return;
        void *q = hostReserve(maxBytes);
        if (!q) return;
        p_ = static_cast<T *>(q);
        inPlace_ = true;
        reservedBytes_ = maxBytes;
    }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return size_; }
    size_t capacity() const { return cap_; }
    bool empty() const { return size_ == 0; }
    bool pinned() const { return pinned_; } // page-locked: a kernel can read / write it in place
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    void clear() { size_ = 0; }
    void reserve(size_t n) {
        if (n <= cap_) return;
        if (inPlace_) {
            if (growInPlace(n * sizeof(T))) return;
            // outgrown the reserved range: one move into an ordinary block, default strategy from here on
            size_t nc = cap_ ? cap_ : 64;
            while (nc < n) nc *= 2;
            bool pinned;
            T *q = alloc(nc, pinnedLimitBytes_, pinned);
            if (size_) memcpy(q, p_, size_ * sizeof(T));
            releaseAll();
            p_ = q; cap_ = nc; pinned_ = pinned;
            return;
        }
        size_t nc = cap_ ? cap_ : 64;
        while (nc < n) nc *= 2;
        bool pinned;
        T *q = alloc(nc, pinnedLimitBytes_, pinned);
        if (size_) memcpy(q, p_, size_ * sizeof(T));
        release(p_, pinned_);
        p_ = q; cap_ = nc; pinned_ = pinned;
    }
    // new elements are zero-filled
    void resize(size_t n) {
        reserve(n);
        if (n > size_) memset(p_ + size_, 0, (n - size_) * sizeof(T));
        size_ = n;
    }
    void push_back(const T &v) { reserve(size_ + 1); p_[size_++] = v; }
    T *append(size_t n) { reserve(size_ + n); T *r = p_ + size_; size_ += n; return r; }
    // asynchronous copies of elements [firstElem, firstElem + nElems) to / from device memory
    void copyToDevice(void *dev, size_t firstElem, size_t nElems, cudaStream_t stream) const {
        const uint8_t *h = reinterpret_cast<const uint8_t *>(p_);
        const size_t base = firstElem * sizeof(T);
        forEachPiece(firstElem, nElems, [&](size_t off, size_t len) {
            GSS_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(dev) + (off - base), h + off, len, cudaMemcpyHostToDevice, stream));
        });
    }
    void copyFromDevice(const void *dev, size_t firstElem, size_t nElems, cudaStream_t stream) {
        uint8_t *h = reinterpret_cast<uint8_t *>(p_);
        const size_t base = firstElem * sizeof(T);
        forEachPiece(firstElem, nElems, [&](size_t off, size_t len) {
            GSS_CUDA(cudaMemcpyAsync(h + off, static_cast<const uint8_t *>(dev) + (off - base), len, cudaMemcpyDeviceToHost, stream));
        });
    }
};

// Device memory that grows in place: a reserved range of virtual addresses with physical chunks
// mapped behind it as the buffer grows (vmem.cc).
namespace vm {
bool available(); // the driver's virtual memory management entry points could be resolved
struct Block {
    struct Chunk {
        unsigned long long handle;
        size_t bytes;
    };
    void *base = nullptr;
    size_t reserved = 0, mapped = 0, gran = 0;
    int device = 0;
    std::vector<Chunk> chunks;
    // make at least `bytes` usable; false: out of device memory (nothing changed).  The data never moves
    // unless the reserved range itself is outgrown (then: same chunks, new addresses, after a stream sync).
    bool grow(size_t bytes, cudaStream_t stream);
    void release();

private:
    bool reserveRange(size_t bytes);
    bool mapChunks(size_t from);
};
} // namespace vm

// Growable device buffer.  tryReserve returns false instead of dying when the device is out of memory
// (the caller then reduces the clause database like the reference, GpuRunner.cu:243-246).
// Two growth strategies: the default allocates a larger block and copies device-to-device on the given
// stream (small, short-lived buffers); setInPlace() buffers -- the clause arenas -- map more physical
// memory behind the same addresses (vm::Block: no copy, no synchronisation, no free).
template <typename T> class DevBuf {
    T *p_ = nullptr;
    size_t cap_ = 0;
    bool wantInPlace_ = false, inPlace_ = false;
    vm::Block block_;
    // a mapping granule is 2 MB: buffers below this size stay ordinary allocations (a database has one
    // arena per clause length, most of them tiny) and move behind a virtual range once, when they pass it
    static constexpr size_t kInPlaceMinBytes = (size_t)8 << 20;

public:
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { free(); }
    // ignored when the driver lacks the entry points
    void setInPlace() { wantInPlace_ = vm::available(); }
    bool inPlace() const { return inPlace_; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t capacity() const { return cap_; }
    void swap(DevBuf &o) {
        std::swap(p_, o.p_);
        std::swap(cap_, o.cap_);
        std::swap(wantInPlace_, o.wantInPlace_);
        std::swap(inPlace_, o.inPlace_);
        std::swap(block_, o.block_);
    }
    // keep = number of leading elements whose contents must survive the growth
    bool tryReserve(size_t n, size_t keep, cudaStream_t stream, bool exact = false) {
        if (n <= cap_) return true;
        if (inPlace_ || (wantInPlace_ && n * sizeof(T) >= kInPlaceMinBytes)) {
            if (!block_.grow(n * sizeof(T), stream)) return false;
            T *q = static_cast<T *>(block_.base);
            if (!inPlace_) { // the one move of this buffer's life
                if (keep && p_) GSS_CUDA(cudaMemcpyAsync(q, p_, keep * sizeof(T), cudaMemcpyDeviceToDevice, stream));
                if (p_) {
                    GSS_CUDA(cudaStreamSynchronize(stream));
                    cudaFree(p_);
                }
                inPlace_ = true;
            }
            p_ = q;
            cap_ = block_.mapped / sizeof(T);
            return true;
        }
        size_t nc = n;
        if (!exact) {
            nc = cap_ ? cap_ : 256;
            while (nc < n) nc *= 2;
        }
        T *q = nullptr;
        cudaError_t e = cudaMalloc(&q, nc * sizeof(T));
        if (e != cudaSuccess && nc != n) { cudaGetLastError(); nc = n; e = cudaMalloc(&q, nc * sizeof(T)); }
        if (e != cudaSuccess) { cudaGetLastError(); return false; }
        if (keep && p_) GSS_CUDA(cudaMemcpyAsync(q, p_, keep * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        if (p_) {
            // the old block may still be read by kernels already queued on the stream
            GSS_CUDA(cudaStreamSynchronize(stream));
            cudaFree(p_);
        }
        p_ = q; cap_ = nc;
        return true;
    }
    void reserve(size_t n, size_t keep, cudaStream_t stream) {
        if (!tryReserve(n, keep, stream)) GSS_DIE("out of device memory");
    }
    void free() {
        if (inPlace_) block_.release();
        else if (p_) cudaFree(p_);
        inPlace_ = false;
        p_ = nullptr;
        cap_ = 0;
    }
};

} // namespace gss
