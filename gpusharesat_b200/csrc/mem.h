// mem.h -- the small arena/staging layer that replaces the reference's CorrespArr/ContigCopy
// substrate (gpuShareLib/CorrespArr.cuh, ContigCopy.cuh).  Not a port: two plain growable
// buffers, one pinned-host, one device, both stream-ordered.
#pragma once
#include "common.h"
#include <cstring>

namespace gss {

// Growable host buffer, page-locked when the allocation is below the pinned budget
// (reference option maxPageLockedMemory, GpuClauseSharer.h:44-47).
template <typename T> class HostBuf {
    T *p_ = nullptr;
    size_t size_ = 0, cap_ = 0;
    bool pinned_ = false;
    size_t pinnedLimitBytes_ = (size_t)1 << 40;

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

public:
    HostBuf() = default;
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
    ~HostBuf() { release(p_, pinned_); }
    void setPinnedLimit(size_t bytes) { pinnedLimitBytes_ = bytes; }
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
};

// Growable device buffer.  Growth allocates a larger block and copies device-to-device on the
// given stream; tryReserve returns false instead of dying when the device is out of memory
// (the caller then reduces the clause database like the reference, GpuRunner.cu:243-246).
template <typename T> class DevBuf {
    T *p_ = nullptr;
    size_t cap_ = 0;

public:
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p_) cudaFree(p_); }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t capacity() const { return cap_; }
    // keep = number of leading elements whose contents must survive the growth
    bool tryReserve(size_t n, size_t keep, cudaStream_t stream, bool exact = false) {
        if (n <= cap_) return true;
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
    void free() { if (p_) cudaFree(p_); p_ = nullptr; cap_ = 0; }
};

} // namespace gss
