// vmem.cc -- device buffers that grow IN PLACE (CUDA virtual memory management).
//
// A clause arena grows for as long as the solvers learn clauses.  Growing a cudaMalloc'ed block means
// allocating a larger one, copying, synchronising and freeing (the reference does exactly that,
// CorrespArr.cu:197-210; round 1 of this library did too and paid 18-81 ms per growth step at 8 M
// clauses).  Here a buffer reserves a range of virtual addresses once and maps physical chunks behind
// it as it grows: the data never moves, nothing is copied, no kernel has to drain, the pointer the
// kernels hold stays valid.  Chunks grow geometrically (an eighth of what is mapped, at least one
// allocation granule), so a buffer of any size consists of a few dozen mappings.  A buffer that
// outgrows its reserved range reserves a larger one and maps the SAME physical chunks there (still no
// copy; this one step has to wait for the kernels that use the old addresses).
//
// The driver entry points are resolved through the runtime (cudaGetDriverEntryPoint): the library
// does not link libcuda, so it still loads on a machine without a driver (the CPU test tier).
#include "mem.h"
#include <cuda.h>
#include <mutex>
#include <sys/mman.h>

namespace gss {

void *hostReserve(size_t bytes) {
    if (getenv("GPUSHARE_NO_VMM")) return nullptr;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    return p == MAP_FAILED ? nullptr : p;
}
void hostUnreserve(void *p, size_t bytes) { munmap(p, bytes); }

namespace vm {

namespace {

struct Api {
    CUresult (*addressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

template <typename F> bool resolve(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const Api &api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        a.ok = resolve("cuMemAddressReserve", a.addressReserve) && resolve("cuMemAddressFree", a.addressFree) &&
               resolve("cuMemCreate", a.create) && resolve("cuMemRelease", a.release) && resolve("cuMemMap", a.map) &&
               resolve("cuMemUnmap", a.unmap) && resolve("cuMemSetAccess", a.setAccess) &&
               resolve("cuMemGetAllocationGranularity", a.granularity);
    });
    return a;
}

CUmemAllocationProp propFor(int device) {
    CUmemAllocationProp p;
    memset(&p, 0, sizeof(p));
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    return p;
}

size_t roundUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

} // namespace

// (GPUSHARE_NO_VMM: plain cudaMalloc growth for buffers created while it is set, for comparison)
bool available() { return api().ok && !getenv("GPUSHARE_NO_VMM"); }

bool Block::reserveRange(size_t bytes) {
    const Api &a = api();
    if (!a.ok) return false;
    if (gran == 0) {
        GSS_CUDA(cudaGetDevice(&device));
        CUmemAllocationProp p = propFor(device);
        if (a.granularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) {
            gran = 0;
            return false;
        }
    }
    bytes = roundUp(bytes, gran);
    CUdeviceptr va = 0;
    if (a.addressReserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) return false;
    base = reinterpret_cast<void *>(va);
    reserved = bytes;
    return true;
}

// map the chunks [from, chunks.size()) at their offsets of the current range
bool Block::mapChunks(size_t from) {
    const Api &a = api();
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t off = 0;
    for (size_t i = 0; i < from; i++) off += chunks[i].bytes;
    for (size_t i = from; i < chunks.size(); i++) {
        const CUdeviceptr at = reinterpret_cast<CUdeviceptr>(base) + off;
        if (a.map(at, chunks[i].bytes, 0, (CUmemGenericAllocationHandle)chunks[i].handle, 0) != CUDA_SUCCESS) return false;
        if (a.setAccess(at, chunks[i].bytes, &acc, 1) != CUDA_SUCCESS) return false;
        off += chunks[i].bytes;
    }
    return true;
}

bool Block::grow(size_t bytes, cudaStream_t stream) {
    const Api &a = api();
    if (!a.ok) return false;
    if (bytes <= mapped) return true;
    if (!base && !reserveRange(std::max<size_t>(bytes * 4, (size_t)1 << 30))) return false;
    if (bytes > reserved) {
        // outgrown the reserved range: a larger range, the same physical chunks behind it (no copy).
        // Kernels already queued use the old addresses: they finish first.
        GSS_CUDA(cudaStreamSynchronize(stream));
        void *oldBase = base;
        const size_t oldReserved = reserved;
        if (mapped) a.unmap(reinterpret_cast<CUdeviceptr>(oldBase), mapped);
        if (!reserveRange(std::max(bytes * 4, oldReserved * 16))) {
            // put the old mapping back: the caller treats this like running out of memory
            base = oldBase;
            reserved = oldReserved;
            if (!mapChunks(0)) GSS_DIE("cannot restore a device mapping");
            return false;
        }
        a.addressFree(reinterpret_cast<CUdeviceptr>(oldBase), oldReserved);
        if (!mapChunks(0)) GSS_DIE("cannot re-map a device buffer into its larger address range");
    }
    // one more chunk: an eighth of what is mapped, at least what is asked for
    const size_t want = roundUp(std::max(bytes - mapped, mapped / 8), gran);
    const size_t least = roundUp(bytes - mapped, gran);
    CUmemAllocationProp p = propFor(device);
    CUmemGenericAllocationHandle h = 0;
    size_t got = std::min(want, reserved - mapped);
    if (a.create(&h, got, &p, 0) != CUDA_SUCCESS) {
        got = least;
        if (a.create(&h, got, &p, 0) != CUDA_SUCCESS) return false; // out of device memory
    }
    chunks.push_back(Chunk{(unsigned long long)h, got});
    if (!mapChunks(chunks.size() - 1)) {
        a.release(h);
        chunks.pop_back();
        return false;
    }
    mapped += got;
    return true;
}

void Block::release() {
    const Api &a = api();
    if (base) {
        if (mapped) a.unmap(reinterpret_cast<CUdeviceptr>(base), mapped);
        for (const Chunk &c : chunks) a.release((CUmemGenericAllocationHandle)c.handle);
        a.addressFree(reinterpret_cast<CUdeviceptr>(base), reserved);
    }
    chunks.clear();
    base = nullptr;
    reserved = mapped = 0;
}

} // namespace vm
} // namespace gss
