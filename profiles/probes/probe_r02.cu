// probe_r02.cu -- round-2 micro-measurements behind the design decisions in DESIGN.md (not product code).
//   1. random row gathers out of a table that fits L2 (64 MB) / does not (256 MB): the practical ceiling
//      of the dense kernel's level-2 gather stream
//   2. SM reads of pinned host memory (zero-copy) vs cudaMemcpyAsync H2D at the batch size (2.3 MB)
//   3. SM writes to pinned host memory (coalesced / record-sized) vs cudaMemcpyAsync D2H at the result size
//   4. launch chains: 5 dependent small kernels, plain vs programmatic dependent launch vs CUDA graph
//   5. host round trips: 64 B D2H + sync, event sync
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_r02 probe_r02.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static double nowUs() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// rows of ROWB bytes; a warp fetches 256/ROWB rows per load instruction (8 B per lane), U loads in flight
template <int ROWB, int U> __global__ void __launch_bounds__(256) k_gather(const uint2 *__restrict__ table, uint32_t nRows,
                                                                           long long nLoadsPerWarp, uint32_t *out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int lanesPerRow = ROWB / 8;
    const int sub = lane / lanesPerRow, off = lane % lanesPerRow;
    uint32_t acc = 0;
    uint32_t seed = warp * 0x9E3779B9u + 17u;
    for (long long i = 0; i < nLoadsPerWarp; i += U) {
        uint2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t r = hash32(seed + (uint32_t)(i + u) * 64u + sub) % nRows;
            v[u] = __ldg(table + (size_t)r * lanesPerRow + off);
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc ^= v[u].x + v[u].y;
    }
    if (acc == 0x12345u) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_read16(const uint4 *__restrict__ src, long long n16, uint4 *dst, uint32_t *out) {
    uint32_t acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = src[i];
        if (dst) dst[i] = v;
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345u) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_write16(uint4 *__restrict__ dst, const uint4 *__restrict__ src, long long n16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// record-sized scattered writes: lane writes one 24 B record at a pseudo-random record slot
__global__ void __launch_bounds__(256) k_write_rec(uint2 *__restrict__ dst, long long nRec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nRec; i += (long long)gridDim.x * blockDim.x) {
        const long long r = (long long)(hash32((uint32_t)i) % (uint32_t)nRec);
        dst[r * 3 + 0] = make_uint2((uint32_t)i, 1);
        dst[r * 3 + 1] = make_uint2((uint32_t)i, 2);
        dst[r * 3 + 2] = make_uint2((uint32_t)i, 3);
    }
}

__global__ void k_small(uint32_t *p, int n) {
#if __CUDA_ARCH__ >= 900
    cudaGridDependencySynchronize();
#endif
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += 1;
#if __CUDA_ARCH__ >= 900
    cudaTriggerProgrammaticLaunchCompletion();
#endif
}

template <typename F> static float timeIt(cudaStream_t s, int iters, F f) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; i++) f();
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms * 1000.f / iters;
}

int main() {
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t *out; CK(cudaMalloc(&out, 64));
    printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"l2_bytes\":%d,\"persist_l2_max\":%d}\n", prop.name, sms, prop.l2CacheSize, prop.persistingL2CacheMaxSize);

    // ---- 1. gathers ----
    {
        const size_t big = (size_t)256 << 20;
        uint2 *table; CK(cudaMalloc(&table, big)); CK(cudaMemset(table, 1, big));
        const long long loadsPerWarp = 4096;
        auto run = [&](const char *name, auto kern, int rowB, size_t tableBytes, int blocksPerSM) {
            const uint32_t nRows = (uint32_t)(tableBytes / rowB);
            const int blocks = sms * blocksPerSM;
            float us = timeIt(s, 3, [&] { kern<<<blocks, 256, 0, s>>>(table, nRows, loadsPerWarp, out); });
            const double bytes = (double)blocks * 8 * loadsPerWarp * 256.0;
            printf("{\"probe\":\"gather\",\"variant\":\"%s\",\"row_bytes\":%d,\"table_mb\":%zu,\"blocks_per_sm\":%d,\"us\":%.1f,\"useful_gbs\":%.0f}\n",
                   name, rowB, tableBytes >> 20, blocksPerSM, us, bytes / us / 1e3);
        };
        for (size_t mb : {(size_t)16, (size_t)64, (size_t)96, (size_t)256}) {
            run("row64_u8", k_gather<64, 8>, 64, mb << 20, 8);
            run("row256_u8", k_gather<256, 8>, 256, mb << 20, 8);
            run("row32_u8", k_gather<32, 8>, 32, mb << 20, 8);
        }
        run("row64_u16", k_gather<64, 16>, 64, (size_t)64 << 20, 4);
        run("row64_u4", k_gather<64, 4>, 64, (size_t)64 << 20, 8);
        run("row128_u8", k_gather<128, 8>, 128, (size_t)64 << 20, 8);
        CK(cudaFree(table));
    }

    // ---- 2/3. host memory over PCIe ----
    {
        const size_t cap = (size_t)32 << 20;
        uint4 *host; CK(cudaHostAlloc((void **)&host, cap, cudaHostAllocMapped | cudaHostAllocPortable));
        for (size_t i = 0; i < cap / 16; i++) host[i] = make_uint4((uint32_t)i, 1, 2, 3);
        uint4 *dev; CK(cudaMalloc(&dev, cap)); CK(cudaMemset(dev, 3, cap));
        for (size_t bytes : {(size_t)64 << 10, (size_t)1 << 20, (size_t)2400000, (size_t)6 << 20, (size_t)24 << 20}) {
            const long long n16 = (long long)(bytes / 16);
            for (int blocksPerSM : {1, 2, 4}) {
                float us = timeIt(s, 5, [&] { k_read16<<<sms * blocksPerSM, 256, 0, s>>>(host, n16, dev, out); });
                printf("{\"probe\":\"pcie_sm_read\",\"bytes\":%zu,\"blocks_per_sm\":%d,\"us\":%.1f,\"gbs\":%.1f}\n", bytes, blocksPerSM, us, bytes / us / 1e3);
            }
            float us = timeIt(s, 5, [&] { cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s); });
            printf("{\"probe\":\"pcie_memcpy_h2d\",\"bytes\":%zu,\"us\":%.1f,\"gbs\":%.1f}\n", bytes, us, bytes / us / 1e3);
            for (int blocksPerSM : {1, 2, 4}) {
                float usw = timeIt(s, 5, [&] { k_write16<<<sms * blocksPerSM, 256, 0, s>>>(host, dev, n16); });
                printf("{\"probe\":\"pcie_sm_write\",\"bytes\":%zu,\"blocks_per_sm\":%d,\"us\":%.1f,\"gbs\":%.1f}\n", bytes, blocksPerSM, usw, bytes / usw / 1e3);
            }
            us = timeIt(s, 5, [&] { cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s); });
            printf("{\"probe\":\"pcie_memcpy_d2h\",\"bytes\":%zu,\"us\":%.1f,\"gbs\":%.1f}\n", bytes, us, bytes / us / 1e3);
        }
        {
            const long long nRec = 150000;
            float us = timeIt(s, 5, [&] { k_write_rec<<<sms * 2, 256, 0, s>>>((uint2 *)host, nRec); });
            printf("{\"probe\":\"pcie_sm_write_records24\",\"records\":%lld,\"us\":%.1f,\"gbs\":%.1f}\n", nRec, us, nRec * 24.0 / us / 1e3);
            us = timeIt(s, 5, [&] { k_write_rec<<<sms * 2, 256, 0, s>>>((uint2 *)dev, nRec); });
            printf("{\"probe\":\"hbm_sm_write_records24\",\"records\":%lld,\"us\":%.1f}\n", nRec, us);
        }
        // ---- 5. host round trips ----
        {
            double t0 = nowUs();
            const int n = 200;
            for (int i = 0; i < n; i++) {
                CK(cudaMemcpyAsync(host, dev, 64, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
            }
            printf("{\"probe\":\"d2h64_plus_sync_host_us\",\"us\":%.2f}\n", (nowUs() - t0) / n);
            cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            t0 = nowUs();
            for (int i = 0; i < n; i++) {
                k_small<<<1, 32, 0, s>>>(out, 1);
                CK(cudaEventRecord(ev, s));
                CK(cudaEventSynchronize(ev));
            }
            printf("{\"probe\":\"launch_plus_event_sync_host_us\",\"us\":%.2f}\n", (nowUs() - t0) / n);
            // kernel writes a flag in pinned host memory, the CPU spins on it
            volatile uint32_t *flag = (volatile uint32_t *)host;
            flag[0] = 0;
            t0 = nowUs();
            for (int i = 0; i < n; i++) {
                const uint32_t before = flag[0];
                k_small<<<1, 32, 0, s>>>((uint32_t *)host, 1);
                while (flag[0] == before) {}
            }
            printf("{\"probe\":\"launch_plus_host_flag_spin_us\",\"us\":%.2f}\n", (nowUs() - t0) / n);
            t0 = nowUs();
            for (int i = 0; i < 2000; i++) k_small<<<1, 32, 0, s>>>(out, 1);
            double tl = (nowUs() - t0) / 2000;
            CK(cudaStreamSynchronize(s));
            printf("{\"probe\":\"launch_host_cost_us\",\"us\":%.2f}\n", tl);
            t0 = nowUs();
            for (int i = 0; i < 2000; i++) CK(cudaMemcpyAsync(dev, host, 4096, cudaMemcpyHostToDevice, s));
            tl = (nowUs() - t0) / 2000;
            CK(cudaStreamSynchronize(s));
            printf("{\"probe\":\"memcpy_async_host_cost_us\",\"us\":%.2f}\n", tl);
        }
        CK(cudaFree(dev)); CK(cudaFreeHost(host));
    }

    // ---- 4. launch chains ----
    {
        uint32_t *p; CK(cudaMalloc(&p, 148 * 256 * 4)); CK(cudaMemset(p, 0, 148 * 256 * 4));
        const int n = 148 * 256;
        float plain = timeIt(s, 50, [&] { for (int k = 0; k < 5; k++) k_small<<<148, 256, 0, s>>>(p, n); });
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.stream = s; cfg.attrs = attr; cfg.numAttrs = 1;
        float pdl = timeIt(s, 50, [&] { for (int k = 0; k < 5; k++) CK(cudaLaunchKernelEx(&cfg, k_small, p, n)); });
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < 5; k++) k_small<<<148, 256, 0, s>>>(p, n);
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        float graph = timeIt(s, 50, [&] { CK(cudaGraphLaunch(ge, s)); });
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < 5; k++) CK(cudaLaunchKernelEx(&cfg, k_small, p, n));
        cudaGraph_t g2; cudaGraphExec_t ge2;
        CK(cudaStreamEndCapture(s, &g2));
        CK(cudaGraphInstantiate(&ge2, g2, 0));
        float graphPdl = timeIt(s, 50, [&] { CK(cudaGraphLaunch(ge2, s)); });
        printf("{\"probe\":\"chain5\",\"plain_us\":%.2f,\"pdl_us\":%.2f,\"graph_us\":%.2f,\"graph_pdl_us\":%.2f}\n", plain, pdl, graph, graphPdl);
        CK(cudaFree(p));
    }
    return 0;
}
