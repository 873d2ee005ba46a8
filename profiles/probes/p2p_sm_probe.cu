// Raw speed of SM-initiated peer traffic between two B200s of one box (no IPC, one process):
// how long does it take to move a 2.3 MB batch with plain stores (push) or loads (pull) from a kernel,
// against a copy-engine transfer?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_sm_probe p2p_sm_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n16, int fence) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
    if (fence) __threadfence_system();
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t cap = 64 << 20;
    void *a, *b;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&b, cap)); CK(cudaMemset(b, 1, cap));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a, cap)); CK(cudaMemset(a, 2, cap));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double sizes[] = {0.064, 0.5, 2.3, 16.0};
    for (double mb : sizes) {
        long long bytes = (long long)(mb * 1e6) / 16 * 16, n16 = bytes / 16;
        for (int mode = 0; mode < 5; mode++) {
            // 0 push 296 blocks, 1 push 1184 blocks, 2 pull 1184 blocks, 3 copy engine, 4 push 1184 + fence.sys
            float best = 1e9;
            for (int rep = 0; rep < 6; rep++) {
                CK(cudaMemsetAsync(a, rep, 1 << 20, s)); // something in between, like a real step
                CK(cudaEventRecord(e0, s));
                if (mode == 0) k_copy<<<296, 256, 0, s>>>((const uint4 *)a, (uint4 *)b, n16, 0);
                else if (mode == 1) k_copy<<<1184, 256, 0, s>>>((const uint4 *)a, (uint4 *)b, n16, 0);
                else if (mode == 2) k_copy<<<1184, 256, 0, s>>>((const uint4 *)b, (uint4 *)a, n16, 0);
                else if (mode == 3) CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, s));
                else k_copy<<<1184, 256, 0, s>>>((const uint4 *)a, (uint4 *)b, n16, 1);
                CK(cudaEventRecord(e1, s));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            const char *names[] = {"SM push, 296 blocks", "SM push, 1184 blocks", "SM pull, 1184 blocks", "copy engine", "SM push 1184 + fence.sys"};
            printf("%6.3f MB  %-26s %7.1f us  %7.1f GB/s\n", mb, names[mode], best * 1e3, bytes / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
