// Random 4-byte reads from ANOTHER GPU's memory over NVLink (one process, peer access), next to the same reads from
// local HBM: what a search kernel pays for a candidate list that lives on a peer. Variants: plain ld.global, the
// non-coherent ld.global.nc, and ld.global.nc.L2::64B (the qualifier the library's gathers use on local memory).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_gather_probe peer_gather_probe.cu ; ./peer_gather_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint32_t load(const uint32_t *p) {
    uint32_t v;
    if (MODE == 0) asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.relaxed.sys.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <int MODE>
__global__ void probe(const uint32_t *table, uint64_t n_words, uint64_t n_reads, uint32_t *out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (; i < n_reads; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull;
        x ^= x >> 29;
        x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 32;
        acc += load<MODE>(table + x % n_words);
    }
    if (acc == 0xFFFFFFFFu) *out = acc;
}

template <int MODE>
float run(const uint32_t *table, uint64_t n_words, uint64_t n_reads, uint32_t *out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    probe<MODE><<<148 * 16, 256>>>(table, n_words, n_reads / 8, out);
    cudaEventRecord(a);
    probe<MODE><<<148 * 16, 256>>>(table, n_words, n_reads, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int n_dev = 0;
    cudaGetDeviceCount(&n_dev);
    const uint64_t n_words = 1ull << 30, n_reads = 1ull << 26;  // 4 GiB table
    uint32_t *local = nullptr, *remote = nullptr, *out = nullptr;
    cudaSetDevice(0);
    cudaMalloc(&local, n_words * 4);
    cudaMalloc(&out, 4);
    cudaMemset(local, 1, n_words * 4);
    if (n_dev > 1) {
        cudaSetDevice(1);
        cudaMalloc(&remote, n_words * 4);
        cudaMemset(remote, 1, n_words * 4);
        cudaDeviceSynchronize();
        cudaSetDevice(0);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, 0, 1);
        printf("device 0 can access device 1: %d (enable: %s)\n", can, cudaGetErrorString(cudaDeviceEnablePeerAccess(1, 0)));
    }
    const char *names[4] = {"ld.global", "ld.global.nc", "ld.global.nc.L2::64B", "ld.global.relaxed.sys"};
    for (int where = 0; where < (remote ? 2 : 1); ++where) {
        const uint32_t *t = where ? remote : local;
        float ms[4] = {run<0>(t, n_words, n_reads, out), run<1>(t, n_words, n_reads, out), run<2>(t, n_words, n_reads, out),
                       run<3>(t, n_words, n_reads, out)};
        for (int m = 0; m < 4; ++m)
            printf("%-6s %-24s %8.3f ms  %7.2f G reads/s\n", where ? "peer" : "local", names[m], ms[m], n_reads / ms[m] / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
