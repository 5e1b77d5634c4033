// Random 8-byte gathers from a table much larger than L2, one kernel per load flavour: does any cache
// operator / prefetch-size qualifier change how many DRAM bytes one gather costs on sm_100a?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_variants gather_variants.cu
//   ./gather_variants                     (times)
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_variants   (bytes per gather)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint64_t load(const uint64_t *p) {
    uint64_t v;
    if (MODE == 0) asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 5) asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 6) asm volatile("ld.global.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 7) asm volatile("ld.global.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 8) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    if (MODE == 9) asm volatile("ld.global.L1::evict_first.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(const uint64_t *__restrict__ table, uint64_t n_words, uint64_t n_gathers,
                                             uint64_t *__restrict__ sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t g = t * 8; g < n_gathers; g += stride * 8) {
        uint64_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint64_t x = (g + j) * 0x9E3779B97F4A7C15ull;
            x ^= x >> 29;
            x *= 0xBF58476D1CE4E5B9ull;
            x ^= x >> 32;
            v[j] = load<MODE>(table + x % n_words);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
}

template <int MODE>
void run(const char *name, const uint64_t *table, uint64_t n_words, uint64_t n, uint64_t *sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<148 * 16, 256>>>(table, n_words, n >> 4, sink);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 16, 256>>>(table, n_words, n, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %6.2f Ggathers/s  (%s)\n", name, ms, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const uint64_t bytes = 16ull << 30, n_words = bytes / 8, n = 1ull << 28;
    uint64_t *table, *sink;
    cudaMalloc(&table, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(table, 1, bytes);
    run<0>("ld.global", table, n_words, n, sink);
    run<1>("ld.global.nc", table, n_words, n, sink);
    run<2>("ld.global.cg", table, n_words, n, sink);
    run<3>("ld.global.cs", table, n_words, n, sink);
    run<4>("ld.global.cv", table, n_words, n, sink);
    run<5>("ld.global.L1::no_allocate", table, n_words, n, sink);
    run<6>("ld.global.L2::64B", table, n_words, n, sink);
    run<7>("ld.global.L2::128B", table, n_words, n, sink);
    run<8>("ld.global.nc.L1::no_allocate.L2::64B", table, n_words, n, sink);
    run<9>("ld.global.L1::evict_first", table, n_words, n, sink);
    return 0;
}
