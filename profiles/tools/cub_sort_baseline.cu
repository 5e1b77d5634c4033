// Library baseline for the build's sort ("beat CUB on the same box"): cub::DeviceRadixSort::SortPairs
// on n (u32 key, u32 value) pairs with uniformly random keys, end_bit = 32 -- the same work as sorting the
// (hash, position) pairs of kmer_index<dna4,16>. NOT part of the product; profiles/ evidence only.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/tools/cub_sort_baseline.cu -o gpurun_out/cub_sort_baseline
//   gpurun_out/cub_sort_baseline 3000000000 32
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>

__global__ void fill(unsigned *k, unsigned *v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long x = (i + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
        x ^= x >> 31;
        x *= 0x94D049BB133111EBull;
        k[i] = (unsigned)(x >> 32);
        v[i] = (unsigned)i;
    }
}

int main(int argc, char **argv) {
    size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100000000ull;
    int end_bit = argc > 2 ? atoi(argv[2]) : 32;
    unsigned *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    void *tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, end_bit);
    cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        fill<<<148 * 8, 256>>>(k0, v0, n);
        cudaEventRecord(a);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, end_bit);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("cub::DeviceRadixSort::SortPairs n=%zu bits=%d: %.3f ms (%.2f Gpairs/s) tmp=%.1f MB err=%s\n", n, end_bit, best,
           n / best / 1e6, tmp_bytes / 1e6, cudaGetErrorString(e));
    return 0;
}
