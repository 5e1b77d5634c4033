// Device-side synthetic input generator (bench.py). Counter-based SplitMix64, bit-identical
// to kmer_index_b200/synth.py: symbol i of stream `seed` = ((splitmix64(base(seed) + i) >> 32) * sigma) >> 32.
#include "launch.h"

namespace kb {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) synth_ranks_kernel(uint8_t *__restrict__ out, uint64_t n, uint64_t start,
                                                          uint32_t sigma, uint64_t base) {
    // 8 symbols per thread, one 8-byte store when aligned
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= n) return;
    uint64_t packed = 0;
    uint8_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint64_t x = splitmix64(base + start + i0 + j);
        r[j] = (uint8_t)(((x >> 32) * sigma) >> 32);
        packed |= (uint64_t)r[j] << (8 * j);
    }
    if (i0 + 8 <= n && ((uintptr_t)(out + i0) & 7) == 0) {
        *reinterpret_cast<uint64_t *>(out + i0) = packed;
    } else {
        for (int j = 0; j < 8 && i0 + j < n; ++j) out[i0 + j] = r[j];
    }
}

void launch_synth_ranks(uint8_t *d_out, uint64_t n, uint64_t start, uint32_t sigma, uint64_t seed, cudaStream_t stream) {
    if (n == 0) return;
    const uint64_t base = seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;
    const uint64_t threads = (n + 7) / 8;
    synth_ranks_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_out, n, start, sigma, base);
}

}  // namespace kb
