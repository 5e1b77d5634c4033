// Probe (not part of the library): the OUTPUT phase of one radix-scatter tile, three ways, to decide how the
// scatter passes should leave shared memory on sm_100a.
//   mode 0  SoA, one STG.32 to keys[] and one to vals[] per element      (what round 1's kernel does)
//   mode 1  AoS (key, value) pairs, one STG.64 per element
//   mode 2  AoS pairs, one cp.async.bulk shared -> global per digit run (16-byte aligned middle) + <= 2 peeled STG.64
// A tile is 8192 staged elements in 256 digit runs of ~32 elements; run (tile, d) lands at base[tile][d] in the
// output, the runs of one digit being contiguous across tiles (exactly the write pattern of an LSD pass).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store_probe bulk_store_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int T = 512, TILE = 8192, RADIX = 256;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

// run lengths: 32 +- x in pairs, so that every tile sums to 8192
__device__ __forceinline__ uint32_t run_len(uint32_t tile, uint32_t d) {
    const uint32_t h = hash32(tile * 128u + (d >> 1));
    const int x = (int)(h % 17u);
    return (uint32_t)(32 + (((d & 1) ^ ((h >> 8) & 1)) ? x : -x));
}

__global__ void base_kernel(uint32_t *base, uint32_t n_tiles, uint64_t per_digit) {
    const uint32_t d = threadIdx.x;
    uint64_t run = (uint64_t)d * per_digit;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        base[(uint64_t)t * RADIX + d] = (uint32_t)run;
        run += run_len(t, d);
    }
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(T, 2) probe(uint2 *out_pairs, uint32_t *out_keys, uint32_t *out_vals,
                                              const uint32_t *__restrict__ base, uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char raw[];
    uint2 *stage = reinterpret_cast<uint2 *>(raw);                                    // TILE + RADIX slots
    uint32_t *soff = reinterpret_cast<uint32_t *>(raw + (TILE + RADIX) * 8);          // smem offset of run d
    uint32_t *gbase = soff + RADIX;                                                    // global element index of run d
    uint32_t *delta = gbase + RADIX;                                                   // gbase - soff
    uint32_t *wsum = delta + RADIX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // run geometry: exclusive scan of the lengths (+ one pad slot where the parity of smem and global differ)
        uint32_t len = 0, gb = 0, pad = 0;
        if (tid < RADIX) {
            len = run_len(tile, tid);
            gb = base[(uint64_t)tile * RADIX + tid];
        }
        if (MODE == 2) {
            // parity fix needs the unpadded offset first: two scans (cheap: 256 values)
            uint32_t incl = len;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (tid < RADIX && lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t ex = incl - len;
            for (int w = 0; w < warp && tid < RADIX; ++w) ex += wsum[w];
            __syncthreads();
            // a run may start one slot later so that (smem slot & 1) == (global index & 1): at most RADIX pads in total.
            // pads accumulate, so decide sequentially per warp with a ballot-free trick: slot = ex + d (every run gets
            // its own spare slot), then parity-correct inside that spare
            const uint32_t slot = ex + (uint32_t)tid;  // room for one pad per earlier run... simplification: 1 spare each
            pad = ((slot ^ gb) & 1u);
            if (tid < RADIX) soff[tid] = slot + pad - (uint32_t)tid + (uint32_t)tid;  // = slot + pad
        } else {
            uint32_t incl = len;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (tid < RADIX && lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t ex = incl - len;
            for (int w = 0; w < warp && tid < RADIX; ++w) ex += wsum[w];
            if (tid < RADIX) soff[tid] = ex;
        }
        if (tid < RADIX) {
            gbase[tid] = gb;
            delta[tid] = gb - soff[tid];
        }
        __syncthreads();
        // staging (stands for the ranked scatter into shared memory): thread d writes its run
        if (tid < RADIX) {
            const uint32_t o = soff[tid];
            for (uint32_t i = 0; i < len; ++i) stage[o + i] = make_uint2(((uint32_t)tid << 24) | i, tile);
        }
        __syncthreads();
        if (MODE == 0 || MODE == 1) {
#pragma unroll
            for (int it = 0; it < TILE / T; ++it) {
                const uint32_t j = it * T + tid;
                const uint2 e = stage[j];
                const uint32_t dst = delta[e.x >> 24] + j;
                if (MODE == 0) {
                    out_keys[dst] = e.x;
                    out_vals[dst] = e.y;
                } else {
                    out_pairs[dst] = e;
                }
            }
        } else {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tid < RADIX && len) {
                uint32_t s = soff[tid], g = gb, l = len;
                if (g & 1u) {  // peel the head so that the middle is 16-byte aligned on both sides
                    out_pairs[g] = stage[s];
                    ++s, ++g, --l;
                }
                const uint32_t mid = l & ~1u;
                if (mid) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_pairs + g),
                                 "r"(smem_addr(stage + s)), "r"(mid * 8u)
                                 : "memory");
                }
                if (l & 1u) out_pairs[g + mid] = stage[s + mid];
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
}

int main(int argc, char **argv) {
    const uint64_t n_pairs = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 30);  // 8 GB of pairs
    const uint32_t n_tiles = (uint32_t)(n_pairs / TILE);
    const uint64_t per_digit = (uint64_t)n_tiles * 32 + (uint64_t)n_tiles * 17 + 64;  // room for the +-16 drift
    uint2 *pairs;
    uint32_t *keys, *vals, *base;
    const uint64_t cap = per_digit * RADIX;
    cudaMalloc(&pairs, cap * 8);
    keys = reinterpret_cast<uint32_t *>(pairs);
    vals = keys + cap;
    cudaMalloc(&base, (uint64_t)n_tiles * RADIX * 4);
    base_kernel<<<1, RADIX>>>(base, n_tiles, per_digit);
    cudaDeviceSynchronize();
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = (TILE + RADIX) * 8 + 3 * RADIX * 4 + 64;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            if (mode == 0) probe<0><<<sms * 2, T, smem>>>(pairs, keys, vals, base, n_tiles);
            if (mode == 1) probe<1><<<sms * 2, T, smem>>>(pairs, keys, vals, base, n_tiles);
            if (mode == 2) probe<2><<<sms * 2, T, smem>>>(pairs, keys, vals, base, n_tiles);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            if (rep && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("mode %d (%s): %.3f ms for %.2f GB written = %.0f GB/s  [%s]\n", mode,
               mode == 0 ? "SoA STG.32 x2" : mode == 1 ? "AoS STG.64" : "AoS bulk store per run", best,
               n_tiles * (double)TILE * 8 / 1e9, n_tiles * (double)TILE * 8 / 1e9 / (best * 1e-3), cudaGetErrorString(e));
    }
    return 0;
}
