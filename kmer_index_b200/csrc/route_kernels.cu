// Query routing for the key-range multi-GPU search (sm_100a). GPU r holds index part r: the k-mers whose hash lies in
// the r-th slice of [0, sigma^k). A query is answered by the GPU that owns the hash of its FIRST k symbols (the seed of
// every plan of a single-k index); whether the query's other parts occur anywhere in the text -- the reference's
// whole-text rules, kmer_index.hpp:216-227 and :119-122 -- comes from a presence bitmap over the whole key space that
// every GPU holds. So a query needs exactly one GPU, and the GPUs exchange queries and results, never candidate lists:
//
//   origin GPU   route_pack      pack each query of its slice of the batch (b-bit words, fixed stride), hash its first k
//                                symbols, append the record to the send block of the owner          -> all-to-all (NCCL)
//   owner GPU    compact_blocks  the records received from all origins -> one packed batch -> the ordinary search
//                pack_return     per record: where its hits start in the owner's position array, its status
//                                                                                   -> all-to-all (blocks + positions)
//   origin GPU   unroute         counts and status back into batch order -> offsets scan -> place: copy every hit list
#include <algorithm>

#include "launch.h"
#include "query_pack.cuh"

namespace kb {

constexpr int kRouteThreads = 256;

// ---- block layouts (all offsets multiples of 8 bytes) ----------------------------------------------------------
// send block   : [0] u32 count | [64] u64 words[capacity * stride] | u32 origin[capacity] | u16 lens[capacity]
// return block : [0] u32 count, [8] u64 n_positions | [64] u32 starts[capacity + 1] | u8 status[capacity]
uint64_t route_block_bytes(uint32_t capacity, uint32_t stride) {
    const uint64_t b = 64 + (uint64_t)capacity * stride * 8 + (uint64_t)capacity * 4 + (uint64_t)capacity * 2;
    return (b + 63) / 64 * 64;
}
uint64_t route_return_block_bytes(uint32_t capacity) {
    const uint64_t b = 64 + ((uint64_t)capacity + 1) * 4 + (uint64_t)capacity;
    return (b + 63) / 64 * 64;
}

struct SendBlock {
    uint32_t *count;
    uint64_t *words;
    uint32_t *origin;
    uint16_t *lens;
};
__host__ __device__ inline SendBlock send_block(uint8_t *base, uint64_t block_bytes, uint32_t b, uint32_t capacity, uint32_t stride) {
    uint8_t *p = base + (uint64_t)b * block_bytes;
    SendBlock s;
    s.count = reinterpret_cast<uint32_t *>(p);
    s.words = reinterpret_cast<uint64_t *>(p + 64);
    s.origin = reinterpret_cast<uint32_t *>(p + 64 + (uint64_t)capacity * stride * 8);
    s.lens = reinterpret_cast<uint16_t *>(p + 64 + (uint64_t)capacity * stride * 8 + (uint64_t)capacity * 4);
    return s;
}
struct ReturnBlock {
    uint32_t *count;
    uint64_t *n_positions;
    uint32_t *starts;
    uint8_t *status;
};
__host__ __device__ inline ReturnBlock return_block(uint8_t *base, uint64_t block_bytes, uint32_t b, uint32_t capacity) {
    uint8_t *p = base + (uint64_t)b * block_bytes;
    ReturnBlock r;
    r.count = reinterpret_cast<uint32_t *>(p);
    r.n_positions = reinterpret_cast<uint64_t *>(p + 8);
    r.starts = reinterpret_cast<uint32_t *>(p + 64);
    r.status = p + 64 + ((uint64_t)capacity + 1) * 4;
    return r;
}

// ---- origin: pack + route ----------------------------------------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(kRouteThreads) route_pack_kernel(RouteArgs a) {
    extern __shared__ uint64_t smem_q[];  // (stride + 2) words per thread
    __shared__ uint32_t s_cnt[64], s_base[64];
    const int tid = threadIdx.x;
    if (tid < 64) s_cnt[tid] = 0;
    __syncthreads();
    const uint64_t q = (uint64_t)blockIdx.x * kRouteThreads + tid;
    uint64_t *qw = smem_q + (size_t)tid * (a.stride + 2);
    uint32_t owner = 0xFFFFFFFFu, rank_in_cta = 0, m = 0;
    if (q < a.n_queries) {
        const uint64_t off0 = a.q_offsets[q];
        const uint64_t m64 = a.q_offsets[q + 1] - off0;
        uint32_t status = 0xFFu;  // "travels": the owner decides
        if (m64 == 0) {
            status = KMER_B200_QUERY_UNDEFINED;  // assert(query.size() > 0), kmer_index.hpp:195
        } else if (a.mode == KMER_B200_MODE_REFERENCE_EXACT && m64 > kQuerySizeRange) {
            status = KMER_B200_QUERY_THROW_INVALID_ARGUMENT;  // kmer_index.hpp:507-509
        } else if (a.mode == KMER_B200_MODE_REFERENCE_EXACT && m64 == kQuerySizeRange) {
            status = KMER_B200_QUERY_UNDEFINED;  // out-of-bounds table read, kmer_index.hpp:512
        } else if (m64 < a.k || m64 > (uint64_t)a.stride * (64 / BITS)) {
            atomicOr(a.flags, 2u);  // shorter than k (prefix slabs span parts) or longer than the stride: not routable
            status = KMER_B200_QUERY_UNDEFINED;
        }
        a.status[q] = (uint8_t)status;
        if (status == 0xFFu) {
            m = (uint32_t)m64;
            const bool bad = pack_query_lane<BITS>(a.q_ranks, off0, m, a.q_offsets[a.n_queries], a.sigma, qw);
            if (bad) atomicOr(a.flags, 1u);
            const uint64_t key = key_at(qw, 0ull, a.k, (uint32_t)BITS, a.sigma);
            owner = (uint32_t)min((uint64_t)(a.n_parts - 1), key / a.part_width);
            rank_in_cta = atomicAdd(&s_cnt[owner], 1u);
        }
    }
    __syncthreads();
    if (tid < (int)a.n_parts && s_cnt[tid])
        s_base[tid] = atomicAdd(send_block(a.blocks, a.block_bytes, tid, a.capacity, a.stride).count, s_cnt[tid]);
    __syncthreads();
    if (owner == 0xFFFFFFFFu) return;
    const uint32_t slot = s_base[owner] + rank_in_cta;
    if (slot >= a.capacity) {
        atomicOr(a.flags, 4u);  // a send block is full: the caller retries with a larger capacity
        return;
    }
    const SendBlock sb = send_block(a.blocks, a.block_bytes, owner, a.capacity, a.stride);
    const uint32_t n_words = (m * BITS + 63) / 64;
    for (uint32_t w = 0; w < a.stride; ++w) sb.words[(uint64_t)slot * a.stride + w] = w < n_words ? qw[w] : 0ull;
    sb.origin[slot] = (uint32_t)q;
    sb.lens[slot] = (uint16_t)m;
}

void launch_route_pack(const RouteArgs &a, cudaStream_t stream) {
    if (a.n_queries == 0) return;
    const unsigned blocks = (unsigned)((a.n_queries + kRouteThreads - 1) / kRouteThreads);
    const size_t smem = (size_t)kRouteThreads * (a.stride + 2) * sizeof(uint64_t);
    if (a.bits == 2) {
        cudaFuncSetAttribute(route_pack_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        route_pack_kernel<2><<<blocks, kRouteThreads, smem, stream>>>(a);
    } else if (a.bits == 4) {
        cudaFuncSetAttribute(route_pack_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        route_pack_kernel<4><<<blocks, kRouteThreads, smem, stream>>>(a);
    } else {
        cudaFuncSetAttribute(route_pack_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        route_pack_kernel<8><<<blocks, kRouteThreads, smem, stream>>>(a);
    }
}

// ---- owner: received blocks -> one packed batch; results -> return blocks ------------------------------------------
__global__ void __launch_bounds__(kRouteThreads) compact_blocks_kernel(uint8_t *blocks, uint64_t block_bytes, uint32_t capacity,
                                                                        uint32_t stride, RoutePrefix pfx, uint64_t *__restrict__ words,
                                                                        uint16_t *__restrict__ lens) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * kRouteThreads + threadIdx.x;
    const uint32_t cnt = pfx.p[b + 1] - pfx.p[b];
    if (i >= cnt) return;
    const SendBlock sb = send_block(blocks, block_bytes, b, capacity, stride);
    const uint64_t j = (uint64_t)pfx.p[b] + i;
    for (uint32_t w = 0; w < stride; ++w) words[j * stride + w] = sb.words[(uint64_t)i * stride + w];
    lens[j] = sb.lens[i];
}

void launch_compact_blocks(uint8_t *d_blocks, uint64_t block_bytes, uint32_t n_parts, uint32_t capacity, uint32_t stride,
                           const RoutePrefix &pfx, uint64_t *d_words, uint16_t *d_lens, cudaStream_t stream) {
    uint32_t mx = 0;
    for (uint32_t b = 0; b < n_parts; ++b) mx = std::max(mx, pfx.p[b + 1] - pfx.p[b]);
    if (mx == 0) return;
    dim3 grid((mx + kRouteThreads - 1) / kRouteThreads, n_parts);
    compact_blocks_kernel<<<grid, kRouteThreads, 0, stream>>>(d_blocks, block_bytes, capacity, stride, pfx, d_words, d_lens);
}

__global__ void __launch_bounds__(kRouteThreads) pack_return_kernel(const uint64_t *__restrict__ offsets, const uint8_t *__restrict__ status,
                                                                     RoutePrefix pfx, uint8_t *ret_blocks, uint64_t ret_bytes,
                                                                     uint32_t capacity) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * kRouteThreads + threadIdx.x;
    const uint32_t cnt = pfx.p[b + 1] - pfx.p[b];
    if (i > cnt) return;
    const ReturnBlock rb = return_block(ret_blocks, ret_bytes, b, capacity);
    const uint64_t base = offsets[pfx.p[b]];
    rb.starts[i] = (uint32_t)(offsets[(uint64_t)pfx.p[b] + i] - base);
    if (i < cnt) rb.status[i] = status[(uint64_t)pfx.p[b] + i];
    if (i == 0) {
        *rb.count = cnt;
        *rb.n_positions = offsets[pfx.p[b + 1]] - base;
    }
}

void launch_pack_return(const uint64_t *d_offsets, const uint8_t *d_status, const RoutePrefix &pfx, uint32_t n_parts, uint8_t *d_ret_blocks,
                        uint64_t ret_bytes, uint32_t capacity, cudaStream_t stream) {
    uint32_t mx = 0;
    for (uint32_t b = 0; b < n_parts; ++b) mx = std::max(mx, pfx.p[b + 1] - pfx.p[b]);
    dim3 grid((mx + 1 + kRouteThreads - 1) / kRouteThreads, n_parts);
    pack_return_kernel<<<grid, kRouteThreads, 0, stream>>>(d_offsets, d_status, pfx, d_ret_blocks, ret_bytes, capacity);
}

// ---- origin: results back into batch order ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRouteThreads) unroute_counts_kernel(uint8_t *send_blocks, uint64_t block_bytes, uint32_t stride,
                                                                        uint8_t *ret_blocks, uint64_t ret_bytes, uint32_t capacity,
                                                                        RoutePrefix sent, uint64_t *__restrict__ counts,
                                                                        uint8_t *__restrict__ status) {
    const uint32_t p = blockIdx.y;
    const uint32_t i = blockIdx.x * kRouteThreads + threadIdx.x;
    if (i >= sent.p[p + 1] - sent.p[p]) return;
    const SendBlock sb = send_block(send_blocks, block_bytes, p, capacity, stride);
    const ReturnBlock rb = return_block(ret_blocks, ret_bytes, p, capacity);
    const uint32_t q = sb.origin[i];
    counts[q] = rb.starts[i + 1] - rb.starts[i];
    status[q] = rb.status[i];
}

__global__ void __launch_bounds__(kRouteThreads) unroute_place_kernel(uint8_t *send_blocks, uint64_t block_bytes, uint32_t stride,
                                                                       uint8_t *ret_blocks, uint64_t ret_bytes, uint32_t capacity,
                                                                       RoutePrefix sent, RoutePrefix64 seg, const uint32_t *__restrict__ recv_pos,
                                                                       const uint64_t *__restrict__ offsets, uint32_t *__restrict__ positions) {
    const uint32_t p = blockIdx.y;
    const uint32_t i = blockIdx.x * kRouteThreads + threadIdx.x;
    if (i >= sent.p[p + 1] - sent.p[p]) return;
    const ReturnBlock rb = return_block(ret_blocks, ret_bytes, p, capacity);
    const uint32_t s0 = rb.starts[i], s1 = rb.starts[i + 1];
    if (s1 == s0) return;
    const SendBlock sb = send_block(send_blocks, block_bytes, p, capacity, stride);
    const uint32_t *src = recv_pos + seg.p[p] + s0;
    uint32_t *dst = positions + offsets[sb.origin[i]];
    for (uint32_t j = 0; j < s1 - s0; ++j) dst[j] = src[j];
}

void launch_unroute_counts(uint8_t *d_send_blocks, uint64_t block_bytes, uint32_t stride, uint8_t *d_ret_blocks, uint64_t ret_bytes,
                           uint32_t capacity, uint32_t n_parts, const RoutePrefix &sent, uint64_t *d_counts, uint8_t *d_status,
                           cudaStream_t stream) {
    uint32_t mx = 0;
    for (uint32_t b = 0; b < n_parts; ++b) mx = std::max(mx, sent.p[b + 1] - sent.p[b]);
    if (mx == 0) return;
    dim3 grid((mx + kRouteThreads - 1) / kRouteThreads, n_parts);
    unroute_counts_kernel<<<grid, kRouteThreads, 0, stream>>>(d_send_blocks, block_bytes, stride, d_ret_blocks, ret_bytes, capacity, sent,
                                                              d_counts, d_status);
}

void launch_unroute_place(uint8_t *d_send_blocks, uint64_t block_bytes, uint32_t stride, uint8_t *d_ret_blocks, uint64_t ret_bytes,
                          uint32_t capacity, uint32_t n_parts, const RoutePrefix &sent, const RoutePrefix64 &seg, const uint32_t *d_recv_pos,
                          const uint64_t *d_offsets, uint32_t *d_positions, cudaStream_t stream) {
    uint32_t mx = 0;
    for (uint32_t b = 0; b < n_parts; ++b) mx = std::max(mx, sent.p[b + 1] - sent.p[b]);
    if (mx == 0) return;
    dim3 grid((mx + kRouteThreads - 1) / kRouteThreads, n_parts);
    unroute_place_kernel<<<grid, kRouteThreads, 0, stream>>>(d_send_blocks, block_bytes, stride, d_ret_blocks, ret_bytes, capacity, sent, seg,
                                                             d_recv_pos, d_offsets, d_positions);
}

// ---- presence bitmap of an index part: bit h = some k-mer of the text has hash h -------------------------------------
__global__ void __launch_bounds__(256) presence_bits_kernel(const uint32_t *__restrict__ dir, uint64_t n_keys, uint32_t *__restrict__ bits32) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool present = h < n_keys && dir[h + 1] > dir[h];
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, present);
    if ((threadIdx.x & 31) == 0 && h < ((n_keys + 31) / 32) * 32) bits32[h >> 5] = b;
}

void launch_presence_bits(const uint32_t *d_dir, uint64_t n_keys, uint64_t *d_bits_at_lo, cudaStream_t stream) {
    if (n_keys == 0) return;
    const uint64_t threads = (n_keys + 31) / 32 * 32;
    presence_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_dir, n_keys, reinterpret_cast<uint32_t *>(d_bits_at_lo));
}


// ---- a directory part as bucket SIZES (1 byte per hash instead of a 4-byte offset) -------------------------------------
// The replicated multi-GPU build all-gathers every part's directory; on a genome-like text almost every bucket holds a
// handful of k-mers, so the parts travel as one byte per hash (a quarter of the bytes over NVLink) and every GPU rebuilds
// the whole directory with a prefix sum. A part with a bucket of 255 or more k-mers reports it and the caller ships
// that index's directory uncompressed.
__global__ void __launch_bounds__(256) bucket_sizes_kernel(const uint32_t *__restrict__ dir, uint64_t n_keys, uint8_t *__restrict__ sizes,
                                                           unsigned long long *__restrict__ n_large) {
    const uint64_t h0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (h0 >= n_keys) return;
    uint32_t large = 0, packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t c = 0;
        if (h0 + j < n_keys) c = dir[h0 + j + 1] - dir[h0 + j];
        if (c >= 255) {
            c = 255;
            ++large;
        }
        packed |= c << (8 * j);
    }
    if (h0 + 4 <= n_keys)
        *reinterpret_cast<uint32_t *>(sizes + h0) = packed;  // n_keys parts start on multiples of 64 hashes: aligned
    else
        for (int j = 0; h0 + j < n_keys; ++j) sizes[h0 + j] = (uint8_t)(packed >> (8 * j));
    if (large) atomicAdd(n_large, (unsigned long long)large);
}

void launch_bucket_sizes(const uint32_t *d_dir, uint64_t n_keys, uint8_t *d_sizes, unsigned long long *d_n_large, cudaStream_t stream) {
    if (n_keys == 0) return;
    const uint64_t threads = (n_keys + 3) / 4;
    bucket_sizes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_dir, n_keys, d_sizes, d_n_large);
}

constexpr int kSizesTile = 256 * 16;  // sizes per CTA (16 per thread, one 16-byte load)

__device__ __forceinline__ uint32_t sum16(uint4 v) {
    // sum of the 16 bytes: pairwise via __vsadu4-free arithmetic (bytes are < 256, the sums fit 16 bits per lane pair)
    uint32_t s = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (w[i] & 0xFF) + ((w[i] >> 8) & 0xFF) + ((w[i] >> 16) & 0xFF) + (w[i] >> 24);
    return s;
}

__global__ void __launch_bounds__(256) sizes_tile_sum_kernel(const uint8_t *__restrict__ sizes, uint64_t n, uint64_t *__restrict__ tile_sums) {
    __shared__ uint32_t warp_sums[8];
    const uint64_t i0 = (uint64_t)blockIdx.x * kSizesTile + (uint64_t)threadIdx.x * 16;
    uint32_t s = 0;
    if (i0 + 16 <= n) {
        s = sum16(*reinterpret_cast<const uint4 *>(sizes + i0));
    } else {
        for (uint64_t i = i0; i < n; ++i) s += sizes[i];
    }
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += warp_sums[w];
        tile_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) sizes_to_dir_kernel(const uint8_t *__restrict__ sizes, uint64_t n, const uint64_t *__restrict__ tile_off,
                                                           uint32_t *__restrict__ dir) {
    __shared__ uint32_t warp_sums[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t i0 = (uint64_t)blockIdx.x * kSizesTile + (uint64_t)threadIdx.x * 16;
    uint8_t b[16];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = 0;
    if (i0 + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sizes + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) b[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
    } else {
        for (int j = 0; j < 16 && i0 + j < n; ++j) b[j] = sizes[i0 + j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += b[j];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t run = (uint32_t)tile_off[blockIdx.x] + incl - s;  // offsets are < 2^32 (32-bit positions)
    for (int w = 0; w < warp; ++w) run += warp_sums[w];
    if (i0 + 16 <= n) {
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            o[j] = run;
            run += b[j];
        }
        uint4 *dst = reinterpret_cast<uint4 *>(dir + i0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {
        for (int j = 0; j < 16 && i0 + j < n; ++j) {
            dir[i0 + j] = run;
            run += b[j];
        }
    }
}

uint64_t sizes_tiles(uint64_t n) { return (n + kSizesTile - 1) / kSizesTile; }

void launch_sizes_tile_sums(const uint8_t *d_sizes, uint64_t n, uint64_t *d_tile_sums, cudaStream_t stream) {
    if (n) sizes_tile_sum_kernel<<<(unsigned)sizes_tiles(n), 256, 0, stream>>>(d_sizes, n, d_tile_sums);
}
void launch_sizes_to_dir(const uint8_t *d_sizes, uint64_t n, const uint64_t *d_tile_off, uint32_t *d_dir, cudaStream_t stream) {
    if (n) sizes_to_dir_kernel<<<(unsigned)sizes_tiles(n), 256, 0, stream>>>(d_sizes, n, d_tile_off, d_dir);
}

}  // namespace kb
