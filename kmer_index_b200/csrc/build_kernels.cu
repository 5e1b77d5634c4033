// Index construction kernels (sm_100a): the replacement for kmer_index_element::create
// (kmer_index.hpp:154-179). Pipeline per element (one k):
//
//   pack_text            bytes (1 B/symbol) -> b-bit packed words                      [once per index]
//   radix_hist<Text>     per-tile digit histogram, keys recomputed from the packed text (0.25 B/base)
//   column_scan_*        exclusive scan of the [tile][digit] matrix in digit-major order
//   radix_scatter<Text>  pass 0: hash + stable scatter of (hash, position); positions are generated
//   radix_hist<Keys> / radix_scatter<Keys>   passes 1..d-1 over (hash, position) pairs
//   directory_fill       run boundaries of the sorted hashes -> bucket directory (CSR offsets)
//
// Every pass is a streaming kernel bounded by HBM bandwidth; DESIGN.md lists the algorithmic bytes.
#include "radix.cuh"
#include "launch.h"

namespace kb {

// ------------------------------------------------------------------------------------------------
// pack_text: one thread produces one 64-bit word from 64/bits input bytes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_text_kernel(const uint8_t *__restrict__ ranks, uint64_t n, uint32_t bits,
                                                        uint32_t sigma, uint64_t n_words, uint64_t *__restrict__ words,
                                                        uint32_t *__restrict__ error_flag) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint32_t spw = 64 / bits;
    const uint64_t s0 = w * spw;
    uint64_t word = 0;
    bool bad = false;
    if (s0 + spw <= n && ((uintptr_t)(ranks + s0) & 7) == 0) {
        // aligned fast path: 8 bytes at a time
        const uint64_t *src = reinterpret_cast<const uint64_t *>(ranks + s0);
        for (uint32_t j = 0; j < spw; j += 8) {
            const uint64_t v = src[j >> 3];
#pragma unroll
            for (uint32_t b = 0; b < 8; ++b) {
                const uint32_t r = (uint32_t)(v >> (8 * b)) & 0xFF;
                bad |= r >= sigma;
                word |= (uint64_t)r << (64 - bits * (j + b + 1));
            }
        }
    } else {
        for (uint32_t j = 0; j < spw && s0 + j < n; ++j) {
            const uint32_t r = ranks[s0 + j];
            bad |= r >= sigma;
            word |= (uint64_t)r << (64 - bits * (j + 1));
        }
    }
    words[w] = word;
    if (bad) atomicOr(error_flag, 1u);
}

void launch_pack_text(const uint8_t *d_ranks, uint64_t n, uint32_t bits, uint32_t sigma, uint64_t n_words_total,
                      uint64_t *d_words, uint32_t *d_error_flag, cudaStream_t stream) {
    // n_words_total includes the two padding words; threads past the text write zeros
    const uint64_t blocks = (n_words_total + 255) / 256;
    pack_text_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_ranks, n, bits, sigma, n_words_total, d_words, d_error_flag);
}

// ------------------------------------------------------------------------------------------------
// key sources
// ------------------------------------------------------------------------------------------------
// A source hands a thread its kSortItems tile items: item r is element first + 32 * r; `avail` = number of
// valid elements starting at `first` (only consulted when FULL is false). KeyT is uint32_t while sigma^k <= 2^32.
template <typename KeyT>
struct TextSource {  // key(i) = hash of the k-mer starting at symbol i; value(i) = i
    using key_type = KeyT;
    static constexpr uint32_t kAtomicItems = 0xFFFFu;  // see tile_rank: the hash extraction already loads the ALU pipe
    PackedText text;
    uint32_t k;
    __device__ __forceinline__ KeyT key(uint64_t i) const {
        return (KeyT)key_from_window(window64(text.words, i, text.bits), k, text.bits, text.sigma);
    }
    template <bool FULL>
    __device__ __forceinline__ void load_keys(uint64_t first, uint32_t avail, KeyT (&out)[kSortItems]) const {
        // items are 32 symbols = 32 * bits bits = `step` whole words apart, so the bit offset inside the word is
        // the same for all of them: walk the words once (step + 1 loads serve an item, step == 1 shares one)
        const uint64_t bit0 = first * text.bits;
        const uint64_t *w = text.words + (bit0 >> 6);
        const uint32_t sh = (uint32_t)(bit0 & 63);
        const uint32_t step = text.bits >> 1;
        const bool pow2 = text.sigma == (1u << text.bits);
        const uint32_t down = 64 - k * text.bits;
        // power-of-two alphabet and k * bits <= 32 (dna4 k <= 16): the key is the top of ONE 32-bit field of the
        // three 32-bit pieces (hi(cur), lo(cur), hi(next)) -- one funnel shift instead of a 64-bit window
        const bool narrow = pow2 && k * text.bits <= 32;
        const bool low = sh >= 32;
        uint64_t cur = w[0];
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            // a valid item's two words are inside the (padded) text; invalid items read nothing
            const bool valid = FULL || (uint32_t)(r * 32) < avail;
            const uint64_t next = valid ? w[r * step + 1] : 0ull;
            if (narrow) {
                const uint32_t a = low ? (uint32_t)cur : (uint32_t)(cur >> 32);
                const uint32_t b = low ? (uint32_t)(next >> 32) : (uint32_t)cur;
                out[r] = valid ? (KeyT)(__funnelshift_l(b, a, sh) >> (down - 32)) : (KeyT)0;
                if (r + 1 < kSortItems) {
                    const bool valid_next = FULL || (uint32_t)((r + 1) * 32) < avail;
                    cur = step == 1 ? next : (valid_next ? w[(r + 1) * step] : 0ull);
                }
                continue;
            }
            const uint64_t win = sh ? ((cur << sh) | (next >> (64 - sh))) : cur;
            const KeyT kk = pow2 ? (KeyT)(win >> down) : (KeyT)key_from_window(win, k, text.bits, text.sigma);
            out[r] = valid ? kk : (KeyT)0;
            if (r + 1 < kSortItems) {
                const bool valid_next = FULL || (uint32_t)((r + 1) * 32) < avail;
                cur = step == 1 ? next : (valid_next ? w[(r + 1) * step] : 0ull);
            }
        }
    }
    template <bool FULL>
    __device__ __forceinline__ void load_vals(uint64_t first, uint32_t, uint32_t (&out)[kSortItems]) const {
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) out[r] = (uint32_t)first + r * 32;
    }
    template <bool FULL>
    __device__ __forceinline__ void prefetch_vals(uint64_t, uint32_t) const {}
};

// k-mers wider than one 64-bit window (k * bits > 64: aa27 k >= 9, dna5 k >= 17 ...; keys are then always 64-bit):
// every key is a Horner walk over two windows, so there is no word sharing between a thread's items to exploit.
struct WideTextSource {
    using key_type = uint64_t;
    static constexpr uint32_t kAtomicItems = 0xFFFFu;
    PackedText text;
    uint32_t k;
    __device__ __forceinline__ uint64_t key(uint64_t i) const { return key_at(text.words, i, k, text.bits, text.sigma); }
    template <bool FULL>
    __device__ __forceinline__ void load_keys(uint64_t first, uint32_t avail, uint64_t (&out)[kSortItems]) const {
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) out[r] = (FULL || (uint32_t)(r * 32) < avail) ? key(first + r * 32) : 0ull;
    }
    template <bool FULL>
    __device__ __forceinline__ void load_vals(uint64_t first, uint32_t, uint32_t (&out)[kSortItems]) const {
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) out[r] = (uint32_t)first + r * 32;
    }
    template <bool FULL>
    __device__ __forceinline__ void prefetch_vals(uint64_t, uint32_t) const {}
};

template <typename KeyT>
struct PairSource {  // materialised (key, value) pairs
    using key_type = KeyT;
#ifndef KB_PAIR_ATOMIC_ITEMS
#define KB_PAIR_ATOMIC_ITEMS 0u  // tuning: which of a thread's 16 items are matched through shared memory in the pair passes
#endif
    static constexpr uint32_t kAtomicItems = KB_PAIR_ATOMIC_ITEMS;
    const KeyT *keys;
    const uint32_t *vals;
    __device__ __forceinline__ KeyT key(uint64_t i) const { return keys[i]; }
    template <bool FULL>
    __device__ __forceinline__ void load_keys(uint64_t first, uint32_t avail, KeyT (&out)[kSortItems]) const {
        const KeyT *p = keys + first;  // one base pointer; the unrolled loads use immediate offsets
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) out[r] = (FULL || (uint32_t)(r * 32) < avail) ? p[r * 32] : (KeyT)0;
    }
    template <bool FULL>
    __device__ __forceinline__ void load_vals(uint64_t first, uint32_t avail, uint32_t (&out)[kSortItems]) const {
        const uint32_t *p = vals + first;
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) out[r] = (FULL || (uint32_t)(r * 32) < avail) ? p[r * 32] : 0u;
    }
    // values are only needed after the ranking; pull their lines into L2 now so the later loads are short
    template <bool FULL>
    __device__ __forceinline__ void prefetch_vals(uint64_t first, uint32_t avail) const {
        const uint32_t *p = vals + first;
#pragma unroll
        for (int r = 0; r < kSortItems; ++r)
            if (FULL || (uint32_t)(r * 32) < avail) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + r * 32));
    }
};

// ------------------------------------------------------------------------------------------------
// radix_hist: tile_hist[tile][digit] = number of elements of the tile with that digit
// ------------------------------------------------------------------------------------------------
template <typename Source>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(Source src, uint64_t n, uint32_t shift, uint32_t mask,
                                                                  uint32_t *__restrict__ tile_hist) {
    // one private histogram per warp: shared-memory atomics only collide inside a warp
    __shared__ uint32_t hist[kSortWarps][kRadix];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t tile_begin = (uint64_t)blockIdx.x * kSortTile;
    uint32_t d[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint64_t i = tile_begin + (uint64_t)r * kSortThreads + tid;
        d[r] = i < n ? ((uint32_t)(src.key(i) >> shift) & mask) : kInvalidDigit;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        // a warp whose 32 digits are all equal (low-entropy text) adds once instead of colliding 32 times
        const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, d[r], 0);
        if (__all_sync(0xFFFFFFFFu, d[r] == d0)) {
            if (lane == 0 && d0 != kInvalidDigit) hist[warp][d0] += 32;
        } else if (d[r] != kInvalidDigit) {
            atomicAdd(&hist[warp][d[r]], 1u);
        }
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s += hist[w][tid];
        tile_hist[(uint64_t)blockIdx.x * kRadix + tid] = s;
    }
}

// The same for materialised 32-bit keys, the three later passes of a dna4 k = 16 build: the generic kernel spends
// 25 warp instructions per 32 keys and is instruction-bound (80 % of the issue slots at 4.9 TB/s). Here a thread
// takes four consecutive keys per 16-byte load and the all-equal shortcut is tested once per load, not per key.
__global__ void __launch_bounds__(kSortThreads) radix_hist_keys32_kernel(const uint32_t *__restrict__ keys, uint64_t n,
                                                                         uint32_t shift, uint32_t mask,
                                                                         uint32_t *__restrict__ tile_hist) {
    __shared__ uint32_t hist[kSortWarps][kRadix];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t tile_begin = (uint64_t)blockIdx.x * kSortTile;
    uint32_t *my = hist[warp];
    if (tile_begin + kSortTile <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(keys + tile_begin) + tid;  // tiles start 32 KB-aligned
        uint4 v[kSortItems / 4];
#pragma unroll
        for (int g = 0; g < kSortItems / 4; ++g) v[g] = p[g * kSortThreads];
#pragma unroll
        for (int g = 0; g < kSortItems / 4; ++g) {
            const uint32_t d0 = (v[g].x >> shift) & mask, d1 = (v[g].y >> shift) & mask;
            const uint32_t d2 = (v[g].z >> shift) & mask, d3 = (v[g].w >> shift) & mask;
            // a warp whose 128 digits are all equal (low-entropy text) adds once instead of colliding
            const uint32_t first = __shfl_sync(0xFFFFFFFFu, d0, 0);
            if (__all_sync(0xFFFFFFFFu, ((d0 ^ first) | (d1 ^ first) | (d2 ^ first) | (d3 ^ first)) == 0)) {
                if (lane == 0) my[first] += 128;
            } else {
                atomicAdd(&my[d0], 1u);
                atomicAdd(&my[d1], 1u);
                atomicAdd(&my[d2], 1u);
                atomicAdd(&my[d3], 1u);
            }
        }
    } else {
        for (uint64_t i = tile_begin + tid; i < n; i += kSortThreads) atomicAdd(&my[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s += hist[w][tid];
        tile_hist[(uint64_t)blockIdx.x * kRadix + tid] = s;
    }
}

// Text-sourced histogram: a thread walks kSortItems CONSECUTIVE positions and slides one 64-bit window over
// them (one window load serves 64/bits - k + 1 k-mers), instead of re-reading two words per k-mer.
__global__ void __launch_bounds__(kSortThreads) radix_hist_text_kernel(PackedText text, uint32_t k, uint64_t n,
                                                                       uint32_t shift, uint32_t mask,
                                                                       uint32_t *__restrict__ tile_hist) {
    __shared__ uint32_t hist[kSortWarps][kRadix];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t i0 = (uint64_t)blockIdx.x * kSortTile + (uint64_t)tid * kSortItems;
    const uint32_t per_window = 64 / text.bits - k + 1;  // k-mers fully inside one window
    uint32_t j = 0;
    while (j < (uint32_t)kSortItems && i0 + j < n) {
        uint64_t w = window64(text.words, i0 + j, text.bits);
        const uint32_t cnt = min(min(per_window, (uint32_t)kSortItems - j), (uint32_t)min((uint64_t)kSortItems, n - i0 - j));
        for (uint32_t t = 0; t < cnt; ++t) {
            const uint32_t d = (uint32_t)(key_from_window(w, k, text.bits, text.sigma) >> shift) & mask;
            atomicAdd(&hist[warp][d], 1u);
            w <<= text.bits;
        }
        j += cnt;
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s += hist[w][tid];
        tile_hist[(uint64_t)blockIdx.x * kRadix + tid] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// column scan: in place, tile_hist[tile][digit] -> global exclusive offset of (digit, tile) in
// digit-major order, i.e. base(d) + sum_{t' < tile} hist[t'][d].
// Three small kernels over the [n_tiles][256] matrix, all with 256 threads = one per digit so that
// every access is a coalesced 1 KB row.
// ------------------------------------------------------------------------------------------------
constexpr int kScanChunk = 64;  // tiles per chunk

__global__ void __launch_bounds__(kRadix) column_sum_kernel(const uint32_t *__restrict__ tile_hist, uint32_t n_tiles,
                                                            uint32_t *__restrict__ chunk_sums) {
    const uint32_t t0 = blockIdx.x * kScanChunk;
    const uint32_t t1 = min(t0 + kScanChunk, n_tiles);
    uint32_t s = 0;
#pragma unroll 8
    for (uint32_t t = t0; t < t1; ++t) s += tile_hist[(uint64_t)t * kRadix + threadIdx.x];
    chunk_sums[(uint64_t)blockIdx.x * kRadix + threadIdx.x] = s;
}

constexpr int kBaseSlices = 4;  // column_base: 4 x 256 threads, each slice walks a quarter of the chunks

__global__ void __launch_bounds__(kRadix * kBaseSlices) column_base_kernel(uint32_t *__restrict__ chunk_sums, uint32_t n_chunks) {
    __shared__ uint32_t warp_sums[kRadix / 32];
    __shared__ uint32_t slice_tot[kBaseSlices][kRadix];
    const int d = threadIdx.x & (kRadix - 1);
    const int slice = threadIdx.x / kRadix;
    const uint32_t per = (n_chunks + kBaseSlices - 1) / kBaseSlices;
    const uint32_t c0 = min(slice * per, n_chunks), c1 = min(c0 + per, n_chunks);
    // per-slice digit totals
    uint32_t total = 0;
#pragma unroll 8
    for (uint32_t c = c0; c < c1; ++c) total += chunk_sums[(uint64_t)c * kRadix + d];
    slice_tot[slice][d] = total;
    __syncthreads();
    uint32_t before = 0, digit_total = 0;  // chunks of earlier slices; all chunks
#pragma unroll
    for (int sl = 0; sl < kBaseSlices; ++sl) {
        const uint32_t t = slice_tot[sl][d];
        if (sl < slice) before += t;
        digit_total += t;
    }
    // exclusive scan of the digit totals over the 256 digits (done redundantly by every slice)
    const int lane = d & 31, warp = d >> 5;
    uint32_t incl = digit_total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (slice == 0 && lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = incl - digit_total;
    for (int w = 0; w < warp; ++w) base += warp_sums[w];
    // chunk_sums[c][d] <- base(d) + sum of earlier chunks
    uint32_t run = base + before;
    for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t v = chunk_sums[(uint64_t)c * kRadix + d];
        chunk_sums[(uint64_t)c * kRadix + d] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(kRadix) column_apply_kernel(uint32_t *__restrict__ tile_hist, uint32_t n_tiles,
                                                              const uint32_t *__restrict__ chunk_sums) {
    const uint32_t t0 = blockIdx.x * kScanChunk;
    const uint32_t t1 = min(t0 + kScanChunk, n_tiles);
    uint32_t run = chunk_sums[(uint64_t)blockIdx.x * kRadix + threadIdx.x];
    uint32_t v[8];
    for (uint32_t t = t0; t < t1; t += 8) {
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) v[j] = (t + j < t1) ? tile_hist[(uint64_t)(t + j) * kRadix + threadIdx.x] : 0;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            if (t + j < t1) tile_hist[(uint64_t)(t + j) * kRadix + threadIdx.x] = run;
            run += v[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// radix_scatter: stable scatter of one tile. Elements are ranked in registers, staged in shared memory
// in digit order and written out as per-digit runs (coalesced: a run of a digit is contiguous in the
// destination).
// ------------------------------------------------------------------------------------------------
// (key, value) staged in digit order. 32-bit keys: one 64-bit shared-memory access per element.
template <typename KeyT, bool ATOMIC>
struct ScatterSmem;
template <bool ATOMIC>
struct ScatterSmem<uint32_t, ATOMIC> {
    RankSmem rank;
    uint32_t warp_msk[ATOMIC ? kSortWarps : 1][kRadix];  // only the shared-memory match (tile_rank) uses it
    uint32_t delta[kRadix];
    uint2 kv[kSortTile];
    __device__ __forceinline__ void put(uint32_t i, uint32_t k, uint32_t v) { kv[i] = make_uint2(k, v); }
    __device__ __forceinline__ void get(uint32_t i, uint32_t &k, uint32_t &v) const {
        const uint2 e = kv[i];
        k = e.x;
        v = e.y;
    }
};
template <bool ATOMIC>
struct ScatterSmem<uint64_t, ATOMIC> {
    RankSmem rank;
    uint32_t warp_msk[ATOMIC ? kSortWarps : 1][kRadix];
    uint32_t delta[kRadix];
    uint64_t keys[kSortTile];
    uint32_t vals[kSortTile];
    __device__ __forceinline__ void put(uint32_t i, uint64_t k, uint32_t v) {
        keys[i] = k;
        vals[i] = v;
    }
    __device__ __forceinline__ void get(uint32_t i, uint64_t &k, uint32_t &v) const {
        k = keys[i];
        v = vals[i];
    }
};

template <typename Source, int BITS, bool FULL, bool BYTE>
__device__ __forceinline__ void scatter_tile(const Source &src, uint64_t tile_begin, uint32_t count, uint32_t shift,
                                             uint32_t mask, const uint32_t *__restrict__ tile_base_row,
                                             typename Source::key_type *__restrict__ out_keys,
                                             uint32_t *__restrict__ out_vals,
                                             ScatterSmem<typename Source::key_type, (Source::kAtomicItems != 0)> &sm) {
    using KeyT = typename Source::key_type;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t e0 = (uint32_t)(warp * kSortItems * 32 + lane);

    // keys live in registers through the ranking; values are fetched afterwards (prefetched to L2 meanwhile),
    // which keeps the 32-bit kernel at <= 64 registers so two CTAs share an SM and overlap each other's phases
    KeyT key[kSortItems];
    uint32_t local_pos2[kSortItems / 2];  // two 16-bit tile positions per register
    src.template load_keys<FULL>(tile_begin + e0, count - min(count, e0), key);
    src.template prefetch_vals<FULL>(tile_begin + e0, count - min(count, e0));
#ifdef KB_ABL_NORANK  // timing ablation (profiles/README.md): identity placement instead of ranking
    for (int r = 0; r < kSortItems; ++r) local_pos2[r >> 1] = (r & 1) ? (local_pos2[r >> 1] | ((e0 + r * 32) << 16)) : (e0 + r * 32);
    if (tid < kRadix) sm.delta[tid] = (uint32_t)tile_begin;
#else
    tile_rank<BITS, FULL, KeyT, BYTE, Source::kAtomicItems>(key, count, shift, mask, local_pos2, sm.rank, sm.warp_msk);
    if (tid < kRadix) sm.delta[tid] = tile_base_row[tid] - sm.rank.excl[tid];
#endif
    {
        uint32_t val[kSortItems];
        src.template load_vals<FULL>(tile_begin + e0, count - min(count, e0), val);
#pragma unroll
        for (int r = 0; r < kSortItems; ++r)
            if (FULL || e0 + r * 32 < count) sm.put((local_pos2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu, key[r], val[r]);
    }
    __syncthreads();
#ifdef KB_ABL_NOOUT  // timing ablation (profiles/README.md): skip the global stores
    if (tile_begin == 0xFFFFFFFFFFull)
#endif
    if (FULL) {
#pragma unroll
        for (int it = 0; it < kSortItems; ++it) {
            const uint32_t j = it * kSortThreads + tid;
            KeyT kk;
            uint32_t vv;
            sm.get(j, kk, vv);
            const uint32_t dst = sm.delta[digit_of<BYTE>(kk, shift, mask)] + j;  // mod 2^32; true destination < n < 2^32
            out_keys[dst] = kk;
            out_vals[dst] = vv;
        }
    } else {
        for (uint32_t j = tid; j < count; j += kSortThreads) {
            KeyT kk;
            uint32_t vv;
            sm.get(j, kk, vv);
            const uint32_t dst = sm.delta[digit_of<BYTE>(kk, shift, mask)] + j;
            out_keys[dst] = kk;
            out_vals[dst] = vv;
        }
    }
}

template <typename Source, int BITS, bool BYTE>
__global__ void __launch_bounds__(kSortThreads, sizeof(typename Source::key_type) == 4 ? 2 : 1)
    radix_scatter_kernel(const Source src, uint64_t n, uint32_t shift, uint32_t mask,
                         const uint32_t *__restrict__ tile_base, typename Source::key_type *__restrict__ out_keys,
                         uint32_t *__restrict__ out_vals) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto &sm = *reinterpret_cast<ScatterSmem<typename Source::key_type, (Source::kAtomicItems != 0)> *>(smem_raw);
    const uint64_t tile_begin = (uint64_t)blockIdx.x * kSortTile;
    const uint32_t count = (uint32_t)min((uint64_t)kSortTile, n - tile_begin);
    const uint32_t *row = tile_base + (uint64_t)blockIdx.x * kRadix;
    if (count == (uint32_t)kSortTile)
        scatter_tile<Source, BITS, true, BYTE>(src, tile_begin, count, shift, mask, row, out_keys, out_vals, sm);
    else
        scatter_tile<Source, BITS, false, BYTE>(src, tile_begin, count, shift, mask, row, out_keys, out_vals, sm);
}

// ------------------------------------------------------------------------------------------------
// directory_fill: dir[j] = number of sorted keys with (key >> shift) < j, for j in [0, dir_entries).
// With S(i) = (key[i] >> shift) + 1, S(-1) = 0 and S(n) = dir_entries, boundary i (between sorted elements i-1
// and i) owns the run dir[S(i-1) .. S(i)) = i. A CTA takes kDirThreads * ITEMS consecutive boundaries, whose runs
// tile one contiguous directory range. That range is produced a slab at a time in shared memory: every
// non-empty run drops its value at its first entry, and because the values grow with the position a "last
// non-zero so far" scan (one ballot + two shuffles per 32 entries) turns the heads into the filled range,
// which leaves as full coalesced lines. The work per boundary does not depend on the run lengths, so sparse
// key spaces and low-entropy texts cost the same per directory entry as dense ones.
// ------------------------------------------------------------------------------------------------
constexpr int kDirThreads = 256;
constexpr int kDirWarps = kDirThreads / 32;

// head[j]: offset of run j's first entry from the start of the CTA's range, ~0 for empty runs; run j carries the
// value value0 + j. The first slab's shared memory arrives zeroed. out = dir + start of the CTA's range.
template <typename RelT, int ITEMS>
__device__ __forceinline__ void directory_rounds(uint32_t *slab, uint32_t *warp_last, const RelT (&head)[ITEMS],
                                                 uint32_t value0, RelT span, uint32_t *__restrict__ out) {
    constexpr uint32_t kSlab = 4 * kDirThreads * ITEMS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);  // lanes <= this one
    uint32_t carry = 0;  // value in force at the end of the previous slab (boundary indices only grow)
    for (RelT c0 = 0; c0 < span; c0 += kSlab) {
        const uint32_t cn = (uint32_t)min((RelT)(span - c0), (RelT)kSlab);
        // the slab's cn entries are split evenly between the warps, in granules of 128 entries (one 16-byte load per lane)
        const uint32_t seg_len = (cn + 128 * kDirWarps - 1) / (128 * kDirWarps) * 128;
        const uint32_t seg0 = warp * seg_len;
        if (c0) {
            for (uint32_t t = tid * 4; t < seg_len * kDirWarps; t += kDirThreads * 4)
                *reinterpret_cast<uint4 *>(slab + t) = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const RelT rel = head[j] - c0;  // wraps to huge when the head is not in this slab (or absent)
            if (rel < (RelT)cn) slab[rel] = value0 + j;
        }
        __syncthreads();
        // each warp: the last head of its segment ...
        uint32_t seg_last = 0;
        for (uint32_t t = 0; t < seg_len; t += 128) {
            const uint4 q = *reinterpret_cast<const uint4 *>(slab + seg0 + t + lane * 4);
            seg_last = max(max(seg_last, q.x), max(max(q.y, q.z), q.w));
        }
        seg_last = __reduce_max_sync(0xFFFFFFFFu, seg_last);
        if (lane == 0) warp_last[warp] = seg_last;
        __syncthreads();
        uint32_t run = carry;
#pragma unroll
        for (int w = 0; w < kDirWarps; ++w) {
            if (w < warp) run = max(run, warp_last[w]);
            carry = max(carry, warp_last[w]);
        }
        // ... then the segment itself: an entry takes the last head at or before it, else what was in force before
        const uint32_t n_here = seg0 < cn ? min(cn - seg0, seg_len) : 0u;
        const uint32_t *in = slab + seg0 + lane;
        uint32_t *o = out + c0 + seg0 + lane;
        // one group of 32 entries: v = this lane's slab entry (entries past cn are zero)
        auto resolve = [&](uint32_t v) -> uint32_t {
            const uint32_t nz = __ballot_sync(0xFFFFFFFFu, v != 0);
            const uint32_t below = nz & le_mask;
            uint32_t src;  // highest head lane at or below this one (bfind gives 0xFFFFFFFF for none: any lane then)
            asm("bfind.u32 %0, %1;" : "=r"(src) : "r"(below));
            const uint32_t got = __shfl_sync(0xFFFFFFFFu, v, src & 31u);
            const uint32_t res = below ? got : run;
            run = __shfl_sync(0xFFFFFFFFu, res, 31);  // what is in force after this group
            return res;
        };
        uint32_t left = n_here;
        for (; left >= 128; left -= 128, in += 128, o += 128) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = in[u * 32];
#pragma unroll
            for (int u = 0; u < 4; ++u) o[u * 32] = resolve(v[u]);
        }
        for (; left >= 32; left -= 32, in += 32, o += 32) o[0] = resolve(in[0]);
        if (left) {  // warp-uniform
            const uint32_t res = resolve(in[0]);
            if ((uint32_t)lane < left) o[0] = res;
        }
        __syncthreads();  // the next round reuses slab and warp_last
    }
}

template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kDirThreads) directory_fill_kernel(const KeyT *__restrict__ keys, uint64_t n_kmers,
                                                                     uint32_t shift, uint64_t dir_entries,
                                                                     uint32_t *__restrict__ dir) {
    constexpr int kTile = kDirThreads * ITEMS;  // boundaries per CTA
    constexpr int kSlab = 4 * kTile;            // directory entries produced per round (the dense directory has <= 4 per key)
    __shared__ __align__(16) uint32_t slab[kSlab];
    __shared__ uint32_t warp_last[kDirWarps];
    __shared__ uint64_t s_end;
    const int tid = threadIdx.x;
    const uint64_t tile_begin = (uint64_t)blockIdx.x * kTile;              // first boundary of the CTA (<= n_kmers)
    const uint64_t tile_end = min(tile_begin + kTile, n_kmers + 1);       // one past its last boundary
    const uint32_t nb = (uint32_t)(tile_end - tile_begin);
    const bool edge = tile_begin == 0 || tile_end == n_kmers + 1;          // S(-1) or S(n) is involved

    const uint64_t i0 = tile_begin + (uint64_t)tid * ITEMS;  // first boundary of this thread
    const KeyT k_base = tile_begin ? keys[tile_begin - 1] : (KeyT)0;  // same address for the whole CTA
    KeyT k[ITEMS + 1];                                                 // keys[i0 - 1 .. i0 + ITEMS - 1]
    k[0] = (i0 >= 1 && i0 - 1 < n_kmers) ? keys[i0 - 1] : (KeyT)0;
    if (i0 + ITEMS <= n_kmers) {
        // aligned vector loads: i0 is a multiple of ITEMS and the array comes from cudaMalloc
        constexpr int kPerVec = 16 / (int)sizeof(KeyT);
        const uint4 *p = reinterpret_cast<const uint4 *>(keys + i0);
#pragma unroll
        for (int v = 0; v < ITEMS / kPerVec; ++v) {
            const uint4 q = p[v];
            if (sizeof(KeyT) == 4) {
                k[1 + 4 * v] = (KeyT)q.x;
                k[2 + 4 * v] = (KeyT)q.y;
                k[3 + 4 * v] = (KeyT)q.z;
                k[4 + 4 * v] = (KeyT)q.w;
            } else {
                k[1 + 2 * v] = (KeyT)(((uint64_t)q.y << 32) | q.x);
                k[2 + 2 * v] = (KeyT)(((uint64_t)q.w << 32) | q.z);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) k[j + 1] = (i0 + j < n_kmers) ? keys[i0 + j] : (KeyT)0;
    }
    // the first slab is zeroed while the loads are in flight
#pragma unroll
    for (int t = 0; t < kSlab / (kDirThreads * 4); ++t)
        *reinterpret_cast<uint4 *>(slab + (t * kDirThreads + tid) * 4) = make_uint4(0, 0, 0, 0);

    const uint32_t b0 = (uint32_t)tid * ITEMS;  // this thread's first boundary inside the tile
    if (!edge) {
        // interior tile: every S is (key >> shift) + 1 with (key >> shift) < 2^32, so offsets from the tile's first
        // entry are exact 32-bit differences of the shifted keys
        const uint32_t slot_base = (uint32_t)(k_base >> shift);
        uint32_t rel[ITEMS + 1];
#pragma unroll
        for (int j = 0; j <= ITEMS; ++j) rel[j] = (uint32_t)(k[j] >> shift) - slot_base;
        uint32_t head[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) head[j] = rel[j + 1] > rel[j] ? rel[j] : 0xFFFFFFFFu;  // all kTile boundaries exist
        if (tid == kDirThreads - 1) s_end = rel[ITEMS];
        __syncthreads();
        directory_rounds<uint32_t, ITEMS>(slab, warp_last, head, (uint32_t)i0, (uint32_t)s_end,
                                          dir + ((uint64_t)slot_base + 1));
    } else {
        const uint64_t base = tile_begin ? (uint64_t)(k_base >> shift) + 1 : 0;
        uint64_t head[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint64_t i = i0 + j;
            head[j] = ~0ull;
            if (b0 + j < nb) {
                const uint64_t lo = (i == 0) ? 0 : (uint64_t)(k[j] >> shift) + 1;                     // S(i - 1)
                const uint64_t hi = (i == n_kmers) ? dir_entries : (uint64_t)(k[j + 1] >> shift) + 1;  // S(i)
                if (hi > lo) head[j] = lo - base;
                if (b0 + j == nb - 1) s_end = hi - base;
            }
        }
        __syncthreads();
        directory_rounds<uint64_t, ITEMS>(slab, warp_last, head, (uint32_t)i0, s_end, dir + base);
    }
}

template <typename KeyT, int ITEMS>
static void launch_directory_fill_t(const KeyT *d_keys, uint64_t n_kmers, uint32_t shift, uint64_t dir_entries, uint32_t *d_dir,
                                    cudaStream_t stream) {
    const uint64_t tile = (uint64_t)kDirThreads * ITEMS;
    const uint64_t blocks = (n_kmers + 1 + tile - 1) / tile;
    cudaFuncSetAttribute(directory_fill_kernel<KeyT, ITEMS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    directory_fill_kernel<KeyT, ITEMS><<<(unsigned)blocks, kDirThreads, 0, stream>>>(d_keys, n_kmers, shift, dir_entries, d_dir);
}

void launch_directory_fill(const void *d_keys, uint32_t key_bytes, uint64_t n_kmers, uint32_t shift, uint64_t dir_entries,
                           uint32_t *d_dir, cudaStream_t stream) {
    // 8 boundaries per thread: 2048 per CTA and a 32 KB slab (4 per thread measured 30 % slower on config 5)
    if (key_bytes == 8)
        launch_directory_fill_t<uint64_t, 8>((const uint64_t *)d_keys, n_kmers, shift, dir_entries, d_dir, stream);
    else
        launch_directory_fill_t<uint32_t, 8>((const uint32_t *)d_keys, n_kmers, shift, dir_entries, d_dir, stream);
}

// ------------------------------------------------------------------------------------------------
// host-side pass drivers
// ------------------------------------------------------------------------------------------------
template <typename Source, int BITS, bool BYTE>
static void launch_scatter_bits(const Source &src, uint64_t n, uint32_t shift, uint32_t mask, const uint32_t *d_tile_base,
                                typename Source::key_type *d_out_keys, uint32_t *d_out_vals, cudaStream_t stream) {
    using Smem = ScatterSmem<typename Source::key_type, (Source::kAtomicItems != 0)>;
    const uint32_t n_tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    // per-device attribute; cheap enough to set on every launch
    cudaFuncSetAttribute(radix_scatter_kernel<Source, BITS, BYTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    radix_scatter_kernel<Source, BITS, BYTE>
        <<<n_tiles, kSortThreads, sizeof(Smem), stream>>>(src, n, shift, mask, d_tile_base, d_out_keys, d_out_vals);
}

// the vote loop is unrolled for the digit width (rounded up to one of the instantiated widths)
template <typename Source>
static void launch_scatter(const Source &src, uint64_t n, uint32_t shift, uint32_t mask, const uint32_t *d_tile_base,
                           typename Source::key_type *d_out_keys, uint32_t *d_out_vals, cudaStream_t stream) {
    const int bits = __builtin_popcount(mask);
    if (bits <= 5)
        launch_scatter_bits<Source, 5, false>(src, n, shift, mask, d_tile_base, d_out_keys, d_out_vals, stream);
    else if (bits == 6)
        launch_scatter_bits<Source, 6, false>(src, n, shift, mask, d_tile_base, d_out_keys, d_out_vals, stream);
    else if (bits == 7)
        launch_scatter_bits<Source, 7, false>(src, n, shift, mask, d_tile_base, d_out_keys, d_out_vals, stream);
    else if (sizeof(typename Source::key_type) == 4 && (shift & 7) == 0) {
        if constexpr (sizeof(typename Source::key_type) == 4)  // byte-aligned digit of a 32-bit key (dna4 k = 12, 16: every pass)
            launch_scatter_bits<Source, 8, true>(src, n, shift, mask, d_tile_base, d_out_keys, d_out_vals, stream);
    } else
        launch_scatter_bits<Source, 8, false>(src, n, shift, mask, d_tile_base, d_out_keys, d_out_vals, stream);
}

void launch_column_scan(uint32_t *d_tile_hist, uint32_t n_tiles, uint32_t *d_chunk_sums, cudaStream_t stream) {
    const uint32_t n_chunks = (n_tiles + kScanChunk - 1) / kScanChunk;
    column_sum_kernel<<<n_chunks, kRadix, 0, stream>>>(d_tile_hist, n_tiles, d_chunk_sums);
    column_base_kernel<<<1, kRadix * kBaseSlices, 0, stream>>>(d_chunk_sums, n_chunks);
    column_apply_kernel<<<n_chunks, kRadix, 0, stream>>>(d_tile_hist, n_tiles, d_chunk_sums);
}

void launch_hist_text(const PackedText &text, uint32_t k, uint64_t n_kmers, uint32_t shift, uint32_t mask,
                      uint32_t *d_tile_hist, cudaStream_t stream) {
    const uint32_t n_tiles = (uint32_t)((n_kmers + kSortTile - 1) / kSortTile);
    if (k * text.bits > 64) {  // no sliding window: one two-window hash per k-mer
        radix_hist_kernel<<<n_tiles, kSortThreads, 0, stream>>>(WideTextSource{text, k}, n_kmers, shift, mask, d_tile_hist);
        return;
    }
    radix_hist_text_kernel<<<n_tiles, kSortThreads, 0, stream>>>(text, k, n_kmers, shift, mask, d_tile_hist);
}

void launch_hist_pairs(const void *d_keys, uint32_t key_bytes, uint64_t n, uint32_t shift, uint32_t mask,
                       uint32_t *d_tile_hist, cudaStream_t stream) {
    const uint32_t n_tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    if (key_bytes == 8) {
        PairSource<uint64_t> src{(const uint64_t *)d_keys, nullptr};
        radix_hist_kernel<<<n_tiles, kSortThreads, 0, stream>>>(src, n, shift, mask, d_tile_hist);
    } else {
        radix_hist_keys32_kernel<<<n_tiles, kSortThreads, 0, stream>>>((const uint32_t *)d_keys, n, shift, mask, d_tile_hist);
    }
}

void launch_scatter_text(const PackedText &text, uint32_t k, uint32_t key_bytes, uint64_t n_kmers, uint32_t shift,
                         uint32_t mask, const uint32_t *d_tile_base, void *d_out_keys, uint32_t *d_out_vals,
                         cudaStream_t stream) {
    if (k * text.bits > 64)
        launch_scatter(WideTextSource{text, k}, n_kmers, shift, mask, d_tile_base, (uint64_t *)d_out_keys, d_out_vals, stream);
    else if (key_bytes == 8)
        launch_scatter(TextSource<uint64_t>{text, k}, n_kmers, shift, mask, d_tile_base, (uint64_t *)d_out_keys, d_out_vals, stream);
    else
        launch_scatter(TextSource<uint32_t>{text, k}, n_kmers, shift, mask, d_tile_base, (uint32_t *)d_out_keys, d_out_vals, stream);
}

void launch_scatter_pairs(const void *d_keys, const uint32_t *d_vals, uint32_t key_bytes, uint64_t n, uint32_t shift,
                          uint32_t mask, const uint32_t *d_tile_base, void *d_out_keys, uint32_t *d_out_vals,
                          cudaStream_t stream) {
    if (key_bytes == 8)
        launch_scatter(PairSource<uint64_t>{(const uint64_t *)d_keys, d_vals}, n, shift, mask, d_tile_base,
                       (uint64_t *)d_out_keys, d_out_vals, stream);
    else
        launch_scatter(PairSource<uint32_t>{(const uint32_t *)d_keys, d_vals}, n, shift, mask, d_tile_base,
                       (uint32_t *)d_out_keys, d_out_vals, stream);
}

uint32_t sort_tile_size() { return kSortTile; }
uint32_t scan_chunk_tiles() { return kScanChunk; }

}  // namespace kb
