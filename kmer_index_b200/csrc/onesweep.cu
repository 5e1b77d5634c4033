// Single-sweep LSD radix sort of (hash, position) pairs with 32-bit hashes (sm_100a): the sorter behind
// kmer_index_element::create (kmer_index.hpp:154-179) for every element whose sigma^k fits 32 bits.
//
//   digit histograms   ALL passes' global digit histograms come from one pass over the packed text, before any
//                      scatter runs. Power-of-two alphabets: a digit of the hash is a (w / bits)-mer of the text, so
//                      every pass's histogram is the SAME c-mer histogram of the text, corrected at the two ends
//                      (cmer_hist_kernel + digit_bases_kernel). Other alphabets: one generic pass that hashes every
//                      k-mer and counts each of its digits (digit_hist_generic_kernel).
//   scatter            one kernel per pass, persistent CTAs (one per SM, 1024 threads), tiles of 8192 elements
//                      claimed from an atomic counter. A tile's input arrives by ONE bulk async copy
//                      (cp.async.bulk global -> shared, completion on an mbarrier) that is issued while the previous
//                      tile is still being ranked and written, so the load phase overlaps the rest. Pairs travel as
//                      8-byte (hash, position) records between passes: one 64-bit load and one 64-bit store per
//                      element. The tile's global digit offsets come from a decoupled look-back over per-tile
//                      status words (state | pass tag | count), so there is no per-tile histogram pass and no column
//                      scan between passes. The first pass reads the text slice of its tile instead of pairs and
//                      generates the positions; the last pass writes the position array (and the sorted hashes).
// Stability (ascending positions inside every bucket, kmer_index.hpp:165) comes from the warp-synchronous ranking of
// radix.cuh (match_digit) applied in element order.
#include <algorithm>
#include <cstdlib>

#include "radix.cuh"
#include "launch.h"

namespace kb {

constexpr int kOsThreads = 1024;
constexpr int kOsItems = 8;
constexpr int kOsWarps = kOsThreads / 32;
constexpr int kOsTile = kOsThreads * kOsItems;  // 8192
static_assert(kOsTile == kSortTile, "both sorters use 8192-element tiles");

// look-back status word: bits 63..62 state, 61..56 pass tag (never 0: memory is zeroed once per build), 31..0 value
constexpr uint64_t kStAggregate = 1ull << 62;
constexpr uint64_t kStInclusive = 2ull << 62;

struct OsSmem {
    alignas(128) uint2 in[kOsTile + 16];   // bulk-copy destination: the tile's pairs, or its slice of the packed text
    alignas(16) uint2 out[kOsTile];        // the tile in digit order
    uint32_t warp_cnt[kOsWarps][kRadix];   // per-warp digit counters, then warp prefixes (+ digit base)
    uint32_t excl[kRadix];                 // exclusive prefix of the tile's digit counts
    uint32_t count[kRadix];                // the tile's digit counts
    uint32_t delta[kRadix];                // global destination of slot j with digit d = delta[d] + j
    uint32_t warp_sums[kRadix / 32];
    alignas(8) uint64_t mbar;              // completes when the tile's bytes have landed in `in`
    uint32_t next_tile;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// bytes: a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint64_t ld_status(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct OsArgs {
    const uint2 *in_pairs;   // passes > 0
    PackedText text;         // pass 0
    uint32_t k;
    uint64_t n;              // elements
    uint32_t n_tiles;
    uint32_t shift, mask;
    uint32_t tag;            // pass tag of the status words, 1..63
    uint32_t final_pass;     // write out_vals (+ out_keys if non-null) instead of out_pairs
    uint32_t debug_no_lookback;  // timing experiment only (wrong results): skip the wait for the predecessors
    const uint32_t *digit_base;  // [256]: exclusive prefix of this pass's global digit histogram
    uint64_t *status;        // [n_tiles][256]
    uint32_t *tile_counter;  // zero before the launch
    uint2 *out_pairs;
    uint32_t *out_keys, *out_vals;
};

// Stable rank of the thread's kOsItems keys inside the tile (see tile_rank in radix.cuh; this is the same algorithm
// for 32 warps x 8 items). Item r of a thread is tile element (warp * kOsItems + r) * 32 + lane. On return
// local_pos2 holds the items' positions in the digit-sorted tile, two 16-bit values per register.
template <int BITS, bool FULL, bool BYTE>
__device__ __forceinline__ void os_tile_rank(const uint32_t (&key)[kOsItems], uint32_t count, uint32_t shift, uint32_t mask,
                                             uint32_t (&local_pos2)[kOsItems / 2], OsSmem &sm) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t e0 = (uint32_t)(warp * kOsItems * 32 + lane);
    uint32_t *my_cnt = sm.warp_cnt[warp];
#pragma unroll
    for (int i = 0; i < kRadix / 32; ++i) my_cnt[i * 32 + lane] = 0;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        const bool valid = FULL || e0 + r * 32 < count;
        const uint32_t peers = match_digit<BITS>(d, FULL ? 0xFFFFFFFFu : __ballot_sync(0xFFFFFFFFu, valid));
        const uint32_t info = (uint32_t)__popc(peers & lt_mask) | ((uint32_t)__popc(peers) << 5) | ((valid ? 1u : 0u) << 11);
        if (r & 1)
            local_pos2[r >> 1] |= info << 16;
        else
            local_pos2[r >> 1] = info;
    }
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        const uint32_t info = (local_pos2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu;
        const bool valid = FULL || ((info >> 11) & 1u);
        const uint32_t below = info & 31u;
        uint32_t pre = 0;
        if (valid) pre = my_cnt[d];
        __syncwarp();
        if (valid && below == 0) my_cnt[d] = pre + ((info >> 5) & 63u);
        __syncwarp();
        const uint32_t pos = pre + below;
        if (r & 1)
            local_pos2[r >> 1] = (local_pos2[r >> 1] & 0xFFFFu) | (pos << 16);
        else
            local_pos2[r >> 1] = (local_pos2[r >> 1] & 0xFFFF0000u) | pos;
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, total, exclusive prefix over the digits
    if (tid < kRadix) {
        uint32_t run = 0;
#pragma unroll 8
        for (int w = 0; w < kOsWarps; ++w) {
            const uint32_t c = sm.warp_cnt[w][tid];
            sm.warp_cnt[w][tid] = run;
            run += c;
        }
        sm.count[tid] = run;
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) sm.warp_sums[warp] = incl;
        sm.excl[tid] = incl - run;
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t base = sm.excl[tid];
        for (int w = 0; w < warp; ++w) base += sm.warp_sums[w];
        sm.excl[tid] = base;
#pragma unroll 8
        for (int w = 0; w < kOsWarps; ++w) sm.warp_cnt[w][tid] += base;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        if (FULL || e0 + r * 32 < count) local_pos2[r >> 1] += my_cnt[d] << (16 * (r & 1));
    }
}

// bytes of tile `tile`'s input and where they start
template <bool TEXT>
__device__ __forceinline__ void os_issue_load(const OsArgs &a, OsSmem &sm, uint32_t tile) {
    const uint64_t tile_begin = (uint64_t)tile * kOsTile;
    const uint32_t count = (uint32_t)min((uint64_t)kOsTile, a.n - tile_begin);
    const void *src;
    uint32_t bytes;
    if (TEXT) {
        // the tile's k-mers start at symbols [tile_begin, tile_begin + count): words from the tile's first word up to the
        // one holding the last k-mer's last symbol, plus one (the window read of the last item); tile_begin * bits is a
        // multiple of 64 * 256, and the text carries three zero words of padding
        const uint64_t w0 = tile_begin * a.text.bits / 64;
        const uint64_t last_bit = ((uint64_t)count - 1 + a.k) * a.text.bits - 1;  // relative to the tile's first bit
        const uint32_t words = (uint32_t)(last_bit / 64) + 2;
        src = a.text.words + w0;
        bytes = ((words + 1) & ~1u) * 8;
    } else {
        src = a.in_pairs + tile_begin;
        bytes = ((count + 1) & ~1u) * 8;  // the pair arrays carry two elements of padding
    }
    mbar_arrive_expect_tx(&sm.mbar, bytes);
    bulk_load(sm.in, src, bytes, &sm.mbar);
}

template <bool TEXT, int BITS, bool BYTE, bool FULL>
__device__ __forceinline__ void os_tile(const OsArgs &a, OsSmem &sm, uint32_t tile, uint32_t count) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t e0 = (uint32_t)(warp * kOsItems * 32 + lane);
    const uint64_t tile_begin = (uint64_t)tile * kOsTile;

    uint32_t key[kOsItems], val[kOsItems];
    if (TEXT) {
        // items are 32 symbols = 32 * bits bits = `step` whole words apart: the bit offset inside the word is the same for
        // all of them, and consecutive items share a word when step == 1 (2-bit symbols)
        const uint64_t *w = reinterpret_cast<const uint64_t *>(sm.in);
        const uint32_t bits = a.text.bits;
        const uint32_t bit0 = e0 * bits;
        const uint32_t w0 = bit0 >> 6, sh = bit0 & 63;
        const uint32_t step = bits >> 1;
        const bool pow2 = a.text.sigma == (1u << bits);
        const uint32_t down = 64 - a.k * bits;
        uint64_t cur = w[w0];
#pragma unroll
        for (int r = 0; r < kOsItems; ++r) {
            const bool valid = FULL || e0 + r * 32 < count;
            const uint64_t next = w[w0 + r * step + 1];  // inside the staged slice for valid items; stale but harmless otherwise
            const uint64_t win = sh ? ((cur << sh) | (next >> (64 - sh))) : cur;
            const uint32_t kk = pow2 ? (uint32_t)(win >> down) : (uint32_t)key_from_window(win, a.k, bits, a.text.sigma);
            key[r] = valid ? kk : 0u;
            val[r] = (uint32_t)tile_begin + e0 + r * 32;
            if (r + 1 < kOsItems) cur = step == 1 ? next : w[w0 + (r + 1) * step];
        }
    } else {
#pragma unroll
        for (int r = 0; r < kOsItems; ++r) {
            const uint2 kv = sm.in[e0 + r * 32];
            const bool valid = FULL || e0 + r * 32 < count;
            key[r] = valid ? kv.x : 0u;
            val[r] = kv.y;
        }
    }
    __syncthreads();  // every thread has its items: the staging buffer can take the next tile
    if (tid == 0) {
        const uint32_t next = atomicAdd(a.tile_counter, 1u);
        sm.next_tile = next;
        if (next < a.n_tiles) os_issue_load<TEXT>(a, sm, next);
    }

    uint32_t local_pos2[kOsItems / 2];
    os_tile_rank<BITS, FULL, BYTE>(key, count, a.shift, a.mask, local_pos2, sm);

    // decoupled look-back (threads 0..255, one digit each) while the other warps already stage their items
    if (tid < kRadix) {
        const uint32_t cnt = sm.count[tid];
        const uint64_t tag = (uint64_t)a.tag << 56;
        uint64_t *row = a.status + (uint64_t)tile * kRadix;
        uint32_t excl = 0;
        if (tile == 0) {
            st_status(row + tid, kStInclusive | tag | cnt);
        } else {
            st_status(row + tid, kStAggregate | tag | cnt);
            const uint64_t *p = row + tid;
            if (!a.debug_no_lookback)
            for (;;) {
                p -= kRadix;
                uint64_t s;
                do {
                    s = ld_status(p);
                } while (((s >> 56) & 0x3F) != a.tag);  // not published in this pass yet
                excl += (uint32_t)s;
                if (s & kStInclusive) break;
            }
            st_status(row + tid, kStInclusive | tag | (uint32_t)(excl + cnt));
        }
        sm.delta[tid] = a.digit_base[tid] + excl - sm.excl[tid];
    }
#pragma unroll
    for (int r = 0; r < kOsItems; ++r)
        if (FULL || e0 + r * 32 < count) sm.out[(local_pos2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu] = make_uint2(key[r], val[r]);
    __syncthreads();

    if (FULL) {
        if (a.final_pass) {
#pragma unroll
            for (int it = 0; it < kOsItems; ++it) {
                const uint32_t j = it * kOsThreads + tid;
                const uint2 e = sm.out[j];
                const uint32_t dst = sm.delta[digit_of<BYTE>(e.x, a.shift, a.mask)] + j;  // mod 2^32; true destination < n < 2^32
                a.out_vals[dst] = e.y;
                if (a.out_keys != nullptr) a.out_keys[dst] = e.x;
            }
        } else {
#pragma unroll
            for (int it = 0; it < kOsItems; ++it) {
                const uint32_t j = it * kOsThreads + tid;
                const uint2 e = sm.out[j];
                a.out_pairs[sm.delta[digit_of<BYTE>(e.x, a.shift, a.mask)] + j] = e;
            }
        }
    } else {
        for (uint32_t j = tid; j < count; j += kOsThreads) {
            const uint2 e = sm.out[j];
            const uint32_t dst = sm.delta[digit_of<BYTE>(e.x, a.shift, a.mask)] + j;
            if (a.final_pass) {
                a.out_vals[dst] = e.y;
                if (a.out_keys != nullptr) a.out_keys[dst] = e.x;
            } else {
                a.out_pairs[dst] = e;
            }
        }
    }
}

template <bool TEXT, int BITS, bool BYTE>
__global__ void __launch_bounds__(kOsThreads, 1) onesweep_kernel(const OsArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    OsSmem &sm = *reinterpret_cast<OsSmem *>(smem_raw);
    if (threadIdx.x == 0) {
        mbar_init(&sm.mbar, 1);
        const uint32_t first = atomicAdd(a.tile_counter, 1u);
        sm.next_tile = first;
        if (first < a.n_tiles) os_issue_load<TEXT>(a, sm, first);
    }
    __syncthreads();
    uint32_t parity = 0;
    uint32_t tile = sm.next_tile;
    while (tile < a.n_tiles) {
        const uint32_t count = (uint32_t)min((uint64_t)kOsTile, a.n - (uint64_t)tile * kOsTile);
        mbar_wait(&sm.mbar, parity);
        parity ^= 1;
        if (count == (uint32_t)kOsTile)
            os_tile<TEXT, BITS, BYTE, true>(a, sm, tile, count);
        else
            os_tile<TEXT, BITS, BYTE, false>(a, sm, tile, count);
        tile = sm.next_tile;  // written before the ranking's barriers of this iteration; rewritten only after the next
                              // iteration's first barrier
    }
}

// ------------------------------------------------------------------------------------------------
// digit histograms of all passes, before the first scatter
// ------------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;

// G[v] = number of text positions j in [0, n - c] whose c symbols form v (power-of-two alphabets: c * bits <= 8)
__global__ void __launch_bounds__(kHistThreads) cmer_hist_kernel(PackedText text, uint32_t c, uint32_t *__restrict__ G) {
    __shared__ uint32_t hist[kHistThreads / 32][kRadix];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (kHistThreads / 32) * kRadix; i += kHistThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t bits = text.bits, spw = 64 / bits;
    const uint64_t n_pos = text.n >= c ? text.n - c + 1 : 0;  // positions j that start a c-mer
    const uint64_t n_words = (n_pos + spw - 1) / spw;
    uint32_t *my = hist[warp];
    // a warp takes 32 consecutive words per round; lane l walks the spw positions of word base + l
    for (uint64_t wbase = ((uint64_t)blockIdx.x * (kHistThreads / 32) + warp) * 32; wbase < n_words;
         wbase += (uint64_t)gridDim.x * (kHistThreads / 32) * 32) {
        const uint64_t w = wbase + lane;
        uint64_t win = 0, next = 0;
        uint32_t valid = 0;
        if (w < n_words) {
            win = text.words[w];
            next = text.words[w + 1];
            const uint64_t left = n_pos - w * spw;
            valid = left < spw ? (uint32_t)left : spw;
        }
        for (uint32_t t = 0; t < spw; ++t) {
            const uint32_t v = (uint32_t)(win >> (64 - c * bits));
            const bool ok = t < valid;
            // low-entropy text: all lanes of the warp hit the same counter -- add once
            const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, v, 0);
            if (__all_sync(0xFFFFFFFFu, ok && v == v0)) {
                if (lane == 0) my[v0] += 32;
            } else if (ok) {
                atomicAdd(&my[v], 1u);
            }
            win = (win << bits) | (next >> (64 - bits));
            next <<= bits;
        }
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kHistThreads / 32; ++w) s += hist[w][tid];
        if (s) atomicAdd(&G[tid], s);
    }
}

// ghist[p][v] = number of k-mers whose digit p is v, for every pass, in one sweep (any alphabet, 32-bit hashes)
__global__ void __launch_bounds__(kHistThreads) digit_hist_generic_kernel(PackedText text, uint32_t k, uint64_t n_kmers,
                                                                          uint32_t n_passes, uint32_t w_bits,
                                                                          uint32_t *__restrict__ ghist) {
    extern __shared__ uint32_t hist_dyn[];  // [4 copies][n_passes][256]
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t per_copy = n_passes * kRadix;
    for (uint32_t i = tid; i < 4 * per_copy; i += kHistThreads) hist_dyn[i] = 0;
    __syncthreads();
    uint32_t *my = hist_dyn + (warp & 3) * per_copy;
    const uint32_t mask = (1u << w_bits) - 1;
    const uint32_t per_window = 64 / text.bits - k + 1;  // k-mers fully inside one window
    constexpr uint32_t kPer = 16;                        // consecutive positions per thread
    for (uint64_t i0 = ((uint64_t)blockIdx.x * kHistThreads + tid) * kPer; i0 < n_kmers;
         i0 += (uint64_t)gridDim.x * kHistThreads * kPer) {
        uint32_t j = 0;
        while (j < kPer && i0 + j < n_kmers) {
            uint64_t w = window64(text.words, i0 + j, text.bits);
            const uint32_t cnt = min(min(per_window, kPer - j), (uint32_t)min((uint64_t)kPer, n_kmers - i0 - j));
            for (uint32_t t = 0; t < cnt; ++t) {
                uint32_t key = (uint32_t)key_from_window(w, k, text.bits, text.sigma);
                for (uint32_t p = 0; p < n_passes; ++p) {
                    atomicAdd(&my[p * kRadix + (key & mask)], 1u);
                    key >>= w_bits;
                }
                w <<= text.bits;
            }
            j += cnt;
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < per_copy; i += kHistThreads) {
        const uint32_t s = hist_dyn[i] + hist_dyn[per_copy + i] + hist_dyn[2 * per_copy + i] + hist_dyn[3 * per_copy + i];
        if (s) atomicAdd(&ghist[i], s);
    }
}

// base[p][v] = number of k-mers whose digit p is < v. from_cmers: the histograms are derived from the text's c-mer
// histogram G (power-of-two alphabet, w_bits = c * bits): digit p < top of the k-mer at i is the c-mer at
// i + k - (p + 1) c, so its histogram is G without the c-mers that start before k - (p + 1) c or after n - (p + 1) c;
// the top digit (ct <= c symbols at offset 0) is G folded onto ct-symbol prefixes, without the starts after n - k.
__global__ void __launch_bounds__(kRadix) digit_bases_kernel(PackedText text, uint32_t k, uint32_t n_passes, uint32_t w_bits,
                                                             uint32_t key_bits, uint32_t from_cmers,
                                                             const uint32_t *__restrict__ hist_in, uint32_t *__restrict__ base) {
    __shared__ uint32_t h[8][kRadix];
    __shared__ uint32_t warp_sums[kRadix / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bits = text.bits;
    if (!from_cmers) {
        for (uint32_t p = 0; p < n_passes; ++p) h[p][tid] = hist_in[p * kRadix + tid];
    } else {
        const uint32_t c = w_bits / bits;
        const uint32_t ct = (key_bits - (n_passes - 1) * w_bits) / bits;
        for (uint32_t p = 0; p + 1 < n_passes; ++p) h[p][tid] = hist_in[tid];
        {
            const uint32_t fold = 1u << ((c - ct) * bits);  // c-mers per ct-symbol prefix
            uint32_t s = 0;
            if ((uint64_t)tid * fold < kRadix)
                for (uint32_t u = 0; u < fold && tid * fold + u < (uint32_t)kRadix; ++u) s += hist_in[tid * fold + u];
            h[n_passes - 1][tid] = s;
        }
        __syncthreads();
        const uint64_t n = text.n;
        // c-mers at the ends of the text that are no digit p of any k-mer: k - c of them per pass
        for (uint32_t p = 0; p + 1 < n_passes; ++p) {
            const uint32_t o = k - (p + 1) * c;  // symbol offset of digit p inside a k-mer
            if ((uint32_t)tid < k - c) {
                const uint64_t j = (uint32_t)tid < o ? (uint64_t)tid : n - k + o + 1 + ((uint32_t)tid - o);
                const uint32_t v = (uint32_t)(window64(text.words, j, bits) >> (64 - c * bits));
                atomicSub(&h[p][v], 1u);
            }
        }
        if ((uint32_t)tid < k - c) {  // top digit: starts j in (n - k, n - c] were counted by G but start no k-mer
            const uint64_t j = n - k + 1 + (uint32_t)tid;
            const uint32_t v = (uint32_t)(window64(text.words, j, bits) >> (64 - ct * bits));
            atomicSub(&h[n_passes - 1][v], 1u);
        }
    }
    __syncthreads();
    for (uint32_t p = 0; p < n_passes; ++p) {
        const uint32_t v = h[p][tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t b = incl - v;
        for (int w = 0; w < warp; ++w) b += warp_sums[w];
        base[p * kRadix + tid] = b;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// key-range parts (multi-GPU build): an index part holds the k-mers whose hash lies in [lo, hi). The text is
// filtered into (hash - lo, position) pairs in position order (count -> scan -> write), which the scatter passes
// then sort like any other pair array.
// ------------------------------------------------------------------------------------------------
constexpr int kFilterThreads = 512;
constexpr int kFilterItems = 16;
constexpr int kFilterTile = kFilterThreads * kFilterItems;  // 8192 text positions per CTA

// the thread's kFilterItems consecutive k-mers: bit j of the result = k-mer i0 + j is owned; keys[j] = hash - lo
__device__ __forceinline__ uint32_t owned_mask(const PackedText &text, uint32_t k, uint64_t n_kmers, uint64_t i0, uint64_t lo,
                                               uint64_t hi, uint32_t (&keys)[kFilterItems]) {
    uint32_t m = 0;
    const uint32_t per_window = 64 / text.bits - k + 1;
    uint32_t j = 0;
    while (j < (uint32_t)kFilterItems && i0 + j < n_kmers) {
        uint64_t w = window64(text.words, i0 + j, text.bits);
        const uint32_t cnt = min(min(per_window, (uint32_t)kFilterItems - j), (uint32_t)min((uint64_t)kFilterItems, n_kmers - i0 - j));
        for (uint32_t t = 0; t < cnt; ++t) {
            const uint64_t key = key_from_window(w, k, text.bits, text.sigma);
            if (key >= lo && key < hi) m |= 1u << (j + t);
            keys[j + t] = (uint32_t)(key - lo);
            w <<= text.bits;
        }
        j += cnt;
    }
    return m;
}

__global__ void __launch_bounds__(kFilterThreads) owned_count_kernel(PackedText text, uint32_t k, uint64_t n_kmers, uint64_t lo,
                                                                      uint64_t hi, uint64_t *__restrict__ tile_counts) {
    __shared__ uint32_t warp_sums[kFilterThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t keys[kFilterItems];
    const uint64_t i0 = (uint64_t)blockIdx.x * kFilterTile + (uint64_t)tid * kFilterItems;
    uint32_t c = i0 < n_kmers ? (uint32_t)__popc(owned_mask(text, k, n_kmers, i0, lo, hi, keys)) : 0u;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane == 0) warp_sums[warp] = c;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kFilterThreads / 32; ++w) s += warp_sums[w];
        tile_counts[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kFilterThreads) owned_write_kernel(PackedText text, uint32_t k, uint64_t n_kmers, uint64_t lo,
                                                                      uint64_t hi, const uint64_t *__restrict__ tile_offsets,
                                                                      uint32_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals) {
    // the tile's owned pairs are compacted in shared memory (position order) and leave as two coalesced streams
    extern __shared__ uint32_t s_stage[];  // keys[kFilterTile], vals[kFilterTile]
    uint32_t *s_keys = s_stage, *s_vals = s_stage + kFilterTile;
    __shared__ uint32_t warp_sums[kFilterThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t keys[kFilterItems];
    const uint64_t i0 = (uint64_t)blockIdx.x * kFilterTile + (uint64_t)tid * kFilterItems;
    const uint32_t m = i0 < n_kmers ? owned_mask(text, k, n_kmers, i0, lo, hi, keys) : 0u;
    const uint32_t c = (uint32_t)__popc(m);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = incl - c, total = 0;
#pragma unroll
    for (int w = 0; w < kFilterThreads / 32; ++w) {
        if (w < warp) before += warp_sums[w];
        total += warp_sums[w];
    }
#pragma unroll
    for (int j = 0; j < kFilterItems; ++j)
        if ((m >> j) & 1u) {
            s_keys[before] = keys[j];
            s_vals[before] = (uint32_t)(i0 + j);
            ++before;
        }
    __syncthreads();
    const uint64_t base = tile_offsets[blockIdx.x];
    for (uint32_t i = tid; i < total; i += kFilterThreads) {
        out_keys[base + i] = s_keys[i];
        out_vals[base + i] = s_vals[i];
    }
}

// ghist[p][v] over a pair array (all passes in one sweep)
__global__ void __launch_bounds__(kHistThreads) digit_hist_pairs_kernel(const uint2 *__restrict__ pairs, uint64_t n, uint32_t n_passes,
                                                                        uint32_t w_bits, uint32_t *__restrict__ ghist) {
    extern __shared__ uint32_t hist_dyn[];  // [4 copies][n_passes][256]
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t per_copy = n_passes * kRadix;
    for (uint32_t i = tid; i < 4 * per_copy; i += kHistThreads) hist_dyn[i] = 0;
    __syncthreads();
    uint32_t *my = hist_dyn + (warp & 3) * per_copy;
    const uint32_t mask = (1u << w_bits) - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * kHistThreads + tid; i < n; i += (uint64_t)gridDim.x * kHistThreads) {
        uint32_t key = pairs[i].x;
        for (uint32_t p = 0; p < n_passes; ++p) {
            atomicAdd(&my[p * kRadix + (key & mask)], 1u);
            key >>= w_bits;
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < per_copy; i += kHistThreads) {
        const uint32_t s = hist_dyn[i] + hist_dyn[per_copy + i] + hist_dyn[2 * per_copy + i] + hist_dyn[3 * per_copy + i];
        if (s) atomicAdd(&ghist[i], s);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int sm_count_cached() {
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev] > 0 ? sms[dev] : 148;
}

uint32_t device_sm_count() { return (uint32_t)sm_count_cached(); }

void launch_digit_histograms(const PackedText &text, uint32_t k, uint64_t n_kmers, uint32_t n_passes, uint32_t w_bits,
                             uint32_t key_bits, bool from_cmers, uint32_t *d_hist_scratch, uint32_t *d_digit_base,
                             cudaStream_t stream) {
    cudaMemsetAsync(d_hist_scratch, 0, 8 * kRadix * sizeof(uint32_t), stream);
    const int sms = sm_count_cached();
    if (from_cmers) {
        cmer_hist_kernel<<<sms * 4, kHistThreads, 0, stream>>>(text, w_bits / text.bits, d_hist_scratch);
    } else {
        const size_t smem = 4 * (size_t)n_passes * kRadix * sizeof(uint32_t);
        digit_hist_generic_kernel<<<sms * 4, kHistThreads, smem, stream>>>(text, k, n_kmers, n_passes, w_bits, d_hist_scratch);
    }
    digit_bases_kernel<<<1, kRadix, 0, stream>>>(text, k, n_passes, w_bits, key_bits, from_cmers ? 1u : 0u, d_hist_scratch,
                                                 d_digit_base);
}

uint32_t filter_tiles(uint64_t n_kmers) { return (uint32_t)((n_kmers + kFilterTile - 1) / kFilterTile); }

void launch_owned_count(const PackedText &text, uint32_t k, uint64_t n_kmers, uint64_t lo, uint64_t hi, uint64_t *d_tile_counts,
                        cudaStream_t stream) {
    owned_count_kernel<<<filter_tiles(n_kmers), kFilterThreads, 0, stream>>>(text, k, n_kmers, lo, hi, d_tile_counts);
}

void launch_owned_write(const PackedText &text, uint32_t k, uint64_t n_kmers, uint64_t lo, uint64_t hi, const uint64_t *d_tile_offsets,
                        uint32_t *d_out_keys, uint32_t *d_out_vals, cudaStream_t stream) {
    const size_t smem = 2 * kFilterTile * sizeof(uint32_t);
    cudaFuncSetAttribute(owned_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    owned_write_kernel<<<filter_tiles(n_kmers), kFilterThreads, smem, stream>>>(text, k, n_kmers, lo, hi, d_tile_offsets, d_out_keys,
                                                                               d_out_vals);
}

void launch_digit_histograms_pairs(const uint2 *d_pairs, uint64_t n, uint32_t n_passes, uint32_t w_bits, uint32_t *d_hist_scratch,
                                   uint32_t *d_digit_base, cudaStream_t stream) {
    cudaMemsetAsync(d_hist_scratch, 0, 8 * kRadix * sizeof(uint32_t), stream);
    const size_t smem = 4 * (size_t)n_passes * kRadix * sizeof(uint32_t);
    digit_hist_pairs_kernel<<<sm_count_cached() * 4, kHistThreads, smem, stream>>>(d_pairs, n, n_passes, w_bits, d_hist_scratch);
    digit_bases_kernel<<<1, kRadix, 0, stream>>>(PackedText{}, 0, n_passes, w_bits, 0, 0u, d_hist_scratch, d_digit_base);
}

template <bool TEXT, int BITS, bool BYTE>
static void launch_onesweep_t(const OsArgs &a, cudaStream_t stream) {
    cudaFuncSetAttribute(onesweep_kernel<TEXT, BITS, BYTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsSmem));
    const uint32_t grid = std::min<uint32_t>(a.n_tiles, (uint32_t)sm_count_cached());
    onesweep_kernel<TEXT, BITS, BYTE><<<grid, kOsThreads, sizeof(OsSmem), stream>>>(a);
}

template <bool TEXT>
static void launch_onesweep_bits(const OsArgs &a, cudaStream_t stream) {
    const int bits = __builtin_popcount(a.mask);
    if (bits <= 4)
        launch_onesweep_t<TEXT, 4, false>(a, stream);
    else if (bits <= 6)
        launch_onesweep_t<TEXT, 6, false>(a, stream);
    else if (bits == 8 && (a.shift & 7) == 0)
        launch_onesweep_t<TEXT, 8, true>(a, stream);
    else
        launch_onesweep_t<TEXT, 8, false>(a, stream);
}

void launch_onesweep_pass(const PackedText *text, uint32_t k, const uint2 *d_in_pairs, uint64_t n, uint32_t shift, uint32_t mask,
                          uint32_t tag, const uint32_t *d_digit_base, uint64_t *d_status, uint32_t *d_tile_counter,
                          uint2 *d_out_pairs, uint32_t *d_out_keys, uint32_t *d_out_vals, cudaStream_t stream) {
    OsArgs a{};
    a.in_pairs = d_in_pairs;
    if (text) a.text = *text;
    a.k = k;
    a.n = n;
    a.n_tiles = (uint32_t)((n + kOsTile - 1) / kOsTile);
    a.shift = shift;
    a.mask = mask;
    a.tag = tag;
    a.final_pass = d_out_pairs == nullptr ? 1u : 0u;
    a.debug_no_lookback = std::getenv("KMER_B200_DEBUG_NO_LOOKBACK") ? 1u : 0u;
    a.digit_base = d_digit_base;
    a.status = d_status;
    a.tile_counter = d_tile_counter;
    a.out_pairs = d_out_pairs;
    a.out_keys = d_out_keys;
    a.out_vals = d_out_vals;
    if (text)
        launch_onesweep_bits<true>(a, stream);
    else
        launch_onesweep_bits<false>(a, stream);
}

}  // namespace kb
