// Stable LSD radix-sort building blocks (sm_100a). Used by
//   * the index build: (hash, position) pairs, first pass fused with the rolling hash over the packed
//     text (keys are never materialised unsorted), and
//   * the per-query segment sort of sub-k results.
// Stability is what makes every bucket's positions ascend, which the reference gets from push_back in
// text order (kmer_index.hpp:165) and relies on for std::binary_search / lower_bound (:242,283,315).
#pragma once

#include "common.cuh"

namespace kb {

constexpr int kRadixBitsMax = 8;
constexpr int kRadix = 1 << kRadixBitsMax;

// Tile geometry: one CTA ranks kSortTile elements; element e of a tile belongs to
// (warp, round, lane) = (e / (32*ITEMS), (e / 32) % ITEMS, e % 32), so rank order == element order.
constexpr int kSortThreads = 512;
constexpr int kSortItems = 16;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortTile = kSortThreads * kSortItems;  // 8192

constexpr uint32_t kInvalidDigit = 0xFFFFFFFFu;

struct RankSmem {
    uint32_t warp_cnt[kSortWarps][kRadix];  // per-warp digit counters, then exclusive warp prefixes
    uint32_t excl[kRadix];                  // exclusive prefix of the tile's digit counts
    uint32_t count[kRadix];                 // the tile's digit counts
    uint32_t warp_sums[kRadix / 32];
};

// Lanes of the warp holding the same digit as this lane, restricted to `peers` on entry. One vote per digit
// bit: the hardware match.any instruction is several times slower than 8 votes on sm_100a, and the ALU pipe
// (one warp instruction per two cycles per SM sub-partition) is what bounds the scatter passes, so every bit
// costs exactly three ALU instructions here: test, conditional complement, and.
template <int BITS>
__device__ __forceinline__ uint32_t match_digit(uint32_t d, uint32_t peers) {
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " .reg .b32 m;\n"
            " and.b32 m, %1, %2;\n"
            " setp.ne.u32 p, m, 0;\n"
            " vote.sync.ballot.b32 m, p, 0xffffffff;\n"
            " @!p not.b32 m, m;\n"
            " and.b32 %0, %0, m;\n"
            "}\n"
            : "+r"(peers)
            : "r"(d), "r"(1u << b));
    }
    return peers;
}

// The digit of a key. BYTE: a byte-aligned 8-bit digit of a 32-bit key (shift in {0, 8, 16, 24}, mask == 0xFF) is one
// PRMT instead of a shift and a mask.
template <bool BYTE, typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, uint32_t shift, uint32_t mask) {
    if constexpr (BYTE) {
        static_assert(sizeof(KeyT) == 4, "byte digits are taken from 32-bit keys");
        return __byte_perm((uint32_t)key, 0u, 0x4440u | (shift >> 3));
    } else {
        return (uint32_t)(key >> shift) & mask;
    }
}

// Stable rank of this thread's kSortItems keys inside the tile by digit (key >> shift) & mask (BITS >= the
// digit width). Item r of this thread is tile element ((warp * kSortItems + r) * 32 + lane); with FULL = false
// elements >= count are ignored. On return
//   local_pos2[r / 2] holds, in its 16-bit half r % 2, the position of item r in the tile's digit-sorted order
//                     (tile positions are < 2^13; undefined for ignored items) -- two per register keeps the
//                     kernel inside its 64-register budget without spilling,
//   sm.count[d]  = number of items with digit d, sm.excl[d] = exclusive prefix of count.
// All kSortThreads threads must call. Ends with a __syncthreads().
template <int BITS, bool FULL, typename KeyT, bool BYTE = false, uint32_t ATOMIC_ITEMS = 0>
__device__ __forceinline__ void tile_rank(const KeyT (&key)[kSortItems], uint32_t count, uint32_t shift,
                                          uint32_t mask, uint32_t (&local_pos2)[kSortItems / 2], RankSmem &sm,
                                          uint32_t (*warp_msk)[kRadix] = nullptr) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t e0 = (uint32_t)(warp * kSortItems * 32 + lane);
    uint32_t *my_cnt = sm.warp_cnt[warp];

  // ATOMIC_ITEMS: bit r set = item r is matched through shared memory (below), clear = by votes (match_digit)
  constexpr bool ATOMIC_MATCH = (ATOMIC_ITEMS & 0xFFFFu) == 0xFFFFu;
  constexpr bool HYBRID_MATCH = ATOMIC_ITEMS != 0 && !ATOMIC_MATCH;
  if constexpr (ATOMIC_MATCH) {
    // The scatter passes are bound by instruction issue (ALU pipe), not by memory: the vote-based match costs 3 ALU
    // instructions per digit bit and item. Here the match runs on the shared-memory pipe instead: every lane ORs its
    // lane bit into the warp's mask word of its digit, reads the mask (= the same-digit lanes) and the digit's running
    // counter back, and the lowest lane of each group adds the group's size and clears the mask. Measured on config 5:
    // the text-sourced pass (which also spends ALU on extracting hashes) 16.7 -> 15.4 ms, the pair passes 13.3 -> 15.7 ms
    // (they become shared-memory bound) -- so only the text-sourced pass uses it.
    uint32_t *my_msk = warp_msk[warp];  // ATOMIC_MATCH: per warp and digit, the lanes holding that digit in this round
#pragma unroll
    for (int i = 0; i < kRadix / 32; ++i) {
        my_cnt[i * 32 + lane] = 0;
        my_msk[i * 32 + lane] = 0;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        const bool valid = FULL || e0 + r * 32 < count;
        if (valid) atomicOr(&my_msk[d], 1u << lane);
        __syncwarp();
        uint32_t peers = 0, pre = 0;
        if (valid) {
            peers = my_msk[d];
            pre = my_cnt[d];
        }
        __syncwarp();
        const uint32_t below = (uint32_t)__popc(peers & lt_mask);
        if (valid && below == 0) {
            my_cnt[d] = pre + (uint32_t)__popc(peers);
            my_msk[d] = 0;
        }
        __syncwarp();
        const uint32_t pos = pre + below;
        if (r & 1)
            local_pos2[r >> 1] |= pos << 16;
        else
            local_pos2[r >> 1] = pos;
    }
  } else if constexpr (HYBRID_MATCH) {
    // some items by votes (ALU pipe), the others through shared memory (LSU pipe): both pipes work
    uint32_t *my_msk = warp_msk[warp];
#pragma unroll
    for (int i = 0; i < kRadix / 32; ++i) {
        my_cnt[i * 32 + lane] = 0;
        my_msk[i * 32 + lane] = 0;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        if ((ATOMIC_ITEMS >> r) & 1u) continue;
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        const bool valid = FULL || e0 + r * 32 < count;
        const uint32_t peers = match_digit<BITS>(d, FULL ? 0xFFFFFFFFu : __ballot_sync(0xFFFFFFFFu, valid));
        const uint32_t info = (uint32_t)__popc(peers & lt_mask) | ((uint32_t)__popc(peers) << 5) | ((valid ? 1u : 0u) << 11);
        if (r & 1)
            local_pos2[r >> 1] = (local_pos2[r >> 1] & 0xFFFFu) | (info << 16);
        else
            local_pos2[r >> 1] = (local_pos2[r >> 1] & 0xFFFF0000u) | info;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        uint32_t pos;
        if ((ATOMIC_ITEMS >> r) & 1u) {
            const bool valid = FULL || e0 + r * 32 < count;
            if (valid) atomicOr(&my_msk[d], 1u << lane);
            __syncwarp();
            uint32_t peers = 0, pre = 0;
            if (valid) {
                peers = my_msk[d];
                pre = my_cnt[d];
            }
            __syncwarp();
            const uint32_t below = (uint32_t)__popc(peers & lt_mask);
            if (valid && below == 0) {
                my_cnt[d] = pre + (uint32_t)__popc(peers);
                my_msk[d] = 0;
            }
            __syncwarp();
            pos = pre + below;
        } else {
            const uint32_t info = (local_pos2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu;
            const bool valid = FULL || ((info >> 11) & 1u);
            const uint32_t below = info & 31u;
            uint32_t pre = 0;
            if (valid) pre = my_cnt[d];
            __syncwarp();
            if (valid && below == 0) my_cnt[d] = pre + ((info >> 5) & 63u);
            __syncwarp();
            pos = pre + below;
        }
        if (r & 1)
            local_pos2[r >> 1] = (local_pos2[r >> 1] & 0xFFFFu) | (pos << 16);
        else
            local_pos2[r >> 1] = (local_pos2[r >> 1] & 0xFFFF0000u) | pos;
    }
  } else {
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&sm.warp_cnt[0][0])[i] = 0;
    __syncthreads();

    // phase A: the same-digit lane masks of all items (independent vote chains: the scheduler overlaps them);
    // kept packed per item as lanes-below count (5 bits) | group size << 5 (6 bits) | valid << 11
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        const bool valid = FULL || e0 + r * 32 < count;
        const uint32_t peers = match_digit<BITS>(d, FULL ? 0xFFFFFFFFu : __ballot_sync(0xFFFFFFFFu, valid));
        const uint32_t info = (uint32_t)__popc(peers & lt_mask) | ((uint32_t)__popc(peers) << 5) | ((valid ? 1u : 0u) << 11);
        if (r & 1)
            local_pos2[r >> 1] |= info << 16;
        else
            local_pos2[r >> 1] = info;
    }
    // phase B: per-warp counters, in item order (the lowest lane of each group bumps the counter of its digit)
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
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
  }
    __syncthreads();

    // per digit: exclusive prefix over warps, total count, exclusive prefix over digits
    if (tid < kRadix) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
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
        sm.excl[tid] = incl - run;  // warp-local exclusive for now
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t base = sm.excl[tid];
        for (int w = 0; w < warp; ++w) base += sm.warp_sums[w];
        sm.excl[tid] = base;
        // fold the digit base into the warp prefixes: one lookup per item below instead of two
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) sm.warp_cnt[w][tid] += base;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t d = digit_of<BYTE>(key[r], shift, mask);
        // no carry between the halves: every final position is < kSortTile <= 2^16
        if (FULL || e0 + r * 32 < count) local_pos2[r >> 1] += my_cnt[d] << (16 * (r & 1));
    }
    __syncthreads();
}

}  // namespace kb
