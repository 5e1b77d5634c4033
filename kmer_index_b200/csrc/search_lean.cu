// Lean count pass (sm_100a) for the common shape of the search: ONE k (kmer_index<alphabet, k>), 2-bit symbols (dna4),
// dense directory, queries of at most 32 * W symbols. Same results as search_query<kPassCount> (search_kernels.cu), which
// remains the general path (multi-k plans, wide alphabets, long queries, sharded / accounting variants, the write pass).
//
// Why a second kernel: ncu on the general count pass (profiles/README.md) showed it front-end bound, not gather bound --
// 1 430 warp instructions per 32 queries, 212 bytes of spill stores per lane under its 32-register cap, 1.9x the
// algorithmic DRAM traffic. Most of that is generality: the packed query staged in shared memory and re-read through
// dynamic indexing, the scheme-table plan, group collectives, accounting hooks. Here the query lives in W registers, a
// part's hash is two funnel shifts, the plan is arithmetic on (m, k), and the whole-text rules of the reference
// (kmer_index.hpp:216-227, :234 -> :119, :314) are evaluated straight from the directory.
// Queries the lean path does not cover (m < k: prefix slabs; candidate lists beyond kHeavyCandidates) are appended to the
// batch's "heavy" list and answered by the general kernel's warp-per-query launch, exactly as long buckets already are.
#include <cstdlib>

#include "launch.h"
#include "query_pack.cuh"

namespace kb {

constexpr int kLeanThreads = 256;
constexpr uint32_t kLeanHeavyCandidates = 2048;  // == kHeavyCandidates of the general kernel

// the W words as one left-aligned bit string: 64 bits starting at bit `bit` (bit < 64 * W; bits past the end are zero)
template <int W>
__device__ __forceinline__ uint64_t window_at(const uint64_t (&w)[W], uint32_t bit) {
    const uint32_t word = bit >> 6, sh = bit & 63;
    uint64_t hi = 0, lo = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if ((uint32_t)i == word) hi = w[i];
        if ((uint32_t)i == word + 1) lo = w[i];
    }
    return sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
}

// pack_query_lane<2> with the words in registers (every index a compile-time constant after unrolling)
template <int W>
__device__ __forceinline__ bool pack_query_regs(const uint8_t *__restrict__ q_ranks, uint64_t off0, uint32_t m, uint64_t q_total,
                                                uint32_t sigma, uint64_t (&qw)[W]) {
    const uint8_t *qr = q_ranks + off0;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(qr) & 7);
    const uint32_t sh = (mis & 3) * 8;
    const bool odd = mis >= 4;
    const uint64_t guard = (0x80u - sigma) * 0x0101010101010101ull;
    uint64_t inval = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t s0 = w * 32;
        uint64_t acc = 0;
        if (s0 < m) {
            const uint32_t n_sym = min(32u, m - s0);
            uint64_t v[4];
            if (off0 + s0 >= 8 && off0 + s0 + 32 + 8 <= q_total) {
                const uint64_t *ap = reinterpret_cast<const uint64_t *>(qr + s0 - mis);
                uint32_t r[10];
#pragma unroll
                for (int i = 0; i <= 4; ++i) {
                    const uint64_t x = ap[i];
                    r[2 * i] = (uint32_t)x;
                    r[2 * i + 1] = (uint32_t)(x >> 32);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t A = odd ? r[2 * c + 1] : r[2 * c];
                    const uint32_t B = odd ? r[2 * c + 2] : r[2 * c + 1];
                    const uint32_t C = odd ? r[2 * c + 3] : r[2 * c + 2];
                    v[c] = ((uint64_t)__funnelshift_r(B, C, sh) << 32) | __funnelshift_r(A, B, sh);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = (uint32_t)(8 * c) < n_sym ? load8(qr + s0 + 8 * c, n_sym - 8 * c, false) : 0ull;
            }
            if (n_sym < 32) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t nv = n_sym > (uint32_t)(8 * c) ? n_sym - 8 * c : 0u;
                    if (nv < 8) v[c] &= (1ull << (8 * nv)) - 1;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                inval |= (v[c] + guard) | v[c];
                acc |= pack8(v[c], 2) << (64 - 16 * (c + 1));
            }
        }
        qw[w] = acc;
    }
    return (inval & 0x8080808080808080ull) != 0;
}

// T[tpos, tpos + len) == q[qpos, qpos + len) with the query in registers; false if the text span leaves the text
template <int W>
__device__ __forceinline__ bool match_span_regs(const PackedText &T, const uint64_t (&qw)[W], uint64_t tpos, uint32_t qpos, uint32_t len) {
    if (tpos + len > T.n) return false;
    while (len) {
        const uint32_t c = len < 32 ? len : 32;
        const uint64_t a = text_window64(T.words, tpos, 2);
        const uint64_t b = window_at<W>(qw, qpos * 2);
        if ((a ^ b) >> (64 - 2 * c)) return false;
        tpos += c;
        qpos += c;
        len -= c;
    }
    return true;
}

template <int W, int MIN_BLOCKS>
__global__ void __launch_bounds__(kLeanThreads, MIN_BLOCKS) search_count_lean_kernel(const SearchArgs a) {
    const uint64_t q = (uint64_t)blockIdx.x * kLeanThreads + threadIdx.x;
    if (q >= a.n_queries) return;
    const int lane = threadIdx.x & 31;
    const DeviceIndex &ix = *a.index;
    const PackedText T = ix.text;
    const Element &E = ix.elem[0];
    const uint32_t k = E.k;

    const bool packed_in = a.q_packed != nullptr;
    const uint64_t off0 = packed_in ? 0 : a.q_offsets[q];
    const uint64_t m64 = packed_in ? (uint64_t)a.q_lens16[q] : a.q_offsets[q + 1] - off0;
    uint32_t status = KMER_B200_QUERY_OK;
    if (m64 == 0) {
        status = KMER_B200_QUERY_UNDEFINED;  // assert(query.size() > 0), kmer_index.hpp:195
    } else if (a.mode == KMER_B200_MODE_REFERENCE_EXACT && m64 > kQuerySizeRange) {
        status = KMER_B200_QUERY_THROW_INVALID_ARGUMENT;  // kmer_index.hpp:507-509
    } else if (a.mode == KMER_B200_MODE_REFERENCE_EXACT && m64 == kQuerySizeRange) {
        status = KMER_B200_QUERY_UNDEFINED;  // kmer_index.hpp:512
    } else if (m64 > a.max_len) {
        status = KMER_B200_QUERY_TOO_LONG_FOR_SHARD;
    }
    auto finish = [&](uint64_t n_hits, uint32_t st, uint32_t flags) {
        a.counts[q] = n_hits;
        a.status[q] = (uint8_t)st;
        a.unsorted[q] = (uint8_t)flags;
    };
    if (status != KMER_B200_QUERY_OK) {
        finish(0, status, 0);
        return;
    }
    const uint32_t m = (uint32_t)m64;
    if (m < k) {
        // prefix slab (kmer_index.hpp:342-345): the general kernel's warp-per-query launch answers it
        a.heavy[1 + atomicAdd(a.heavy, 1u)] = (uint32_t)q;
        finish(0, KMER_B200_QUERY_OK, 2);
        return;
    }

    uint64_t qw[W];
    if (packed_in) {
        const uint64_t *src = a.q_packed + q * a.q_stride;
#pragma unroll
        for (int w = 0; w < W; ++w) qw[w] = (uint32_t)w < a.q_stride ? src[w] : 0ull;
    } else {
        if (pack_query_regs<W>(a.q_ranks, off0, m, a.q_offsets[a.n_queries], T.sigma, qw)) atomicOr(a.error_flag, 1u);
    }

    // ---- plan (kmer_index_element::search, kmer_index.hpp:193-346; CORRECT mode: always the contiguous comparison)
    const uint32_t P = m / k, rest = m - P * k;
    const bool exact = m == k;
    bool throw_after = false, buggy = false;
    if (a.mode == KMER_B200_MODE_REFERENCE_EXACT && !exact) {
        throw_after = rest > 0 && ix.pow_sigma[k - rest] > 10000000ull;  // :234 -> :119
        buggy = rest != 0 && P > 2;                                       // :314
    }
    const uint32_t down = 64 - 2 * k;
    Range seed = bucket_of(E, window_at<W>(qw, 0) >> down);
    uint32_t seed_d = 0;
    bool all_present = seed.cnt != 0;
    if ((throw_after || buggy) && all_present) {
        // the reference looks every full part up first (:216-227); only these plans' outcome depends on it
        Range last{0, 0};
        bool last_foreign = false;
        const uint64_t *pb = ix.presence[0];
        for (uint32_t j = 1; j < P; ++j) {
            const uint64_t key = window_at<W>(qw, j * k * 2) >> down;
            last_foreign = pb != nullptr && (key < E.key_lo || key >= E.key_hi);
            if (last_foreign) {
                // another part's hash (key-range multi-GPU search): presence from the replicated bitmap; it cannot seed
                last = Range{0, (gather64(pb + (key >> 6)) >> (key & 63)) & 1ull};
            } else {
                last = bucket_of(E, key);
            }
            if (last.cnt == 0) {
                all_present = false;
                break;
            }
        }
        if (all_present && buggy && !last_foreign && last.cnt < seed.cnt) {  // seed from the shorter of the two constrained parts
            seed = last;
            seed_d = (P - 1) * k;
        }
    }
    if (!all_present) {  // :224 return result_t()
        finish(0, KMER_B200_QUERY_OK, 0);
        return;
    }
    if (throw_after) {
        finish(0, KMER_B200_QUERY_THROW_INVALID_ARGUMENT, 0);
        return;
    }
    if (seed.cnt > (uint64_t)kLeanHeavyCandidates) {
        a.heavy[1 + atomicAdd(a.heavy, 1u)] = (uint32_t)q;
        finish(0, KMER_B200_QUERY_OK, 2);
        return;
    }

    // ---- candidates: verified against the text
    uint64_t n_hits = 0;
    if (exact && ix.owned == T.n) {
        n_hits = seed.cnt;
    } else {
        const uint32_t last_q = (P - 1) * k;
        // a bucket lies in one part of a peer-positions index; the part table sits in the launch parameters
        const uint32_t *cand = E.pos + seed.lo;
        if (a.parts0.n != 0 && seed.cnt != 0) {
            uint32_t r = 0;
#pragma unroll
            for (uint32_t j = 1; j < (uint32_t)kMaxPosParts; ++j) r += (j < a.parts0.n && seed.lo >= (uint64_t)a.parts0.first[j]) ? 1u : 0u;
            cand = a.parts0.ptr[r] + (seed.lo - a.parts0.first[r]);
        }
        for (uint64_t c = 0; c < seed.cnt; ++c) {
            const uint32_t at = gather32(cand + c);
            if (at < seed_d) continue;
            const uint64_t p = at - seed_d;
            if (p >= ix.owned) continue;
            bool ok;
            if (exact) {
                ok = true;
            } else if (!buggy) {
                ok = match_span_regs<W>(T, qw, p, 0, m);  // the whole query: the seed part compares equal by construction
            } else {
                // part 0 in place, the middle parts compared against the LAST part (:314), the last part + rest in place
                ok = match_span_regs<W>(T, qw, p, 0, k);
                for (uint32_t j = 1; ok && j + 1 < P; ++j) ok = match_span_regs<W>(T, qw, p + j * k, last_q, k);
                ok = ok && match_span_regs<W>(T, qw, p + last_q, last_q, k + rest);
            }
            n_hits += ok ? 1 : 0;
        }
    }
    finish(n_hits, KMER_B200_QUERY_OK, 0);
    if (a.hits != nullptr && n_hits > 0) {
        // the write pass only visits the queries listed here (one atomic per set of lanes arriving together)
        const uint32_t act = __activemask();
        const int leader = __ffs(act) - 1;
        uint32_t slot = 0;
        if (lane == leader) slot = atomicAdd(a.hits, (uint32_t)__popc(act));
        slot = __shfl_sync(act, slot, leader) + __popc(act & ((1u << lane) - 1));
        a.hits[1 + slot] = (uint32_t)q;
    }
}

// true if the lean kernel took the launch (the caller still launches the heavy pass of the general kernel afterwards)
bool launch_search_count_lean(const SearchArgs &a, cudaStream_t stream) {
    if (!a.lean_ok || a.heavy == nullptr || a.n_queries == 0 || a.max_len > 128) return false;
    const unsigned blocks = (unsigned)((a.n_queries + kLeanThreads - 1) / kLeanThreads);
    // 8 CTAs per SM (32 registers, 84 bytes of spills that stay in L1) against 6 (40 registers, none): the pass is bound by
    // the latency of its dependent gathers, so the extra resident warps win -- config 5: 5.50 ms against 5.67.
    // KMER_B200_LEAN_BLOCKS=6 selects the other variant.
    static const int dense = [] {
        const char *e = std::getenv("KMER_B200_LEAN_BLOCKS");
        return e ? std::atoi(e) : 8;
    }();
    if (a.max_len <= 64) {
        if (dense == 8)
            search_count_lean_kernel<2, 8><<<blocks, kLeanThreads, 0, stream>>>(a);
        else
            search_count_lean_kernel<2, 6><<<blocks, kLeanThreads, 0, stream>>>(a);
    } else {
        search_count_lean_kernel<4, 5><<<blocks, kLeanThreads, 0, stream>>>(a);
    }
    return true;
}

}  // namespace kb
