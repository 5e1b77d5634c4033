// See host_pack.h. Plain C++ (no CUDA): compiled by the host compiler through nvcc.
#include "host_pack.h"

#include <immintrin.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace kb {

struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<void(unsigned)> fn;
    unsigned n_tasks = 0, next = 0, running = 0;
    uint64_t generation = 0;
    bool stop = false;

    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lock(mu);
        for (;;) {
            cv_work.wait(lock, [&] { return stop || (generation != seen && next < n_tasks); });
            if (stop) return;
            while (next < n_tasks) {
                const unsigned i = next++;
                ++running;
                lock.unlock();
                fn(i);
                lock.lock();
                --running;
            }
            seen = generation;
            if (running == 0) cv_done.notify_all();
        }
    }
};

HostPool::HostPool() : impl_(new Impl) {
    unsigned n = std::thread::hardware_concurrency();
    n = std::max(1u, std::min(n, 32u));
    for (unsigned t = 0; t < n; ++t) impl_->workers.emplace_back([this] { impl_->loop(); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lock(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv_work.notify_all();
    for (auto &w : impl_->workers) w.join();
    delete impl_;
}

HostPool &HostPool::instance() {
    static HostPool pool;
    return pool;
}

unsigned HostPool::threads() const { return (unsigned)impl_->workers.size(); }

void HostPool::submit(unsigned n_tasks, std::function<void(unsigned)> fn) {
    std::unique_lock<std::mutex> lock(impl_->mu);
    impl_->cv_done.wait(lock, [&] { return impl_->next >= impl_->n_tasks && impl_->running == 0; });
    impl_->fn = std::move(fn);
    impl_->n_tasks = n_tasks;
    impl_->next = 0;
    ++impl_->generation;
    lock.unlock();
    impl_->cv_work.notify_all();
}

void HostPool::wait() {
    std::unique_lock<std::mutex> lock(impl_->mu);
    impl_->cv_done.wait(lock, [&] { return impl_->next >= impl_->n_tasks && impl_->running == 0; });
}

uint64_t query_lengths_host(const uint64_t *q_offsets, uint64_t q_begin, uint64_t q_end, uint16_t *lens, unsigned part,
                            unsigned n_parts) {
    const uint64_t n = q_end - q_begin;
    const uint64_t lo = q_begin + n * part / n_parts, hi = q_begin + n * (part + 1) / n_parts;
    uint64_t mx = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t m = q_offsets[i + 1] - q_offsets[i];
        lens[i - q_begin] = (uint16_t)m;
        mx = std::max(mx, m);
    }
    return mx;
}

namespace {

// 8 ranks (byte j = symbol j) -> 8 * bits bits, symbol 0 on top. This file is compiled for x86-64-v3 (BMI2: one PEXT
// per 8 symbols); the shift-and-mask form is the portable fallback.
template <int BITS>
inline uint64_t pack8(uint64_t v) {
#if defined(__BMI2__)
    constexpr uint64_t mask = BITS == 2 ? 0x0303030303030303ull : BITS == 4 ? 0x0F0F0F0F0F0F0F0Full : ~0ull;
    return _pext_u64(__builtin_bswap64(v), mask);
#else
    uint64_t r = 0;
    for (int j = 0; j < 8; ++j) r |= ((v >> (8 * j)) & ((1ull << BITS) - 1)) << (BITS * (7 - j));
    return r;
#endif
}

#if defined(__AVX2__)
alignas(32) static const uint8_t kTailMask[64] = {
    0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
    0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};  // then 32 zeros

// the first min(left, 32) bytes of a 32-byte load kept, the rest zeroed
static inline __m256i keep_first(__m256i v, uint32_t left) {
    const uint32_t keep = left < 32 ? left : 32;
    return _mm256_and_si256(v, _mm256_loadu_si256(reinterpret_cast<const __m256i *>(kTailMask + 32 - keep)));
}

// 32 two-bit ranks (one per byte) -> 64 bits, symbol 0 in the top two bits: two multiply-adds merge neighbours
// (4 a + b, then 16 x + y), a byte shuffle collects the eight finished bytes
static inline uint64_t pack32x2(__m256i v) {
    const __m256i x = _mm256_maddubs_epi16(v, _mm256_set1_epi16(0x0104));
    const __m256i y = _mm256_madd_epi16(x, _mm256_set1_epi32(0x00010010));
    const __m256i sh = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1, -1, -1, -1,
                                        -1, -1, -1, -1, -1, -1, -1);
    const __m256i z = _mm256_shuffle_epi8(y, sh);
    return ((uint64_t)(uint32_t)_mm256_extract_epi32(z, 0) << 32) | (uint32_t)_mm256_extract_epi32(z, 4);
}

// 32 four-bit ranks -> two 64-bit words (16 symbols each)
static inline void pack32x4(__m256i v, uint64_t &w0, uint64_t &w1) {
    const __m256i x = _mm256_maddubs_epi16(v, _mm256_set1_epi16(0x0110));  // 16 a + b in every 16-bit lane
    const __m256i sh = _mm256_setr_epi8(14, 12, 10, 8, 6, 4, 2, 0, -1, -1, -1, -1, -1, -1, -1, -1, 14, 12, 10, 8, 6, 4, 2, 0, -1, -1,
                                        -1, -1, -1, -1, -1, -1);
    const __m256i z = _mm256_shuffle_epi8(x, sh);
    w0 = (uint64_t)_mm256_extract_epi64(z, 0);
    w1 = (uint64_t)_mm256_extract_epi64(z, 2);
}
#endif

// `safe_end`: reading at any address below it is inside the caller's buffer (a query's last chunk is fetched with one
// unconditional wide load and masked, instead of a variable-length copy)
template <int BITS>
bool pack_range(const uint8_t *q_ranks, const uint64_t *q_offsets, uint64_t q_begin, uint64_t lo, uint64_t hi, uint32_t sigma,
                uint32_t stride, uint64_t *words, const uint8_t *safe_end) {
    constexpr uint32_t SPW = 64 / BITS;  // symbols per word
    constexpr uint32_t CPW = SPW / 8;    // 8-symbol chunks per word
    const uint64_t guard = (uint64_t)(0x80u - (sigma > 128 ? 128u : sigma)) * 0x0101010101010101ull;
    uint64_t inval = 0;  // bit 7 of a byte ends up set <=> that rank is >= sigma (sigma <= 128)
    bool bad = false;
#if defined(__AVX2__)
    __m256i vmax = _mm256_setzero_si256();  // largest rank seen (BITS 2 / 4 paths)
#endif
    uint64_t off = q_offsets[lo];
    for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t next = q_offsets[i + 1];
        const uint8_t *src = q_ranks + off;
        const uint32_t m = (uint32_t)(next - off);
        off = next;
        uint64_t *dst = words + (i - q_begin) * stride;
        const uint32_t n_words = (m + SPW - 1) / SPW;
        uint32_t w = 0;
#if defined(__AVX2__)
        if ((BITS == 2 || BITS == 4) && src + (size_t)n_words * SPW + 32 <= safe_end) {
            if (BITS == 2) {
                for (; w < n_words; ++w) {
                    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32 * w));
                    v = keep_first(v, m - 32 * w);
                    vmax = _mm256_max_epu8(vmax, v);
                    dst[w] = pack32x2(v);
                }
            } else {
                for (uint32_t s = 0; s < m; s += 32) {
                    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + s));
                    v = keep_first(v, m - s);
                    vmax = _mm256_max_epu8(vmax, v);
                    uint64_t a, b;
                    pack32x4(v, a, b);
                    dst[w++] = a;
                    if (w < n_words) dst[w++] = b;
                }
            }
        } else
#endif
        if (src + (size_t)n_words * SPW + 8 <= safe_end) {
            // branch-free per word: always CPW 8-byte loads (those behind the query's end are masked to zero; they stay
            // inside the batch buffer), one PEXT each
            for (; w < n_words; ++w) {
                const uint8_t *p = src + w * SPW;
                const int32_t left = (int32_t)(m - w * SPW);  // symbols of the query from this word on, >= 1
                uint64_t acc = 0;
#pragma GCC unroll 8
                for (uint32_t c = 0; c < CPW; ++c) {
                    uint64_t v;
                    std::memcpy(&v, p + 8 * c, 8);
                    const int32_t l = left - (int32_t)(8 * c);
                    const uint64_t keep = l >= 8 ? ~0ull : (l <= 0 ? 0ull : ((1ull << (8 * l)) - 1));
                    v &= keep;
                    inval |= (v + guard) | v;
                    acc |= pack8<BITS>(v) << (64 - 8 * BITS * (c + 1));
                }
                dst[w] = acc;
            }
        } else {  // the last few queries of the batch: byte by byte
            for (; w < n_words; ++w) {
                uint64_t acc = 0;
                for (uint32_t c = 0; c < CPW; ++c) {
                    uint64_t v = 0;
                    for (uint32_t j = 0; j < 8; ++j) {
                        const uint32_t sidx = w * SPW + 8 * c + j;
                        if (sidx < m) v |= (uint64_t)src[sidx] << (8 * j);
                    }
                    inval |= (v + guard) | v;
                    acc |= pack8<BITS>(v) << (64 - 8 * BITS * (c + 1));
                }
                dst[w] = acc;
            }
        }
        for (; w < stride; ++w) dst[w] = 0;
        if (sigma > 128)
            for (uint32_t j = 0; j < m; ++j) bad |= src[j] >= sigma;
    }
    if (sigma <= 128) bad = (inval & 0x8080808080808080ull) != 0;
#if defined(__AVX2__)
    {
        alignas(32) uint8_t mx[32];
        _mm256_store_si256(reinterpret_cast<__m256i *>(mx), vmax);
        for (int j = 0; j < 32; ++j) bad |= mx[j] >= sigma;
    }
#endif
    return !bad;
}

// ---- streaming pack: a contiguous run of ranks -> words, query boundaries ignored ---------------------------------
#if defined(__AVX2__)
// 128 two-bit ranks -> four words: per 32-byte vector two multiply-adds leave one finished byte (4 symbols, the first on
// top) in every 32-bit lane; two saturating packs bring the 32 finished bytes of four vectors together, a byte shuffle
// and a lane permutation put them in word order (byte 7 - l of word t = symbols 4 l .. 4 l + 3 of vector t)
static inline __m256i pack128x2(__m256i v0, __m256i v1, __m256i v2, __m256i v3) {
    const __m256i m1 = _mm256_set1_epi16(0x0104), m2 = _mm256_set1_epi32(0x00010010);
    const __m256i y0 = _mm256_madd_epi16(_mm256_maddubs_epi16(v0, m1), m2);
    const __m256i y1 = _mm256_madd_epi16(_mm256_maddubs_epi16(v1, m1), m2);
    const __m256i y2 = _mm256_madd_epi16(_mm256_maddubs_epi16(v2, m1), m2);
    const __m256i y3 = _mm256_madd_epi16(_mm256_maddubs_epi16(v3, m1), m2);
    const __m256i b = _mm256_packus_epi16(_mm256_packus_epi32(y0, y1), _mm256_packus_epi32(y2, y3));
    // low half: symbols 0..15 of vectors 0..3 (4 bytes each), high half: symbols 16..31; reverse every 4-byte group
    const __m256i rev = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12, 3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8,
                                         15, 14, 13, 12);
    const __m256i r = _mm256_shuffle_epi8(b, rev);
    return _mm256_permutevar8x32_epi32(r, _mm256_setr_epi32(4, 0, 5, 1, 6, 2, 7, 3));
}
#endif

template <int BITS>
bool pack_stream_range(const uint8_t *ranks, uint64_t n, uint32_t sigma, uint64_t *words, uint64_t w_lo, uint64_t w_hi) {
    constexpr uint32_t SPW = 64 / BITS;
    constexpr uint32_t CPW = SPW / 8;
    const uint64_t full_hi = std::min<uint64_t>(w_hi, n / SPW);  // words below it are whole
    uint64_t w = w_lo;
    bool bad = false;
#if defined(__AVX2__)
    __m256i vmax = _mm256_setzero_si256();
    if (BITS == 2) {
        for (; w + 4 <= full_hi; w += 4) {
            const __m256i *src = reinterpret_cast<const __m256i *>(ranks + w * 32);
            const __m256i v0 = _mm256_loadu_si256(src), v1 = _mm256_loadu_si256(src + 1), v2 = _mm256_loadu_si256(src + 2),
                          v3 = _mm256_loadu_si256(src + 3);
            vmax = _mm256_max_epu8(_mm256_max_epu8(vmax, _mm256_max_epu8(v0, v1)), _mm256_max_epu8(v2, v3));
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(words + w), pack128x2(v0, v1, v2, v3));
        }
    } else if (BITS == 4) {
        for (; w + 2 <= full_hi; w += 2) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(ranks + w * 16));
            vmax = _mm256_max_epu8(vmax, v);
            pack32x4(v, words[w], words[w + 1]);
        }
    }
    {
        alignas(32) uint8_t mx[32];
        _mm256_store_si256(reinterpret_cast<__m256i *>(mx), vmax);
        for (int j = 0; j < 32; ++j) bad |= mx[j] >= sigma;
    }
#endif
    for (; w < w_hi; ++w) {  // the words the vector loops left, and the one the run ends in: symbol by symbol
        uint64_t acc = 0;
        for (uint32_t c = 0; c < CPW; ++c) {
            uint64_t v = 0;
            for (uint32_t j = 0; j < 8; ++j) {
                const uint64_t s = w * SPW + 8 * c + j;
                if (s < n) {
                    bad |= ranks[s] >= sigma;
                    v |= (uint64_t)ranks[s] << (8 * j);
                }
            }
            acc |= pack8<BITS>(v) << (64 - 8 * BITS * (c + 1));
        }
        words[w] = acc;
    }
    return !bad;
}

}  // namespace

uint64_t pack_stream_words(uint64_t n, uint32_t bits) { return (n + 64 / bits - 1) / (64 / bits); }

bool pack_stream_host(const uint8_t *ranks, uint64_t n, uint32_t bits, uint32_t sigma, uint64_t *words, unsigned part,
                      unsigned n_parts) {
    const uint64_t n_words = pack_stream_words(n, bits);
    // parts are cut at multiples of four words (the vector loop's step)
    const uint64_t quads = (n_words + 3) / 4;
    const uint64_t w_lo = std::min(n_words, quads * part / n_parts * 4), w_hi = std::min(n_words, quads * (part + 1) / n_parts * 4);
    if (bits == 2) return pack_stream_range<2>(ranks, n, sigma, words, w_lo, w_hi);
    if (bits == 4) return pack_stream_range<4>(ranks, n, sigma, words, w_lo, w_hi);
    return pack_stream_range<8>(ranks, n, sigma, words, w_lo, w_hi);
}

bool pack_queries_host(const uint8_t *q_ranks, const uint64_t *q_offsets, uint64_t q_begin, uint64_t q_end, uint32_t bits,
                       uint32_t sigma, uint32_t stride, uint64_t *words, unsigned part, unsigned n_parts) {
    const uint64_t n = q_end - q_begin;
    const uint64_t lo = q_begin + n * part / n_parts, hi = q_begin + n * (part + 1) / n_parts;
    const uint8_t *safe_end = q_ranks + q_offsets[q_end];  // the batch's last byte + 1 (the buffer may end there)
    if (bits == 2) return pack_range<2>(q_ranks, q_offsets, q_begin, lo, hi, sigma, stride, words, safe_end);
    if (bits == 4) return pack_range<4>(q_ranks, q_offsets, q_begin, lo, hi, sigma, stride, words, safe_end);
    return pack_range<8>(q_ranks, q_offsets, q_begin, lo, hi, sigma, stride, words, safe_end);
}

}  // namespace kb
