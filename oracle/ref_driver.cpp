// TEST INFRASTRUCTURE ONLY (oracle/). Never linked, imported or executed by the product path.
//
// C-ABI wrapper around the UNMODIFIED reference implementation (Clemapfel/kmer_index), compiled
// from the headers where they lie under /root/reference (see oracle/Makefile: the only change is
// the one-token `constexpr` fix on kmer_index.hpp:401 applied to a throw-away copy in a temp dir,
// without which nothing in the reference compiles). Output: oracle/_ref/libkmer_ref.so.
//
// It is used (a) to pin oracle/kmer_oracle.c, (b) to generate tests/golden/ fixtures, and
// (c) as the CPU baseline ("kind": "reference") in bench.py.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <future>
#include <iostream>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <cmath>
#include <unordered_map>
#include <condition_variable>
#include <queue>
#include <functional>
#include <map>
#include <mutex>

// expose the scheme tables (_optimal_nk_sum, _use_multi_search_scheme) for pinning; all std
// headers the reference pulls in are already included above, so only reference code is affected.
#define private public
#define protected public
#include <kmer_index.hpp>
#include <choose_best_k.hpp>
#undef private
#undef protected

namespace
{
    struct index_base
    {
        virtual ~index_base() = default;
        // returns 0 ok, 1 std::invalid_argument, 2 std::out_of_range, 3 other exception
        virtual int search(const uint8_t* q, size_t m, std::vector<uint32_t>& out) const = 0;
        virtual size_t scheme(size_t m, uint32_t* out, size_t cap, int* use_multi) const = 0;
    };

    template<typename alphabet_t, size_t... ks>
    struct index_impl final : index_base
    {
        using index_t = kmer::kmer_index<alphabet_t, uint32_t, ks...>;
        std::unique_ptr<index_t> idx;

        index_impl(const uint8_t* ranks, size_t n, size_t n_threads)
        {
            static_assert(sizeof(alphabet_t) == 1);
            std::vector<alphabet_t> text(n);
            std::memcpy(static_cast<void*>(text.data()), ranks, n);
            // same call as kmer::make_kmer_index<ks...>(text, n_threads) (kmer_index.hpp:569-579)
            idx = std::make_unique<index_t>(text, n_threads);
        }

        int search(const uint8_t* q, size_t m, std::vector<uint32_t>& out) const override
        {
            std::vector<alphabet_t> query(m);
            std::memcpy(static_cast<void*>(query.data()), q, m);
            try
            {
                out = idx->search(query).to_vector();   // the observable used by test_main.cpp:41-42
                return 0;
            }
            catch (std::invalid_argument const&) { out.clear(); return 1; }
            catch (std::out_of_range const&)     { out.clear(); return 2; }
            catch (...)                          { out.clear(); return 3; }
        }

        size_t scheme(size_t m, uint32_t* out, size_t cap, int* use_multi) const override
        {
            auto const& s = idx->_optimal_nk_sum.at(m);
            *use_multi = idx->_use_multi_search_scheme.at(m) ? 1 : 0;
            for (size_t i = 0; i < s.size() && i < cap; ++i) out[i] = uint32_t(s[i]);
            return s.size();
        }
    };

    template<typename alphabet_t, size_t... ks>
    bool match(uint32_t sigma, const uint32_t* in_ks, uint32_t n_ks)
    {
        constexpr size_t want[] = {ks...};
        if (sigma != alphabet_t::alphabet_size || n_ks != sizeof...(ks)) return false;
        for (size_t i = 0; i < sizeof...(ks); ++i)
            if (in_ks[i] != want[i]) return false;
        return true;
    }

#define KREF_TRY(ALPHA, ...)                                                            \
    if (match<ALPHA, __VA_ARGS__>(sigma, ks, n_ks))                                     \
        return probe_only ? reinterpret_cast<index_base*>(1)                            \
                          : new index_impl<ALPHA, __VA_ARGS__>(ranks, n, n_threads);

    index_base* dispatch(uint32_t sigma, const uint32_t* ks, uint32_t n_ks, const uint8_t* ranks,
                         size_t n, size_t n_threads, bool probe_only)
    {
        using namespace seqan3;
        // BASELINE.json configs
        KREF_TRY(dna4, 10)
        KREF_TRY(dna4, 12)
        KREF_TRY(dna4, 5, 7, 9, 11, 13)
        KREF_TRY(dna15, 8)
        KREF_TRY(aa27, 5)
        KREF_TRY(dna4, 16)
        // low-entropy / edge-case pinning and test_main.cpp:76-78 shapes
        KREF_TRY(dna4, 3)
        KREF_TRY(dna4, 5)
        KREF_TRY(dna4, 6)
        KREF_TRY(dna4, 14)
        KREF_TRY(dna4, 9, 10)
        KREF_TRY(dna4, 10, 11, 12)
        KREF_TRY(dna15, 5)
        KREF_TRY(dna15, 10)
        KREF_TRY(dna15, 5, 6, 7)
        KREF_TRY(dna15, 10, 11, 12)
        KREF_TRY(dna5, 4)
        KREF_TRY(aa27, 3)
        KREF_TRY(aa27, 9, 10)
        // k-mers wider than 64 packed bits / hashes wider than 32 bits
        KREF_TRY(aa27, 12)
        KREF_TRY(dna5, 18)
        KREF_TRY(dna4, 20)
        KREF_TRY(dna15, 16)
        return nullptr;
    }
}

extern "C"
{
    int kref_supported(uint32_t sigma, const uint32_t* ks, uint32_t n_ks)
    {
        return dispatch(sigma, ks, n_ks, nullptr, 0, 1, true) != nullptr;
    }

    // build; *build_seconds = wall time of the reference constructor (text already in its vector)
    void* kref_create(uint32_t sigma, const uint32_t* ks, uint32_t n_ks, const uint8_t* ranks, uint64_t n,
                      uint32_t n_threads, double* build_seconds)
    {
        auto t0 = std::chrono::steady_clock::now();
        index_base* p = nullptr;
        try { p = dispatch(sigma, ks, n_ks, ranks, n, std::max(1u, n_threads), false); }
        catch (...) { p = nullptr; }
        auto t1 = std::chrono::steady_clock::now();
        if (build_seconds) *build_seconds = std::chrono::duration<double>(t1 - t0).count();
        return p;
    }

    void kref_destroy(void* h) { delete static_cast<index_base*>(h); }

    // Q queries (ranks concatenated, q_off[Q+1]) striped over the reference's own thread_pool, each
    // task calling index.search(q).to_vector(). counts/status are caller-allocated [Q]; *positions is
    // malloc'ed (free with kref_free) and holds the per-query sorted hit lists back to back.
    int kref_search_batch(void* h, const uint8_t* q, const uint64_t* q_off, uint64_t Q, uint32_t n_threads,
                          uint64_t* counts, uint8_t* status, uint32_t** positions, uint64_t* total,
                          double* search_seconds, int keep_positions)
    {
        auto* idx = static_cast<index_base*>(h);
        n_threads = std::max(1u, n_threads);
        std::vector<std::vector<uint32_t>> per_thread_pos(n_threads);
        auto t0 = std::chrono::steady_clock::now();
        {
            kmer::detail::thread_pool pool{n_threads};
            std::vector<std::future<void>> futures;
            for (uint32_t t = 0; t < n_threads; ++t)
            {
                uint64_t lo = Q * t / n_threads, hi = Q * (t + 1) / n_threads;
                futures.emplace_back(pool.execute([=, &per_thread_pos]() {
                    std::vector<uint32_t> out;
                    for (uint64_t i = lo; i < hi; ++i)
                    {
                        status[i] = uint8_t(idx->search(q + q_off[i], q_off[i + 1] - q_off[i], out));
                        counts[i] = out.size();
                        if (keep_positions)
                            per_thread_pos[t].insert(per_thread_pos[t].end(), out.begin(), out.end());
                    }
                }));
            }
            for (auto& f : futures) f.get();
        }
        auto t1 = std::chrono::steady_clock::now();
        if (search_seconds) *search_seconds = std::chrono::duration<double>(t1 - t0).count();
        uint64_t tot = 0;
        for (auto& v : per_thread_pos) tot += v.size();
        if (total) *total = tot;
        if (positions)
        {
            *positions = static_cast<uint32_t*>(std::malloc(std::max<uint64_t>(tot, 1) * sizeof(uint32_t)));
            uint64_t o = 0;
            for (auto& v : per_thread_pos)
            {
                if (!v.empty()) std::memcpy(*positions + o, v.data(), v.size() * sizeof(uint32_t));
                o += v.size();
            }
        }
        return 0;
    }

    void kref_free(void* p) { std::free(p); }

    // scheme table row for query length m (kmer_index.hpp:404-405 after choose_search_scheme :407-476)
    uint64_t kref_scheme(void* h, uint64_t m, uint32_t* out, uint64_t cap, int* use_multi)
    {
        return static_cast<index_base*>(h)->scheme(m, out, cap, use_multi);
    }

    uint64_t kref_fast_pow(uint64_t base, uint8_t exp) { return kmer::detail::fast_pow(base, exp); }

    // choose_best_k<dna4>(interval, n_k) (choose_best_k.hpp:12-60); alphabet is unused by the function body
    uint64_t kref_choose_best_k(const uint64_t* lens, uint64_t n_lens, uint64_t n_k, uint64_t* out)
    {
        std::vector<size_t> interval(lens, lens + n_lens);
        auto r = choose_best_k<seqan3::dna4>(interval, n_k);
        for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
        return r.size();
    }

    uint32_t kref_hardware_concurrency() { return std::max(1u, std::thread::hardware_concurrency()); }
}
