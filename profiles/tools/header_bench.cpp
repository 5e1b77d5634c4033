// What a user of the drop-in header gets for a large batch held the reference's way -- std::vector<std::vector<dna4>>,
// pageable memory, one heap block per query -- next to the flat C-ABI call on the same batch. Not part of the library.
// Build: g++ -std=c++20 -O2 -I include profiles/tools/header_bench.cpp -o header_bench -L kmer_index_b200 -lkmer_b200
//        -Wl,-rpath,$PWD/kmer_index_b200 -pthread   (one line)
// usage: header_bench [text_symbols] [queries]
#include <kmer_index.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 3000000000ull;
    const uint64_t Q = argc > 2 ? strtoull(argv[2], nullptr, 10) : 100000000ull;
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::vector<kmer::dna4> text(n);
    std::vector<std::vector<kmer::dna4>> queries(Q);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (uint64_t i = n * t / T; i < n * (t + 1) / T; ++i) text[i] = kmer::dna4(uint8_t(mix(i) & 3));
                for (uint64_t q = Q * t / T; q < Q * (t + 1) / T; ++q) {
                    const uint64_t h = mix(q ^ 0xABCDEFull);
                    const uint32_t m = 16 + uint32_t(h % 49);
                    auto& v = queries[q];
                    v.resize(m);
                    if (q % 64 == 0 && n > m) {  // a planted window now and then, so that there are hits
                        const uint64_t at = (h >> 8) % (n - m);
                        for (uint32_t j = 0; j < m; ++j) v[j] = kmer::dna4(uint8_t(mix(at + j) & 3));
                    } else {
                        for (uint32_t j = 0; j < m; ++j) v[j] = kmer::dna4(uint8_t(mix(h + j) & 3));
                    }
                }
            });
        for (auto& x : th) x.join();
    }
    std::printf("text %.2e symbols, %.2e queries of 16-64 symbols in std::vector<std::vector<dna4>>\n", double(n), double(Q));
    auto t0 = std::chrono::steady_clock::now();
    auto index = kmer::make_kmer_index<16>(text);
    std::printf("make_kmer_index<16>(text)            %9.1f ms  (pageable std::vector storage handed over in place)\n", ms_since(t0));
    for (int rep = 0; rep < 3; ++rep) {
        t0 = std::chrono::steady_clock::now();
        auto res = index.search_batch(queries);
        const double ms = ms_since(t0);
        std::printf("index.search_batch(queries)          %9.1f ms  %.3e queries/s  hits %llu\n", ms, Q / (ms * 1e-3),
                    (unsigned long long)res.positions.size());
    }
    // the same batch, flat: what the C ABI takes directly (ranks back to back + offsets)
    std::vector<uint8_t> flat;
    std::vector<uint64_t> off(Q + 1, 0);
    for (uint64_t q = 0; q < Q; ++q) off[q + 1] = off[q] + queries[q].size();
    flat.resize(off[Q]);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (uint64_t q = Q * t / T; q < Q * (t + 1) / T; ++q)
                    for (size_t j = 0; j < queries[q].size(); ++j) flat[off[q] + j] = uint8_t(queries[q][j].to_rank());
            });
        for (auto& x : th) x.join();
    }
    for (int rep = 0; rep < 3; ++rep) {
        t0 = std::chrono::steady_clock::now();
        kmer_b200_result* r = nullptr;
        if (kmer_b200_search_batch(index.native_handle(), flat.data(), off.data(), Q, UINT32_MAX, &r) != 0) {
            std::printf("error: %s\n", kmer_b200_last_error());
            return 1;
        }
        const double ms = ms_since(t0);
        std::printf("kmer_b200_search_batch (flat, pageable) %6.1f ms  %.3e queries/s  hits %llu\n", ms, Q / (ms * 1e-3),
                    (unsigned long long)kmer_b200_result_n_positions(r));
        kmer_b200_result_free(r);
    }
    return 0;
}
