// Probe (not part of the library): can the GPU box's host cores pack 1-byte dna4 ranks into 2-bit words faster than
// PCIe moves the unpacked bytes (~55 GB/s)? Decides whether kmer_b200_search_batch packs on the host before H2D.
// Build: g++ -O3 -march=x86-64-v3 -pthread -o host_pack_probe host_pack_probe.cpp
#include <immintrin.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void pack2(const uint8_t *src, uint64_t n_sym, uint64_t *dst) {  // n_sym multiple of 32
    for (uint64_t w = 0; w < n_sym / 32; ++w) {
        uint64_t acc = 0;
        for (int c = 0; c < 4; ++c) {
            uint64_t v;
            memcpy(&v, src + w * 32 + c * 8, 8);
            // bswap puts symbol 0 in the top byte, pext keeps the low 2 bits of every byte: 16 packed bits
            const uint64_t f = _pext_u64(__builtin_bswap64(v), 0x0303030303030303ull);
            acc |= f << (48 - 16 * c);
        }
        dst[w] = acc;
    }
}

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 32);
    uint8_t *src = (uint8_t *)aligned_alloc(4096, n);
    uint64_t *dst = (uint64_t *)aligned_alloc(4096, n / 4);
    uint8_t *cpy = (uint8_t *)aligned_alloc(4096, n);
    const unsigned hw = std::thread::hardware_concurrency();
    {   // first touch in parallel
        std::vector<std::thread> th;
        for (unsigned t = 0; t < hw; ++t)
            th.emplace_back([=] {
                const uint64_t lo = n * t / hw, hi = n * (t + 1) / hw;
                for (uint64_t i = lo; i < hi; ++i) src[i] = (uint8_t)((i * 2654435761u) >> 30);
                memset(cpy + lo, 0, hi - lo);
                memset((uint8_t *)dst + lo / 4, 0, (hi - lo) / 4);
            });
        for (auto &x : th) x.join();
    }
    printf("hardware_concurrency %u, %.1f GB of ranks\n", hw, n / 1e9);
    for (unsigned nt : {1u, 2u, 4u, 8u, 16u, 32u, 64u}) {
        if (nt > hw && nt != 1) continue;
        for (int what = 0; what < 2; ++what) {
            double best = 1e30;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nt; ++t)
                    th.emplace_back([=] {
                        const uint64_t lo = (n * t / nt) & ~31ull, hi = (n * (t + 1) / nt) & ~31ull;
                        if (what == 0) pack2(src + lo, hi - lo, dst + lo / 32);
                        else memcpy(cpy + lo, src + lo, hi - lo);
                    });
                for (auto &x : th) x.join();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            printf("threads %2u  %-6s %.1f ms  %.1f GB/s of input\n", nt, what == 0 ? "pack2" : "memcpy", best * 1e3, n / 1e9 / best);
        }
    }
    uint64_t s = 0;
    for (uint64_t i = 0; i < n / 32; i += 4097) s += dst[i];
    printf("checksum %llu\n", (unsigned long long)s);
    return 0;
}
