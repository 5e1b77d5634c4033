// TEST INFRASTRUCTURE ONLY (oracle/): eager stand-in for seqan3::views::kmer_hash with an
// ungapped shape. The value per window is forced by the reference itself: build inserts
// `kmer_hash` values (kmer_index.hpp:157-165) and search looks up `hash()` values
// (kmer_index.hpp:56-73), so both must equal sum_i rank(t[p+i]) * sigma^(k-1-i).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <seqan3/alphabet/concept.hpp>

namespace seqan3
{
    struct ungapped { std::uint8_t value; };

    struct shape
    {
        std::size_t k;
        shape(ungapped u) : k(u.value) {}
    };

    namespace views
    {
        struct kmer_hash_closure { std::size_t k; };

        inline kmer_hash_closure kmer_hash(shape s) { return {s.k}; }

        template<typename range_t>
        std::vector<std::size_t> operator|(range_t&& text, kmer_hash_closure c)
        {
            using alphabet_t = std::remove_cvref_t<decltype(*text.begin())>;
            constexpr std::size_t sigma = alphabet_size<alphabet_t>;
            std::vector<std::size_t> out;
            const std::size_t n = text.size();
            if (n < c.k) return out;
            out.reserve(n - c.k + 1);
            std::size_t top = 1;                       // sigma^(k-1)
            for (std::size_t i = 1; i < c.k; ++i) top *= sigma;
            std::size_t h = 0;
            for (std::size_t i = 0; i < c.k; ++i) h = h * sigma + to_rank(text[i]);
            out.push_back(h);
            for (std::size_t p = 1; p + c.k <= n; ++p)
            {
                h = (h - to_rank(text[p - 1]) * top) * sigma + to_rank(text[p + c.k - 1]);
                out.push_back(h);
            }
            return out;
        }
    }
}
