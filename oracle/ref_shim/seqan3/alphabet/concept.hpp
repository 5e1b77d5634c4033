// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the seqan3 alphabet
// concept so the reference headers under /root/reference compile without seqan3.
// Own code, not reference code. seqan3 is an un-vendored, unpinned dependency of
// the reference (kmer_index.hpp:8-12); only `alphabet`, `alphabet_size`,
// `to_rank` and rank-only alphabet types are needed on the build/search path.
#pragma once
#include <concepts>
#include <cstddef>
#include <cstdint>

namespace seqan3
{
    template<typename T>
    concept alphabet = requires(T const a) {
        { a.to_rank() } -> std::convertible_to<std::size_t>;
        { T::alphabet_size } -> std::convertible_to<std::size_t>;
    };

    template<typename T>
    inline constexpr std::size_t alphabet_size = T::alphabet_size;

    template<alphabet T>
    constexpr auto to_rank(T const a) { return a.to_rank(); }

    // rank-only alphabet of SIGMA symbols: one byte per symbol, as seqan3's dna4/dna15/aa27 store them
    template<std::size_t SIGMA>
    struct rank_alphabet
    {
        static constexpr std::size_t alphabet_size = SIGMA;
        std::uint8_t r{0};
        constexpr std::uint8_t to_rank() const { return r; }
        constexpr rank_alphabet& assign_rank(std::uint8_t x) { r = x; return *this; }
        friend constexpr bool operator==(rank_alphabet a, rank_alphabet b) { return a.r == b.r; }
        friend constexpr bool operator!=(rank_alphabet a, rank_alphabet b) { return a.r != b.r; }
    };

    using dna4  = rank_alphabet<4>;
    using dna5  = rank_alphabet<5>;
    using dna15 = rank_alphabet<15>;
    using aa27  = rank_alphabet<27>;
}
