// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for seqan3::range_innermost_value_t (kmer_index.hpp:574).
#pragma once
#include <ranges>
#include <type_traits>
namespace seqan3
{
    template<typename R>
    using range_innermost_value_t = std::ranges::range_value_t<std::remove_cvref_t<R>>;
}
