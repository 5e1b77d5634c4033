// TEST INFRASTRUCTURE ONLY (oracle/): robin_hood::unordered_map (martinus/robin-hood-hashing,
// unpinned in the reference, kmer_index.hpp:23,52) is used as a plain container whose iteration
// order is never observed, so std::unordered_map is result-equivalent.
#pragma once
#include <unordered_map>
namespace robin_hood
{
    template<typename K, typename V>
    using unordered_map = std::unordered_map<K, V>;
}
