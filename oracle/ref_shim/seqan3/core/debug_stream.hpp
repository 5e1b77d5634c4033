// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for seqan3::debug_stream. The reference only
// prints diagnostics through it (choose_best_k.hpp:53); swallow everything.
#pragma once
namespace seqan3
{
    struct null_debug_stream
    {
        template<typename T>
        null_debug_stream& operator<<(T const&) { return *this; }
    };
    inline null_debug_stream debug_stream{};
}
