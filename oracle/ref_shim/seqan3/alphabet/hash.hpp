// TEST INFRASTRUCTURE ONLY (oracle/): nothing from seqan3/alphabet/hash.hpp is used on the path.
#pragma once
#include <seqan3/alphabet/concept.hpp>
