// oracle/shim/absl/container/flat_hash_set.h — std::unordered_set stand-in. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <unordered_set>

#include "absl/hash/hash.h"

namespace absl {
template <typename K, typename H = absl::Hash<K>, typename E = std::equal_to<K>>
using flat_hash_set = std::unordered_set<K, H, E>;
}  // namespace absl
