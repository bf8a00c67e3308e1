// oracle/shim/absl/container/flat_hash_map.h — std::unordered_map stand-in. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <unordered_map>

#include "absl/hash/hash.h"

namespace absl {
template <typename K, typename V, typename H = absl::Hash<K>, typename E = std::equal_to<K>>
using flat_hash_map = std::unordered_map<K, V, H, E>;
}  // namespace absl
