// oracle/shim/absl/hash/hash.h — stand-in for the absl::Hash subset the reference uses
// (abseil-cpp 20250814.1, subprojects/abseil-cpp.wrap:2; not vendored, not fetchable).
// TEST INFRASTRUCTURE ONLY. Real absl hashes are salted per process, so hash VALUES are not a
// parity target — only equality classes are. This shim uses a
// deterministic 64-bit mix so oracle runs are reproducible.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <type_traits>
#include <typeindex>
#include <utility>
#include <vector>

namespace absl {

namespace shim_detail {
inline uint64_t mix64(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  return h;
}
inline uint64_t mix_bytes(uint64_t h, const void* p, size_t n) {
  const auto* b = static_cast<const unsigned char*>(p);
  h = mix64(h, n);
  while (n >= 8) {
    uint64_t w;
    std::memcpy(&w, b, 8);
    h = mix64(h, w);
    b += 8;
    n -= 8;
  }
  if (n) {
    uint64_t w = 0;
    std::memcpy(&w, b, n);
    h = mix64(h, w);
  }
  return h;
}
}  // namespace shim_detail

class MixState;
template <typename T>
MixState shim_hash_one(MixState h, const T& v);

// The concrete hash state ("H" in AbslHashValue(H, const T&)).
class MixState {
 public:
  uint64_t v = 0x243f6a8885a308d3ULL;
  MixState() = default;
  MixState(const MixState&) = default;
  MixState& operator=(const MixState&) = default;

  template <typename... Ts>
  static MixState combine(MixState h, const Ts&... vs) {
    ((h = shim_hash_one(std::move(h), vs)), ...);
    return h;
  }
  template <typename T>
  static MixState combine_contiguous(MixState h, const T* p, size_t n) {
    if constexpr (std::is_arithmetic_v<T> || std::is_enum_v<T>) {
      h.v = shim_detail::mix_bytes(h.v, p, n * sizeof(T));
    } else {
      h.v = shim_detail::mix64(h.v, n);
      for (size_t i = 0; i < n; ++i) h = shim_hash_one(std::move(h), p[i]);
    }
    return h;
  }
};

template <typename T, typename = void>
struct has_absl_hash_value : std::false_type {};
template <typename T>
struct has_absl_hash_value<T, std::void_t<decltype(AbslHashValue(std::declval<MixState>(), std::declval<const T&>()))>>
    : std::true_type {};

template <typename T>
MixState shim_hash_one(MixState h, const T& v) {
  if constexpr (std::is_arithmetic_v<T> || std::is_enum_v<T>) {
    uint64_t w = 0;
    std::memcpy(&w, &v, sizeof(T) < 8 ? sizeof(T) : 8);
    h.v = shim_detail::mix64(h.v, w);
    return h;
  } else if constexpr (std::is_same_v<T, std::type_index>) {
    h.v = shim_detail::mix64(h.v, static_cast<uint64_t>(v.hash_code()));
    return h;
  } else if constexpr (std::is_same_v<T, std::string> || std::is_same_v<T, std::string_view>) {
    h.v = shim_detail::mix_bytes(h.v, v.data(), v.size());
    return h;
  } else if constexpr (std::is_pointer_v<T>) {
    h.v = shim_detail::mix64(h.v, reinterpret_cast<uintptr_t>(v));
    return h;
  } else if constexpr (has_absl_hash_value<T>::value) {
    return AbslHashValue(std::move(h), v);
  } else {
    static_assert(sizeof(T) == 0, "absl shim: unsupported type in hash combine");
    return h;
  }
}

template <typename T>
MixState AbslHashValue(MixState h, const std::vector<T>& v) {
  return MixState::combine_contiguous(std::move(h), v.data(), v.size());
}
template <typename A, typename B>
MixState AbslHashValue(MixState h, const std::pair<A, B>& p) {
  return MixState::combine(std::move(h), p.first, p.second);
}
template <typename T>
MixState AbslHashValue(MixState h, const std::shared_ptr<T>& p) {
  return MixState::combine(std::move(h), p.get());
}

// Type-erased handle (absl::HashState): wraps a pointer to the concrete state.
class HashState {
 public:
  static HashState Create(MixState* s) { return HashState(s); }
  HashState(const HashState&) = default;
  HashState(HashState&&) = default;
  HashState& operator=(const HashState&) = default;
  HashState& operator=(HashState&&) = default;

  template <typename... Ts>
  static HashState combine(HashState h, const Ts&... vs) {
    *h.s_ = MixState::combine(*h.s_, vs...);
    return h;
  }
  template <typename T>
  static HashState combine_contiguous(HashState h, const T* p, size_t n) {
    *h.s_ = MixState::combine_contiguous(*h.s_, p, n);
    return h;
  }

 private:
  explicit HashState(MixState* s) : s_(s) {}
  MixState* s_;
};

template <typename T>
uint64_t HashOf(const T& v) {
  MixState h;
  h = shim_hash_one(std::move(h), v);
  return h.v;
}

template <typename T>
struct Hash {
  size_t operator()(const T& v) const { return static_cast<size_t>(HashOf(v)); }
};

}  // namespace absl
