// py_s3fifo.h — host-side S3FIFOCache / ShardedS3FIFOCache for the Python tools (cache_utils.py, tournament.py hold
// such objects; reference: src/s3fifo_cache.h:15-318, bound at py_wrapper.cc:222-259). The self-play engine itself
// uses the device table (az_engine_logic.h: cache_find / cache_insert); this class is the reference's host container
// with the same observable behaviour — admission (unseen keys enter the Small queue, keys remembered by the ghost list
// enter Main), eviction (Small first: entries hit while in Small are promoted to Main with frequency 0, the others are
// evicted into the ghost list; Main with one second chance per frequency point), a 2-bit saturating frequency bumped by
// find(), an existing key is never overwritten, and the six counters.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace b2az_host {

class Fifo {  // fixed-capacity ring of slot ids / keys
 public:
  explicit Fifo(size_t cap) : buf_(cap ? cap : 1), cap_(cap ? cap : 1) {}
  bool empty() const { return n_ == 0; }
  size_t size() const { return n_; }
  void push(uint64_t x) { buf_[(head_ + n_) % cap_] = x; ++n_; }
  uint64_t pop() { const uint64_t x = buf_[head_]; head_ = (head_ + 1) % cap_; --n_; return x; }
  // overwrite the oldest entry with x and return it (the ring is full)
  uint64_t replace_oldest(uint64_t x) { const uint64_t old = buf_[head_]; buf_[head_] = x; head_ = (head_ + 1) % cap_; return old; }
 private:
  std::vector<uint64_t> buf_;
  size_t cap_, head_ = 0, n_ = 0;
};

class S3FIFOCache {
 public:
  S3FIFOCache(uint32_t max_size, uint32_t ghost_size, uint32_t num_policy, uint32_t num_value)
      : cap_(max_size), ghost_cap_(ghost_size), np_(num_policy), nv_(num_value), small_(max_size), main_(max_size),
        ghost_(ghost_size) {
    entries_.reserve(max_size);
  }
  bool find(uint64_t key, float* policy_out, float* value_out) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = index_.find(key);
    if (it == index_.end()) {
      ++misses_;
      if (ghost_cap_ > 0 && ghost_keys_.count(key)) ++reinserts_;
      return false;
    }
    ++hits_;
    Entry& e = entries_[it->second];
    if (e.freq < 3) ++e.freq;
    std::memcpy(policy_out, e.data.get(), np_ * sizeof(float));
    std::memcpy(value_out, e.data.get() + np_, nv_ * sizeof(float));
    return true;
  }
  void insert(uint64_t key, const float* policy, const float* value) {
    std::lock_guard<std::mutex> lk(mu_);
    if (cap_ == 0 || index_.count(key)) return;
    const bool remembered = ghost_cap_ > 0 && ghost_keys_.erase(key) > 0;  // its ring entry goes stale, as in the reference
    const uint32_t slot = entries_.size() < cap_ ? new_slot() : evict();
    Entry& e = entries_[slot];
    e.key = key;
    e.freq = 0;
    if (!e.data) e.data.reset(new float[np_ + nv_]);
    std::memcpy(e.data.get(), policy, np_ * sizeof(float));
    std::memcpy(e.data.get() + np_, value, nv_ * sizeof(float));
    index_[key] = slot;
    (remembered ? main_ : small_).push(slot);
  }
  size_t hits() const { return hits_; }
  size_t misses() const { return misses_; }
  size_t evictions() const { return evictions_; }
  size_t reinserts() const { return reinserts_; }
  size_t size() const { return index_.size(); }
  size_t max_size() const { return cap_; }
  uint32_t num_policy() const { return np_; }
  uint32_t num_value() const { return nv_; }

 private:
  struct Entry {
    uint64_t key = 0;
    uint8_t freq = 0;
    std::unique_ptr<float[]> data;  // allocated when the slot is first used: an empty cache costs no memory
  };
  uint32_t new_slot() {
    entries_.emplace_back();
    return (uint32_t)entries_.size() - 1;
  }
  void remember(uint64_t key) {
    if (ghost_cap_ == 0) return;
    if (ghost_.size() >= ghost_cap_) ghost_keys_.erase(ghost_.replace_oldest(key));
    else ghost_.push(key);
    ghost_keys_.insert(key);
  }
  uint32_t evict() {
    while (!small_.empty()) {
      const uint32_t slot = (uint32_t)small_.pop();
      Entry& e = entries_[slot];
      if (e.freq) {  // hit while in Small: promoted
        e.freq = 0;
        main_.push(slot);
        continue;
      }
      remember(e.key);
      index_.erase(e.key);
      ++evictions_;
      return slot;
    }
    for (;;) {
      const uint32_t slot = (uint32_t)main_.pop();
      Entry& e = entries_[slot];
      if (e.freq) {  // second chance
        --e.freq;
        main_.push(slot);
        continue;
      }
      index_.erase(e.key);
      ++evictions_;
      return slot;
    }
  }
  uint32_t cap_, ghost_cap_, np_, nv_;
  std::vector<Entry> entries_;
  std::unordered_map<uint64_t, uint32_t> index_;
  Fifo small_, main_, ghost_;
  std::unordered_set<uint64_t> ghost_keys_;
  size_t hits_ = 0, misses_ = 0, evictions_ = 0, reinserts_ = 0;
  std::mutex mu_;
};

class ShardedS3FIFOCache {  // shard = key % shards (s3fifo_cache.h:229-318)
 public:
  ShardedS3FIFOCache(uint32_t max_size, uint32_t shards, uint32_t ghost_size, uint32_t num_policy, uint32_t num_value) {
    if (shards == 0) shards = 1;
    for (uint32_t i = 0; i < shards; ++i)
      shards_.push_back(std::make_unique<S3FIFOCache>(max_size / shards, ghost_size / shards, num_policy, num_value));
  }
  bool find(uint64_t key, float* p, float* v) { return shard(key).find(key, p, v); }
  void insert(uint64_t key, const float* p, const float* v) { shard(key).insert(key, p, v); }
#define B2AZ_SUM(name) size_t name() const { size_t s = 0; for (auto& c : shards_) s += c->name(); return s; }
  B2AZ_SUM(hits) B2AZ_SUM(misses) B2AZ_SUM(evictions) B2AZ_SUM(reinserts) B2AZ_SUM(size) B2AZ_SUM(max_size)
#undef B2AZ_SUM
 private:
  S3FIFOCache& shard(uint64_t key) { return *shards_[key % shards_.size()]; }
  std::vector<std::unique_ptr<S3FIFOCache>> shards_;
};

}  // namespace b2az_host
