// az_selfplay.h — PlayManager::play (play_manager.cc:258-600) over the tafl games, on the device: the scheduler loop
// of the reference (search `mcts_visits` simulations with the tree of the seat to move, temperature schedule, acting
// rule, training-sample capture, both seats' trees re-rooted, game end / restart, Gumbel arming and root noise on the
// reused root) with one warp per GAME SLOT, built on the wide-tree search of az_forest.h.
//
// Layout. Slot g owns trees 2g and 2g+1 of a forest (GameData::mcts[0..1], play_manager.h:36) and ONE pcg32 stream
// (kept in tree 2g, ForestView::rng_pair): in the reference every draw of a worker thread — the child shuffles of
// BOTH seats' trees (update_root expands an unexpanded root, mcts.cc:154-156), Dirichlet / Gumbel noise, pick_move —
// comes from the thread-local generator, so slot g == the unmodified PlayManager with concurrent_games = 1 and
// games_to_play = games_per_slot run after MCTS::seed_thread_rng(seed + g). Every slot needs exactly `visits`
// simulations per move (no playout cap here), so all slots advance in lock step: k_sp_search (the hot kernel:
// visits x (find_leaf, evaluator, process_result)) then k_sp_move (cold, once per move).
// Training samples are staged per slot ([max_turns] rows of canonical + policy target) and appended to the output
// ring when the game ends, last move first like PlayManager's partial_history.pop_back() loop (play_manager.cc:446-460).
#pragma once

namespace b2az {

constexpr int kSpMaxPerms = 8;
struct SpSlot {                       // b2az_tafl_selfplay_slot in include/b2az.h (same layout)
  u32 active;                         // the slot is still cycling (play_manager.cc:506-508)
  u32 games_started, games_completed;
  u32 pending;                        // samples staged for the current game (GameData::partial_history.size())
  u32 move_count, full_move_count;    // GameData counters of the current game
  u32 total_move_count, total_full_move_count, game_length;  // PlayManager::total_move_count_ / full_move_count_ / game_length_
  u32 picked;                         // scratch: the move pick_move returned
  u32 error;                          // 1 = output ring full (samples dropped)
  u32 capped;                         // GameData::capped: the current search is a fast (playout-cap) search
  double g_leaf_depth, g_entropy, g_valid_moves;  // GameData::total_avg_leaf_depth / total_search_entropy / total_valid_moves
  double leaf_depth, entropy, valid_moves;        // PlayManager::total_avg_leaf_depth_ / total_search_entropy_ / total_valid_moves_
  unsigned long long simulations;
  float scores[3];
  u32 playthrough;                    // GameData::playthrough (set once per game slot, never cleared: play_manager.cc:328-329)
  // fast (capped) searches: GameData::fast_* / PlayManager::fast_* (play_manager.cc:436-446, 489-505)
  u32 fast_move_count, total_fast_move_count;
  double g_fast_leaf_depth, g_fast_entropy, fast_leaf_depth, fast_entropy;
  float resign_scores[3];             // resign_scores_
  u16 resign_streak[2];               // GameData::resign_streak (never cleared between the games of a slot, like the reference)
  Pcg32 coin;                         // playout-cap and resign-playthrough coins (see sp_coin)
};
static_assert(sizeof(SpSlot) == 192, "b2az_tafl_selfplay_slot layout");

// Device position cache of the wide-tree engine (S3FIFOCache / ShardedS3FIFOCache, s3fifo_cache.h:15-318, for the tafl
// games and Star Gambit): 4-way set-associative table keyed by the 64-bit state key (the reference's cache is keyed by
// the 64-bit absl hash alone as well), dense value rows pi[A] + v[3] per entry (what the reference stores), a 2-bit
// frequency per way: a hit bumps it, an insert into a full set replaces the way with the lowest frequency and ages the
// others. An existing key is never overwritten (s3fifo_cache.h:84). Lookups (k_sp_find_leaf) and inserts
// (k_sp_process_result) run in different launches, so a reader never meets a half-written row.
struct SpCache {
  u64* keys;     // [sets * 4], 0 = empty
  u8* freq;      // [sets * 4]
  u32* stamp;    // [sets * 4] epoch of the launch that wrote the way
  u32 epoch;     // this insert launch's epoch (> 0)
  float* v;      // [sets * 4][3]
  float* pi;     // [sets * 4][A]
  u32 sets;      // power of two; 0 = no cache
  unsigned long long* ctr;  // hits, misses, inserts, evictions
};
// PlayManager's per-variant tables (variant_scores_ / variant_metrics_, play_manager.cc:468-484), one set per slot
struct SpVariantAcc {                 // b2az_variant_stats in include/b2az.h (same layout)
  float scores[3];
  u32 games_completed;
  u32 game_length, total_move_count, full_move_count, fast_move_count;
  double leaf_depth, entropy, valid_moves, fast_leaf_depth, fast_entropy;
};
static_assert(sizeof(SpVariantAcc) == 72, "b2az_variant_stats layout");
struct SpView {
  SpVariantAcc* variants;  // [n_games][4], null unless the game has variants (StarGambitUnifiedGS)
  SpCache cache;
  u32* wait;          // [n_games] (cache only): the slot's leaf missed the cache and waits for the evaluator
  u64* leaf_key;      // [n_games] (cache only): the waiting leaf's key
  SpSlot* slots;
  u32 n_games, games_per_slot, visits;  // visits: the largest search budget of any seat (launch sizing)
  // seat permutations (play_manager.cc:46-90, 213-221): slot g plays permutation g % n_perms (the reference hands the
  // permutations out round robin, i % num_perms at construction and games_started_ % num_perms afterwards: with
  // n_games a multiple of n_perms and the slots finishing in order that is g % n_perms for every game of slot g)
  u32 n_perms;
  u32 seat_visits[kSpMaxPerms][2], seat_cap_visits[kSpMaxPerms][2];  // seat_visits_ / seat_cap_visits_ (play_manager.cc:70-90)
  u8 perm_group[kSpMaxPerms][2];  // seat_perms_[perm][seat]: the model group that searches for the seat (play_manager.cc:577)
  float seat_resign_threshold[kSpMaxPerms][2];   // seat_resign_threshold_ (-2 = off), play_manager.cc:335-366
  u32 seat_resign_consecutive[kSpMaxPerms][2];   // seat_resign_consecutive_
  u32 random_groups;  // bit i: model group i is EvalType::RANDOM — its searches run dumb_eval inline (play_manager.cc:578-587)
  u8* leaf_group;     // [n_games] model group of the slot's waiting leaf (several groups only)
  u32 playout_cap, fast_search_uses_gumbel;
  u32 gumbel_targets;  // PlayParams::gumbel_enabled: the policy target is gumbel_improved_policy() (play_manager.cc:412-419)
  float playout_cap_percent, resign_percent, resign_playthrough_percent;
  float start_temp, final_temp, half_life;
  u32 history_enabled, policy_target_pruning, tree_reuse;
  u32 stage_rows;     // staging rows per slot: the game's max_turns (tafl) / a bound on the actions of a Star Gambit game
  u32 st_stride;      // floats per staged position (FGame::stage_floats: the planes for tafl, the position record for Star Gambit)
  u32 n_variant_half_life;
  float variant_half_life[4];  // temp_decay_half_life_by_variant
  float* st_canon;    // [n_games][stage_rows][st_stride]
  float* st_pi;       // [n_games][max_turns][A]
  u8* st_player;      // [n_games][max_turns]
  float* scratch_pi;  // [n_games][A]: the acting distribution
  float *out_canon, *out_v, *out_pi;  // output ring [out_cap] rows
  u32* out_slot;
  u32 out_cap;
  u32* out_count;     // rows appended since the last drain
};

#ifndef B2AZ_HOST_EMU
// a fresh MCTS object (make_mcts, play_manager.cc): empty root, counters zero, Gumbel state reset with no target
__device__ __forceinline__ void sp_reset_search(const ForestView& F, u32 t) {
  ForestTree& R = F.trees[t];
  R.n = 0; R.v = 0.0f; R.d = 0.0f; R.blk = 0; R.k = 0; R.player = 0; R.term = 0;
  R.depth = 0; R.total_leaf_depth = 0; R.bump = 1; R.half = 0; R.nif = 0; R.expanded = 0; R.in_flight = 0;
  if (F.gum) { F.gum[t].num_sims_target = 0; fg_reset(F.gum[t]); }
}
// The reference flips its playout-cap and resign-playthrough coins with a thread_local std::default_random_engine seeded
// from std::random_device (play_manager.cc:261-262): unseedable, so there is nothing to replay. Here every slot has its
// OWN coin stream, separate from the stream that drives the search: at the deterministic corners (playout_cap_percent
// 0 or 1, resign_playthrough_percent 0 or 1) a slot is then still the reference bit for bit, in between only the
// coin values differ (tests/test_tafl_selfplay.py).
__device__ __forceinline__ bool sp_coin(SpSlot& G, float p) { return rng_uniform01(G.coin) < p; }
// the search budget of the seat to move (play_manager.cc:284-285)
__device__ __forceinline__ u32 sp_perm(const SpView& S, u32 g) { return S.n_perms > 1u ? g % S.n_perms : 0u; }
__device__ __forceinline__ u32 sp_goal(const SpView& S, const SpSlot& G, u32 g, u32 cp) {
  const u32 p = sp_perm(S, g);
  return G.capped ? S.seat_cap_visits[p][cp] : S.seat_visits[p][cp];
}
// MCTS::set_gumbel_num_sims on the tree of the seat to move — the full budget, or for a capped search the cap when
// fast_search_uses_gumbel and 0 (= PUCT for this search) otherwise — then under tree reuse the root temperature and
// fresh noise on a root that has been visited (play_manager.cc:523-553; also the first call of a run, 556-568)
__device__ __forceinline__ void sp_arm(const ForestView& F, const SpView& S, const SpSlot& G, u32 tn, u32 lane, bool reused_root) {
  if (lane == 0) {
    const u32 cp = tn & 1u, pm = sp_perm(S, tn >> 1);
    const u32 target = G.capped ? (S.fast_search_uses_gumbel ? S.seat_cap_visits[pm][cp] : 0u) : S.seat_visits[pm][cp];
    if (F.gum) { F.gum[tn].num_sims_target = target; fg_reset(F.gum[tn]); }
    if (reused_root) {
      ForestTree& R = F.trees[tn];
      if (R.n > 0 && R.blk != 0) {
        u32* pool = F.pool + (size_t)tn * F.words_per_tree;
        fr_apply_root_policy_temp(F, tn, pool, R.blk, R.k);
        if (FSEAT(F, tn).epsilon > 0.0f && !G.capped) {
          Pcg32 rng = FOREST_RNG(F, tn);
          fr_add_root_noise(F, tn, rng, pool, R.blk, R.k);
          FOREST_RNG(F, tn) = rng;
        }
      }
    }
  }
  __syncwarp();
}

template <int GAME>
__global__ void k_sp_init(const AZ_GC_F ForestView F, const AZ_GC_F SpView S, unsigned long long seed) {
  for (u32 g = GLOBAL_TID; g < S.n_games; g += GLOBAL_NT) {
    pcg32_seed(F.trees[2u * g].rng, seed + g);  // slot g == a reference run after MCTS::seed_thread_rng(seed + g)
    SpSlot& G = S.slots[g];
    G.active = 1;
    G.games_started = 1;
    pcg32_seed_stream(G.coin, seed + g, 0x5EEDC01ull);
    {  // base_gs_->copy() + randomize_start() for the first game of the slot (Unified variant mix only)
      const int v = FGame<GAME>::pick_variant(F, G.coin);
      if (v >= 0) { FGame<GAME>::init(F, 2u * g, v); FGame<GAME>::init(F, 2u * g + 1u, v); }
    }
    // game.initialized = true; the first playout-cap coin; set_gumbel_num_sims on the first seat's tree (play_manager.cc:556-568)
    G.capped = (S.playout_cap && sp_coin(G, S.playout_cap_percent)) ? 1u : 0u;
    const u32 cp0 = 0u, t0 = 2u * g + cp0;  // player 0 opens every game of this engine (attackers / Star Gambit's P0)
    const u32 pm0 = sp_perm(S, g);
    const u32 target = G.capped ? (S.fast_search_uses_gumbel ? S.seat_cap_visits[pm0][cp0] : 0u) : S.seat_visits[pm0][cp0];
    if (F.gum) { F.gum[t0].num_sims_target = target; fg_reset(F.gum[t0]); }
  }
}

// the hot kernel: `n_sims` x (MCTS::find_leaf, dumb_eval, MCTS::process_result) on the tree of the seat to move
template <int GAME>
__global__ void __launch_bounds__(128, GAME == B2AZ_FOREST_SG ? 4 : B2AZ_FOREST_MINB) k_sp_search(const AZ_GC_F ForestView F, const AZ_GC_F SpView S, u32 n_sims) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 g = GLOBAL_TID >> 5; g < S.n_games; g += GLOBAL_NT >> 5) {
    if (!S.slots[g].active) continue;
    const u32 cp = FGame<GAME>::root_player(F, 2u * g), t = 2u * g + cp;
    const bool noise = FSEAT(F, t).epsilon > 0.0f && !S.slots[g].capped;  // seat_epsilon > 0 && !capped
    const u32 goal = sp_goal(S, S.slots[g], g, cp), have = F.trees[t].depth;
    const u32 todo = have < goal ? (goal - have < n_sims ? goal - have : n_sims) : 0u;  // this slot's own budget
    for (u32 i = 0; i < todo; ++i) {
      forest_find_leaf<GAME, false>(F, t, sm[wib], lane, false, F.trees[t].leaf, nullptr);
      forest_process_result<GAME, true, false>(F, t, nullptr, nullptr, lane, noise, F.trees[t].leaf);
    }
    if (lane == 0) S.slots[g].simulations += todo;
  }
}
// The same search with the 16 warps of a 512-thread CTA in lock step (forest_find_leaf LOCK): every tree level is one
// CTA-wide round, the expansions run together, the backups run together. Slots that need fewer simulations (capped
// searches, retired slots) sit the rounds out.
template <int GAME>
__global__ void __launch_bounds__(512, GAME == B2AZ_FOREST_SG ? 1 : 2) k_sp_search_lock(const AZ_GC_F ForestView F, const AZ_GC_F SpView S, u32 n_sims) {
  __shared__ ForestSmem<GAME> sm[16];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 g0 = blockIdx.x * 16u; g0 < S.n_games; g0 += gridDim.x * 16u) {
    const u32 g = g0 + wib;
    const bool has = g < S.n_games && S.slots[g].active != 0;
    u32 t = 0, todo = 0;
    bool noise = false;
    if (has) {
      const u32 cp = FGame<GAME>::root_player(F, 2u * g);
      t = 2u * g + cp;
      noise = FSEAT(F, t).epsilon > 0.0f && !S.slots[g].capped;
      const u32 goal = sp_goal(S, S.slots[g], g, cp), have = F.trees[t].depth;
      todo = have < goal ? (goal - have < n_sims ? goal - have : n_sims) : 0u;
    }
    for (u32 i = 0;; ++i) {
      const bool live = has && i < todo;
      if (!__syncthreads_or(live ? 1 : 0)) break;
      forest_find_leaf<GAME, false, true>(F, t, sm[wib], lane, false, F.trees[t].leaf, nullptr, live);
      __syncthreads();
      if (live) forest_process_result<GAME, true, false>(F, t, nullptr, nullptr, lane, noise, F.trees[t].leaf);
    }
    if (has && lane == 0) S.slots[g].simulations += todo;
  }
}
__device__ __forceinline__ int sp_cache_lookup(const SpCache& c, u64 key, u32 lane) {  // entry index or -1
  const u32 set = (u32)(key >> 17) & (c.sets - 1u);
  const u64 k = lane < 4u ? c.keys[(size_t)set * 4u + lane] : 0ULL;
  const u32 m = __ballot_sync(0xFFFFFFFFu, lane < 4u && k == key);
  if (!m) return -1;
  const u32 e = set * 4u + (u32)(__ffs((int)m) - 1);
  if (lane == 0) { const u8 f = c.freq[e]; if (f < 3) c.freq[e] = (u8)(f + 1); }
  return (int)e;
}
// One way is written by at most one slot per launch: the writer first wins the way's `stamp` word for this launch's epoch,
// and a way stamped with the current epoch is never chosen as a victim — otherwise a second slot could evict the
// freshly claimed way while its 10 KB row is still being written and leave a row mixed from two evaluations.
__device__ __forceinline__ void sp_cache_insert(const SpCache& c, u64 key, const float* v, const float* pi, u32 A, u32 lane) {
  const u32 set = (u32)(key >> 17) & (c.sets - 1u);
  const size_t i0 = (size_t)set * 4u;
  const u64 k = lane < 4u ? c.keys[i0 + lane] : ~0ULL;
  const u32 f = lane < 4u ? c.freq[i0 + lane] : 255u;
  const u32 st = lane < 4u ? c.stamp[i0 + lane] : c.epoch;
  if (__ballot_sync(0xFFFFFFFFu, lane < 4u && k == key)) return;  // the first value wins
  const bool cand = lane < 4u && st != c.epoch;
  // an empty way first, else the lowest frequency (lowest way on ties), among the ways not written in this launch
  u32 best = cand ? (((k == 0ULL ? 0u : 1u + f) << 2) | lane) : 0xFFFFFFFFu;
#pragma unroll
  for (int o = 2; o > 0; o >>= 1) {
    const u32 ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
    best = ob < best ? ob : best;
  }
  best = __shfl_sync(0xFFFFFFFFu, best, 0);
  if (best == 0xFFFFFFFFu) return;  // every way of the set was written in this launch: a cache may forget
  const u32 way = best & 3u;
  const u64 old = __shfl_sync(0xFFFFFFFFu, k, (int)way);
  const u32 old_st = __shfl_sync(0xFFFFFFFFu, st, (int)way);
  const u32 e = set * 4u + way;
  u32 ok = 0;
  if (lane == 0) ok = atomicCAS(&c.stamp[e], old_st, c.epoch) == old_st ? 1u : 0u;
  ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
  if (!ok) return;  // another slot took the way in this launch
  if (lane < 4u && old != 0ULL && lane != way && f > 0u && f != 255u) c.freq[i0 + lane] = (u8)(f - 1u);  // ageing
  if (lane == 0) {
    c.keys[e] = key;
    c.freq[e] = 0;
    atomicAdd(&c.ctr[2], 1ULL);
    if (old != 0ULL) atomicAdd(&c.ctr[3], 1ULL);
  }
  if (lane < 3u) c.v[(size_t)e * 3u + lane] = v[lane];
  for (u32 i = lane; i < A; i += 32u) c.pi[(size_t)e * A + i] = pi[i];
}

// the evaluator-in-the-middle form of the same step (EvalType::NN): leaves' canonical planes out, (v, pi) rows in;
// row g of both belongs to slot g. With the position cache a slot keeps simulating while its leaves hit
// (play_manager.cc:589-594) and stops at its first miss (wait[g] = 1) or when its search is complete.
template <int GAME>
__global__ void __launch_bounds__(128) k_sp_find_leaf(const AZ_GC_F ForestView F, const AZ_GC_F SpView S, float* canon) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 g = GLOBAL_TID >> 5; g < S.n_games; g += GLOBAL_NT >> 5) {
    if (!S.slots[g].active) {
      if (S.wait && lane == 0) S.wait[g] = 0;
      continue;
    }
    const u32 cp = FGame<GAME>::root_player(F, 2u * g), t = 2u * g + cp;
    float* row = canon + (size_t)g * FGame<GAME>::canon(F);
    // the model group that searches for this seat under the slot's seat permutation (play_manager.cc:577)
    const u32 group = S.perm_group[sp_perm(S, g)][cp];
    if (S.leaf_group && lane == 0) S.leaf_group[g] = (u8)group;
    const bool random_group = ((S.random_groups >> group) & 1u) != 0;
    if (!S.cache.sets && !random_group) {
      forest_find_leaf<GAME, false>(F, t, sm[wib], lane, true, F.trees[t].leaf, row);
      if (S.wait && lane == 0) S.wait[g] = 1;
      continue;
    }
    if (random_group) {  // eval_types_[group] == RANDOM: dumb_eval inline, nothing for the evaluator (play_manager.cc:578-587)
      const u32 goal = sp_goal(S, S.slots[g], g, cp), have = F.trees[t].depth;
      const u32 todo = have < goal ? goal - have : 0u;
      const bool noise = FSEAT(F, t).epsilon > 0.0f && !S.slots[g].capped;
      for (u32 i = 0; i < todo; ++i) {
        forest_find_leaf<GAME, false>(F, t, sm[wib], lane, false, F.trees[t].leaf, nullptr);
        forest_process_result<GAME, true, false>(F, t, nullptr, nullptr, lane, noise, F.trees[t].leaf);
      }
      if (lane == 0) { S.slots[g].simulations += todo; S.wait[g] = 0; }
      __syncwarp();
      continue;
    }
    const bool noise = FSEAT(F, t).epsilon > 0.0f && !S.slots[g].capped;
    const u32 goal = sp_goal(S, S.slots[g], g, cp);
    u32 pending = 0;
    for (u32 guard = 0; guard <= goal; ++guard) {
      if (F.trees[t].depth >= goal) break;  // the search is complete: k_sp_move plays the move
      u64 key = 0;
      forest_find_leaf<GAME, false>(F, t, sm[wib], lane, true, F.trees[t].leaf, row, true, &key);
      if (group) key ^= 0x9E3779B97F4A7C15ull;  // one table for both model groups (the reference keeps a cache per group)
      if (key == 0ULL) key = 1ULL;
      const int e = sp_cache_lookup(S.cache, key, lane);
      if (e < 0) {
        pending = 1;
        if (lane == 0) { S.leaf_key[g] = key; atomicAdd(&S.cache.ctr[1], 1ULL); }
        break;
      }
      if (lane == 0) atomicAdd(&S.cache.ctr[0], 1ULL);
      forest_process_result<GAME, false, false>(F, t, S.cache.v, S.cache.pi, lane, noise, F.trees[t].leaf, (u32)e);
      if (lane == 0) S.slots[g].simulations += 1;
      __syncwarp();
    }
    if (lane == 0) S.wait[g] = pending;
    __syncwarp();
  }
}
template <int GAME>
__global__ void __launch_bounds__(128) k_sp_process_result(const AZ_GC_F ForestView F, const AZ_GC_F SpView S, const float* ev_v, const float* ev_pi) {
  const u32 lane = threadIdx.x & 31u;
  for (u32 g = GLOBAL_TID >> 5; g < S.n_games; g += GLOBAL_NT >> 5) {
    if (!S.slots[g].active) continue;
    if (S.wait && !S.wait[g]) continue;  // (cache) nothing of this slot waits for the evaluator
    const u32 t = 2u * g + FGame<GAME>::root_player(F, 2u * g);
    forest_process_result<GAME, false, false>(F, t, ev_v, ev_pi, lane, FSEAT(F, t).epsilon > 0.0f && !S.slots[g].capped, F.trees[t].leaf,
                                              /*row=*/g);  // row g of the evaluator's output belongs to slot g
    if (S.cache.sets) {  // update_inferences' insert_many (play_manager.cc:619-642)
      const u32 A = FGame<GAME>::actions(F);
      sp_cache_insert(S.cache, S.leaf_key[g], ev_v + (size_t)g * 3, ev_pi + (size_t)g * A, A, lane);
      if (lane == 0) S.wait[g] = 0;
    }
    if (lane == 0) S.slots[g].simulations += 1;
  }
}

// "Actually play a move" (play_manager.cc:283-553) for every slot whose search has reached its visit count
#ifndef B2AZ_SP_MOVE_MINB
#define B2AZ_SP_MOVE_MINB 8  /* 64 registers, 32 warps per SM: the lane-0 stretches of the move step are latency bound (measured: 1 / 4 / 6 / 8 -> Brandubh 112.6 / 112.7 / 115.7 / 118.2 M sims/s, OpenTafl 46.8 / 46.9 / 46.5 / 49.1) */
#endif
template <int GAME>
__global__ void __launch_bounds__(128, GAME == B2AZ_FOREST_SG ? 4 : B2AZ_SP_MOVE_MINB) k_sp_move(const AZ_GC_F ForestView F, const AZ_GC_F SpView S) {
  typedef FGame<GAME> GM;
  const u32 A = GM::actions(F), CANON = GM::canon(F);
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 g = GLOBAL_TID >> 5; g < S.n_games; g += GLOBAL_NT >> 5) {
    SpSlot& G = S.slots[g];
    if (!G.active) continue;
    const u32 cp = FGame<GAME>::root_player(F, 2u * g), t = 2u * g + cp;
    ForestTree& R = F.trees[t];
    if (R.depth < sp_goal(S, G, g, cp)) continue;  // mcts.depth() >= goal_depth
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    const bool capped = G.capped != 0;
    // temperature schedule (play_manager.cc:285-302)
    float temp = S.start_temp;
    float half_life = S.half_life;
    {  // temp_decay_half_life_by_variant (play_manager.cc:289-296)
      const int vid = GM::root_variant(F, 2u * g);
      if (S.n_variant_half_life && vid >= 0 && vid < (int)S.n_variant_half_life) half_life = S.variant_half_life[vid];
    }
    if (half_life != 0.0f) {
      const float lambda = fdiv(0.693f, half_life);
      temp = fsub(temp, S.final_temp);
      temp = fmul(temp, az_expf(fmul(-lambda, (float)GM::root_turn(F, 2u * g))));
      temp = fadd(temp, S.final_temp);
    }
    // resign_percent (play_manager.cc:305-333): MCTS::root_value (mcts.h:78-100) against 1 - resign_percent
    u32 resign_term = 0;  // 1 + index of the score that would be awarded
    if (S.resign_percent > 0.0f && !G.playthrough) {
      if (lane == 0) {
        float q = 0.0f, d = 0.0f;
        bool found = false;
        const u32 b = R.blk, k = b ? R.k : 0u;
        for (u32 j = 0; j < k; ++j) {
          const float qj = u2f(pool[fb_q(b, k) + j]);
          if (pool[fb_n(b, k) + j] > 0 && qj > q) { q = qj; d = u2f(pool[fb_d(b, k) + j]); found = true; }
        }
        if (!found && R.n > 0) { q = R.v; d = R.d; }
        const float w = fsub(q, fdiv(d, 2.0f));
        const float l = (float)dsub(dsub(1.0, (double)w), (double)d);
        const double resign_val = dsub(1.0, (double)S.resign_percent);
        u32 tm = 0;
        if ((double)w > resign_val) tm = cp + 1u;
        else if ((double)l > resign_val) tm = ((cp + 1u) % 2u) + 1u;
        else if ((double)d > resign_val) tm = 3u;
        if (tm != 0) {
          if (sp_coin(G, S.resign_playthrough_percent)) G.playthrough = 1;
          else resign_term = tm;
        }
      }
      resign_term = __shfl_sync(0xFFFFFFFFu, resign_term, 0);
    }
    // per-seat opt-in resign (play_manager.cc:335-366): the seat's expected score W - L at or below its threshold for
    // `consecutive` own moves in a row
    {
      const u32 pm = sp_perm(S, g);
      const float seat_thresh = S.seat_resign_threshold[pm][cp];
      if (resign_term == 0 && !G.playthrough && seat_thresh > -2.0f) {
        if (lane == 0) {
          float q = 0.0f, d = 0.0f;
          bool found = false;
          const u32 b = R.blk, k = b ? R.k : 0u;
          for (u32 j = 0; j < k; ++j) {
            const float qj = u2f(pool[fb_q(b, k) + j]);
            if (pool[fb_n(b, k) + j] > 0 && qj > q) { q = qj; d = u2f(pool[fb_d(b, k) + j]); found = true; }
          }
          if (!found && R.n > 0) { q = R.v; d = R.d; }
          const float w = fsub(q, fdiv(d, 2.0f));
          const float l = (float)dsub(dsub(1.0, (double)w), (double)d);
          const float v_self = fsub(w, l);
          if (v_self <= seat_thresh) G.resign_streak[cp] = (u16)(G.resign_streak[cp] < 0xFFFFu ? G.resign_streak[cp] + 1u : 0xFFFFu);
          else G.resign_streak[cp] = 0;
          const u32 need = S.seat_resign_consecutive[pm][cp] > 1u ? S.seat_resign_consecutive[pm][cp] : 1u;
          if (G.resign_streak[cp] >= need) resign_term = ((cp + 1u) % 2u) + 1u;  // the opponent wins
        }
        resign_term = __shfl_sync(0xFFFFFFFFu, resign_term, 0);
      }
    }
    // acting rule (play_manager.cc:367-406): Gumbel's final action only after a full search
    float* act = S.scratch_pi + (size_t)g * A;
    u32 chosen = 0xFFFFFFFFu;
    if (FSEAT(F, t).gumbel_enabled && !capped) {
      if (lane == 0) chosen = fg_final_action(F, t, R, F.gum[t], pool);
      chosen = __shfl_sync(0xFFFFFFFFu, chosen, 0);
      if (chosen == 0xFFFFFFFFu) {  // the search never initialised: pick_move(probs(0)) (mcts.cc:379-381)
        forest_probs<GAME>(F, t, 0.0f, act, &G.picked, 1u, 0u, lane);
        chosen = G.picked;
      }
    } else {
      forest_probs<GAME>(F, t, temp, act, &G.picked, 1u, 0u, lane);
      chosen = G.picked;
    }
    // training sample (play_manager.cc:407-424): full searches only
    if (S.history_enabled && !capped) {
      // (the staging area of a slot is a ring: a game with more full searches than `stage_rows` — only Star Gambit's
      // action count has no small bound — keeps its most recent stage_rows samples)
      const size_t row = (size_t)g * S.stage_rows + G.pending % S.stage_rows;
      {
        typename GM::Pos pos;  // GameState::canonicalized() of the root position
        GM::open(F, t, lane, sm[wib], pos);
        GM::stage(pos, sm[wib], S.st_canon + row * S.st_stride, lane);
      }
      float* pi = S.st_pi + row * A;
      if (S.gumbel_targets) {  // params_.gumbel_enabled — the GLOBAL flag, whatever the seat's own search is (play_manager.cc:412-419)
        for (u32 m = lane; m < A; m += 32u) pi[m] = 0.0f;
        __syncwarp();
        if (lane == 0) fg_improved_policy(F, t, R, pool, pi);
        __syncwarp();
      } else {
        forest_probs<GAME>(F, t, 1.0f, pi, nullptr, 0u, (S.policy_target_pruning && FSEAT(F, t).epsilon > 0.0f) ? 1u : 0u, lane);
      }
      if (lane == 0) S.st_player[row] = (u8)cp;
    }
    if (lane == 0) {
      if (S.history_enabled && !capped) G.pending += 1;
      // metrics (play_manager.cc:436-446; MCTS::avg_leaf_depth mcts.h:112, normalized_root_entropy mcts.cc:737-750)
      const float ald = R.depth == 0 ? 0.0f : fdiv((float)R.total_leaf_depth, (float)R.depth);
      float ent = 0.0f;
      const u32 b = R.blk, k = b ? R.k : 0u;
      if (k > 1 && R.n > 1) {
        const float log_k = az_logf((float)k), total_n = (float)R.n;
        float e = 0.0f;
        for (u32 j = 0; j < k; ++j) {
          const u32 nj = pool[fb_n(b, k) + j];
          if (nj > 0) {
            const float p = fdiv((float)nj, total_n);
            e = fsub(e, fmul(p, az_logf(p)));
          }
        }
        ent = fdiv(e, log_k);
      }
      if (!capped) {
        G.g_leaf_depth += (double)ald;
        G.g_entropy += (double)ent;
        G.full_move_count += 1;
      } else {
        G.g_fast_leaf_depth += (double)ald;
        G.g_fast_entropy += (double)ent;
        G.fast_move_count += 1;
      }
      G.g_valid_moves += (double)k;
      G.move_count += 1;
    }
    __syncwarp();
    // for (auto& m : game.mcts) m.update_root(*game.gs, chosen_m); game.gs->play_move(chosen_m)  — seat 0's tree first:
    // an unexpanded root draws its child shuffle from the shared stream
    forest_update_root<GAME>(F, 2u * g, chosen, sm[wib], lane);
    __syncwarp();
    forest_update_root<GAME>(F, 2u * g + 1u, chosen, sm[wib], lane);
    __syncwarp();
    u32 term = GM::root_terminal(F, 2u * g);
    if (term == 0 && resign_term != 0) term = resign_term;  // play_manager.cc:440-444
    else resign_term = 0;
    if (term != 0) {
      const float s0 = term == 1 ? 1.0f : 0.0f, s1 = term == 2 ? 1.0f : 0.0f, sd = term == 3 ? 1.0f : 0.0f;
      if (S.history_enabled) {
        const u32 cnt = G.pending < S.stage_rows ? G.pending : S.stage_rows;
        u32 base = 0;
        if (lane == 0) base = atomicAdd(S.out_count, cnt);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (u32 i = 0; i < cnt; ++i) {  // partial_history.back() first
          const u32 dst = base + i;
          if (dst >= S.out_cap) { if (lane == 0) G.error |= 1u; break; }
          const size_t src = (size_t)g * S.stage_rows + (G.pending - 1u - i) % S.stage_rows;
          GM::unstage(F, S.st_canon + src * S.st_stride, sm[wib], S.out_canon + (size_t)dst * CANON, lane);
          for (u32 e = lane; e < A; e += 32u) S.out_pi[(size_t)dst * A + e] = S.st_pi[src * A + e];
          if (lane == 0) {
            // relative_values games store the outcome in the frame of the sample's mover (absolute_to_relative,
            // play_manager.cc:451-455)
            const bool sw = F.relative_values && S.st_player[src] == 1;
            S.out_v[(size_t)dst * 3 + 0] = sw ? s1 : s0; S.out_v[(size_t)dst * 3 + 1] = sw ? s0 : s1; S.out_v[(size_t)dst * 3 + 2] = sd;
            S.out_slot[dst] = g;
          }
        }
      }
      bool retire = false;
      if (lane == 0) {
        G.pending = 0;
        G.scores[0] = fadd(G.scores[0], s0); G.scores[1] = fadd(G.scores[1], s1); G.scores[2] = fadd(G.scores[2], sd);
        if (resign_term != 0) {
          G.resign_scores[0] = fadd(G.resign_scores[0], s0); G.resign_scores[1] = fadd(G.resign_scores[1], s1);
          G.resign_scores[2] = fadd(G.resign_scores[2], sd);
        }
        G.games_completed += 1;
        G.game_length += GM::root_turn(F, 2u * g);
        const int vid = GM::root_variant(F, 2u * g);
        if (S.variants && vid >= 0 && vid < 4) {  // play_manager.cc:468-484
          SpVariantAcc& V = S.variants[(size_t)g * 4u + (u32)vid];
          V.scores[0] = fadd(V.scores[0], s0); V.scores[1] = fadd(V.scores[1], s1); V.scores[2] = fadd(V.scores[2], sd);
          V.games_completed += 1;
          V.game_length += GM::root_turn(F, 2u * g);
          V.leaf_depth += G.g_leaf_depth; V.entropy += G.g_entropy; V.valid_moves += G.g_valid_moves;
          V.fast_leaf_depth += G.g_fast_leaf_depth; V.fast_entropy += G.g_fast_entropy;
          V.total_move_count += G.move_count; V.full_move_count += G.full_move_count; V.fast_move_count += G.fast_move_count;
        }
        G.leaf_depth += G.g_leaf_depth; G.entropy += G.g_entropy; G.valid_moves += G.g_valid_moves;
        G.total_move_count += G.move_count; G.total_full_move_count += G.full_move_count;
        G.fast_leaf_depth += G.g_fast_leaf_depth; G.fast_entropy += G.g_fast_entropy; G.total_fast_move_count += G.fast_move_count;
        G.g_leaf_depth = 0; G.g_entropy = 0; G.g_valid_moves = 0; G.move_count = 0; G.full_move_count = 0;
        G.g_fast_leaf_depth = 0; G.g_fast_entropy = 0; G.fast_move_count = 0;
        if (G.games_started >= S.games_per_slot) {
          G.active = 0;  // `continue`: the slot is not pushed back
          retire = true;
        } else {
          G.games_started += 1;
          const int nv = GM::pick_variant(F, G.coin);  // game.gs->randomize_start()
          for (u32 j = 0; j < 2u; ++j) {  // game.gs = base_gs_->copy(); fresh MCTS per seat
            GM::init(F, 2u * g + j, nv);
            sp_reset_search(F, 2u * g + j);
          }
        }
      }
      retire = __shfl_sync(0xFFFFFFFFu, retire ? 1u : 0u, 0) != 0;
      __syncwarp();
      if (retire) continue;
    }
    // a move has been played: the next search's playout cap (play_manager.cc:523-524; `&&` short-circuits the coin)
    if (lane == 0) G.capped = (S.playout_cap && sp_coin(G, S.playout_cap_percent)) ? 1u : 0u;
    __syncwarp();
    const u32 tn = 2u * g + GM::root_player(F, 2u * g);
    if (!S.tree_reuse) {
      // set_gumbel_num_sims happens BEFORE the trees are replaced by fresh MCTS objects (play_manager.cc:531-545), so
      // without tree reuse the new objects have no simulation target (the reference's behaviour, reproduced)
      if (lane == 0) { sp_reset_search(F, 2u * g); sp_reset_search(F, 2u * g + 1u); }
      __syncwarp();
    } else {
      sp_arm(F, S, G, tn, lane, true);
    }
  }
}
// out[0] = active slots, out[1] = OR of every tree's sticky error bits (a search on a full slab is a search on a wrong
// tree: the calls that synchronise report it instead of returning 0)
__global__ void k_sp_count_active(const AZ_GC_F ForestView F, const AZ_GC_F SpView S, u32* out) {
  u32 c = 0, err = 0;
  for (u32 g = GLOBAL_TID; g < S.n_games; g += GLOBAL_NT) {
    c += S.slots[g].active ? 1u : 0u;
    err |= F.trees[2u * g].error | F.trees[2u * g + 1u].error | (S.slots[g].error ? 0x100u : 0u);
  }
  if (c) atomicAdd(out, c);
  if (err) atomicOr(out + 1, err);
}
#endif  // !B2AZ_HOST_EMU

}  // namespace b2az

// which search kernel: B2AZ_SP_LOCKSTEP=0 / 1 overrides; default = lock step for Star Gambit (measured)
static inline bool sp_lockstep(uint32_t game) {
  static const int env = [] { const char* e = getenv("B2AZ_SP_LOCKSTEP"); return e ? atoi(e) : -1; }();
  return env >= 0 ? env != 0 : game >= 10u;
}

struct b2az_tafl_selfplay {
  b2az_forest* forest = nullptr;
  b2az::SpView view;
  uint32_t* active_dev = nullptr;
  float *ev_v = nullptr, *ev_pi = nullptr, *leaf_canon = nullptr;
  uint32_t hist_head = 0;  // rows of the sample ring already handed out (the ring restarts once it has been emptied)
  std::vector<float> h_canon, h_v, h_pi;  // host staging of the reference-API flavour (leaf_batch_host / submit_eval_host)
  std::vector<b2az::SpSlot> h_slots;
  std::vector<uint32_t> h_tree_err, h_wait;
  std::vector<uint8_t> h_group_all, h_group;  // model group per slot / per row of the last leaf_batch_host
};

extern "C" {

int b2az_tafl_selfplay_destroy(b2az_tafl_selfplay* sp) {
  using namespace b2az;
  if (!sp) return 0;
  dev_free(sp->view.slots); dev_free(sp->view.st_canon); dev_free(sp->view.st_pi); dev_free(sp->view.st_player);
  dev_free(sp->view.scratch_pi); dev_free(sp->view.out_canon); dev_free(sp->view.out_v); dev_free(sp->view.out_pi);
  dev_free(sp->view.out_slot); dev_free(sp->view.out_count); dev_free(sp->active_dev);
  dev_free(sp->view.cache.keys); dev_free(sp->view.cache.freq); dev_free(sp->view.cache.stamp); dev_free(sp->view.cache.v); dev_free(sp->view.cache.pi);
  dev_free(sp->view.cache.ctr); dev_free(sp->view.wait); dev_free(sp->view.leaf_key); dev_free(sp->view.leaf_group); dev_free(sp->view.variants);
  dev_free(sp->ev_v); dev_free(sp->ev_pi); dev_free(sp->leaf_canon);
  b2az_forest_destroy(sp->forest);
  delete sp;
  return 0;
}

int b2az_tafl_selfplay_create(const b2az_tafl_selfplay_params* p, int device, b2az_tafl_selfplay** out) {
  using namespace b2az;
  if (!p || !out) return fail(B2AZ_EINVAL, "null argument");
  if (p->n_games == 0 || p->n_games > 0x7FFFFFFFu / 2u) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: bad n_games");
  if (p->games_per_slot == 0 || (p->visits == 0 && (p->seat_visits[0] == 0 || p->seat_visits[1] == 0)))
    return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: games_per_slot and visits must be positive");
  if (p->forest.max_in_flight != 0) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: PlayManager runs one leaf per game (max_in_flight must be 0)");
  if (p->n_seat_perms > (uint32_t)b2az::kSpMaxPerms) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: at most 8 seat permutations");
  if (p->n_seat_perms > 1 && p->n_games % p->n_seat_perms != 0)
    return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: n_games must be a multiple of the number of seat permutations (slot g plays permutation g % n)");
  for (uint32_t pm = 0; pm < p->n_seat_perms; ++pm)
    if (p->seat_perms[pm][0] > 1 || p->seat_perms[pm][1] > 1) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: a model group index must be 0 or 1");
  b2az_forest_params fp = p->forest;
  fp.n_trees = 2u * p->n_games;
  if (p->has_seat_search) {  // the forest allocates its Dirichlet / Gumbel scratch from the globals: the union of the seats
    for (uint32_t pm = 0; pm < std::max(1u, p->n_seat_perms) && pm < (uint32_t)b2az::kSpMaxPerms; ++pm)
      for (int seat = 0; seat < 2; ++seat) {
        fp.epsilon = std::max(fp.epsilon, p->seat_epsilon[pm][seat]);
        if (p->seat_gumbel_enabled[pm][seat]) {
          if (!fp.gumbel_enabled) { fp.gumbel_m = p->seat_gumbel_m[pm][seat]; fp.gumbel_c_visit = p->seat_gumbel_c_visit[pm][seat]; fp.gumbel_c_scale = p->seat_gumbel_c_scale[pm][seat]; }
          fp.gumbel_enabled = 1;
          if (p->seat_gumbel_m[pm][seat] == 0) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay: seat_gumbel_m must be > 0");
        }
      }
  }
  if (fp.words_per_tree == 0) {
    // sized from the search: each half of a tree's slab holds the kept subtree plus one search's new nodes; an expanded
    // node is 1 + 8 k words. Budget 8 x visits nodes at a typical branching (Brandubh 64, 11x11 boards 200) per half —
    // a search that outgrows it is reported as B2AZ_ENOMEM by the calls that synchronise, never silently truncated.
    const uint64_t k_typ = p->forest.game == B2AZ_TAFL_BRANDUBH ? 64u : 200u;
    uint64_t vmax = std::max<uint64_t>(p->visits, std::max(p->seat_visits[0], p->seat_visits[1]));
    for (uint32_t pm = 0; pm < p->n_seat_perms; ++pm) vmax = std::max<uint64_t>(vmax, std::max(p->perm_seat_visits[pm][0], p->perm_seat_visits[pm][1]));
    const uint64_t want = 2ull * (1ull + 8ull * vmax * (1ull + 8ull * k_typ));
    fp.words_per_tree = (uint32_t)std::min<uint64_t>(want, 1ull << 26);
  }
  b2az_forest* f = nullptr;
  if (int rc = b2az_forest_create(&fp, device, &f)) return rc;
#ifdef B2AZ_HOST_EMU
  return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
#else
  b2az_tafl_selfplay* sp = new b2az_tafl_selfplay();
  sp->forest = f;
  f->view.rng_pair = 1;
  SpView& S = sp->view;
  memset(&S, 0, sizeof(S));
  S.n_games = p->n_games; S.games_per_slot = p->games_per_slot;
  S.n_perms = std::max(1u, p->n_seat_perms);
  S.visits = 0;
  bool several_groups = false;
  for (uint32_t pm = 0; pm < S.n_perms; ++pm)
    for (int seat = 0; seat < 2; ++seat) {
      const uint32_t sv = p->perm_seat_visits[pm][seat] ? p->perm_seat_visits[pm][seat] : p->seat_visits[seat];
      const uint32_t cv = p->perm_seat_cap_visits[pm][seat] ? p->perm_seat_cap_visits[pm][seat] : p->seat_cap_visits[seat];
      S.seat_visits[pm][seat] = sv ? sv : p->visits;
      S.seat_cap_visits[pm][seat] = cv ? cv : (p->playout_cap_depth ? p->playout_cap_depth : 25u);
      S.visits = std::max(S.visits, std::max(S.seat_visits[pm][seat], S.seat_cap_visits[pm][seat]));
      S.perm_group[pm][seat] = p->n_seat_perms ? p->seat_perms[pm][seat] : 0;
      several_groups |= S.perm_group[pm][seat] != 0;
    }
  S.random_groups = (p->group_random[0] ? 1u : 0u) | (p->group_random[1] ? 2u : 0u);
  for (uint32_t pm = 0; pm < (uint32_t)kSpMaxPerms; ++pm)
    for (int seat = 0; seat < 2; ++seat) {
      const bool on = p->has_seat_search && pm < S.n_perms;
      S.seat_resign_threshold[pm][seat] = on ? p->seat_resign_threshold[pm][seat] : -2.0f;
      S.seat_resign_consecutive[pm][seat] = on ? p->seat_resign_consecutive[pm][seat] : 1u;
    }
  S.gumbel_targets = p->forest.gumbel_enabled ? 1u : 0u;
  S.playout_cap = p->playout_cap_randomization ? 1u : 0u;
  S.fast_search_uses_gumbel = p->fast_search_uses_gumbel ? 1u : 0u;
  S.playout_cap_percent = p->playout_cap_percent;
  S.resign_percent = p->resign_percent; S.resign_playthrough_percent = p->resign_playthrough_percent;
  S.start_temp = p->start_temp; S.final_temp = p->final_temp; S.half_life = p->temp_decay_half_life;
  S.history_enabled = p->history_enabled ? 1u : 0u;
  S.policy_target_pruning = p->policy_target_pruning ? 1u : 0u;
  S.tree_reuse = p->tree_reuse ? 1u : 0u;
  S.n_variant_half_life = std::min(p->n_variant_half_life, 4u);
  for (int i = 0; i < 4; ++i) S.variant_half_life[i] = p->variant_half_life[i];
  S.out_cap = S.history_enabled ? (p->hist_capacity ? p->hist_capacity : p->n_games * fp.max_turns) : 1u;
  S.stage_rows = fp.max_turns;
  S.st_stride = f->view.sg_state ? (uint32_t)(sizeof(SGState) / 4u + 1u) : f->canon;
  const size_t G = p->n_games, MT = fp.max_turns, A = f->actions, C = f->canon;
  auto bail = [&](int rc) { b2az_tafl_selfplay_destroy(sp); return rc; };
  if (p->has_seat_search) {  // make_mcts(perm, seat): every seat's own search settings (play_manager.cc:602-617)
    ForestView& FV = f->view;
    SeatSearch sets[kFSeatSets];
    memset(sets, 0, sizeof(sets));
    for (uint32_t pm = 0; pm < S.n_perms; ++pm)
      for (int seat = 0; seat < 2; ++seat)
        sets[pm * 2 + seat] = SeatSearch{p->seat_epsilon[pm][seat], p->seat_root_temp[pm][seat], p->seat_gumbel_c_visit[pm][seat],
                                         p->seat_gumbel_c_scale[pm][seat], p->seat_gumbel_m[pm][seat],
                                         (u8)(p->seat_root_fpu_zero[pm][seat] ? 1 : 0), (u8)(p->seat_gumbel_enabled[pm][seat] ? 1 : 0),
                                         (u8)(p->seat_gumbel_enabled[pm][seat] && p->seat_gumbel_full[pm][seat] ? 1 : 0), 0};
    // tree t = 2 * slot + seat; slot plays permutation slot % n_perms
    std::vector<SeatSearch> h((size_t)FV.n_trees);
    for (uint32_t t = 0; t < FV.n_trees; ++t) h[t] = sets[((t >> 1) % S.n_perms) * 2u + (t & 1u)];
    if (cudaMemcpy(const_cast<SeatSearch*>(FV.seat), h.data(), h.size() * sizeof(SeatSearch), cudaMemcpyHostToDevice) != cudaSuccess)
      return bail(fail(B2AZ_ECUDA, "b2az_tafl_selfplay: seat table upload failed"));
  }
  if (int rc = dev_alloc(&S.slots, G)) return bail(rc);
  if (int rc = dev_alloc_raw(&S.scratch_pi, G * A)) return bail(rc);
  if (S.history_enabled) {
    if (int rc = dev_alloc_raw(&S.st_canon, G * MT * S.st_stride)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.st_pi, G * MT * A)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.st_player, G * MT)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.out_canon, (size_t)S.out_cap * C)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.out_v, (size_t)S.out_cap * 3)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.out_pi, (size_t)S.out_cap * A)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.out_slot, (size_t)S.out_cap)) return bail(rc);
  }
  if (p->cache_entries) {  // max_cache_size (one model group): sets of four, rounded up to a power of two
    uint32_t sets = 1;
    while (sets < 0x40000000u && (uint64_t)sets * 4u < p->cache_entries) sets <<= 1;
    S.cache.sets = sets;
    const size_t E = (size_t)sets * 4u;
    if (int rc = dev_alloc(&S.cache.keys, E)) return bail(rc);
    if (int rc = dev_alloc(&S.cache.freq, E)) return bail(rc);
    if (int rc = dev_alloc(&S.cache.stamp, E)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.cache.v, E * 3)) return bail(rc);
    if (int rc = dev_alloc_raw(&S.cache.pi, E * A)) return bail(rc);
    if (int rc = dev_alloc(&S.cache.ctr, 4)) return bail(rc);
    if (int rc = dev_alloc(&S.leaf_key, G)) return bail(rc);
  }
  // wait[g]: the slot's leaf waits for the evaluator (cache: its leaves may all have hit; an EvalType::RANDOM group next
  // to an NN one: its searches never wait)
  if (S.cache.sets || S.random_groups)
    if (int rc = dev_alloc(&S.wait, G)) return bail(rc);
  if (several_groups)
    if (int rc = dev_alloc(&S.leaf_group, G)) return bail(rc);
  if (fp.game >= 20u)
    if (int rc = dev_alloc(&S.variants, G * 4)) return bail(rc);
  if (int rc = dev_alloc(&S.out_count, 1)) return bail(rc);
  if (int rc = dev_alloc(&sp->active_dev, 2)) return bail(rc);
  if (fp.game == 24u) {
    float tot = 0.0f;
    for (int i = 0; i < 4; ++i) tot += p->variant_probs[i] > 0.0f ? p->variant_probs[i] : 0.0f;
    for (int i = 0; i < 4; ++i) f->view.sg_probs[i] = tot > 0.0f ? (p->variant_probs[i] > 0.0f ? p->variant_probs[i] : 0.0f) : 0.25f;
  }
  FOREST_DISPATCH(f, (k_sp_init<G_><<<148, 128>>>(f->view, S, p->forest.seed)));
  if (cudaGetLastError() != cudaSuccess) return bail(fail(B2AZ_ECUDA, "k_sp_init launch failed"));
  if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(B2AZ_ECUDA, "k_sp_init failed"));
  *out = sp;
  return 0;
#endif
}

#ifdef B2AZ_HOST_EMU
int b2az_tafl_selfplay_play(b2az_tafl_selfplay*, void*, uint32_t, uint32_t*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_find_leaf(b2az_tafl_selfplay*, void*, const float**) FOREST_NO_CUDA()
int b2az_tafl_selfplay_process_result(b2az_tafl_selfplay*, void*, const float*, const float*, int, uint32_t*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_drain_history(b2az_tafl_selfplay*, void*, uint32_t, float*, float*, float*, uint32_t*, uint32_t*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_slots(b2az_tafl_selfplay*, void*, b2az_tafl_selfplay_slot*, uint32_t*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_get_stats(b2az_tafl_selfplay*, void*, b2az_stats*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_leaf_batch_host(b2az_tafl_selfplay*, void*, uint32_t, float*, uint32_t*, uint32_t*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_submit_eval_host(b2az_tafl_selfplay*, void*, const uint32_t*, const float*, const float*, uint32_t) FOREST_NO_CUDA()
int b2az_tafl_selfplay_variant_stats(b2az_tafl_selfplay*, void*, b2az_variant_stats*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_leaf_groups_host(b2az_tafl_selfplay*, uint8_t*, uint32_t) FOREST_NO_CUDA()
int b2az_tafl_selfplay_root_state(b2az_tafl_selfplay*, uint32_t, void*, uint32_t, void*, uint32_t, uint32_t*) FOREST_NO_CUDA()
int b2az_tafl_selfplay_perm_stats(b2az_tafl_selfplay*, void*, b2az_perm_stats*, uint32_t*) FOREST_NO_CUDA()
#else
#define SP_CTAS(sp) std::max(1u, std::min(((sp)->view.n_games + 3u) / 4u, 148u * 8u))
static int sp_active(b2az_tafl_selfplay* sp, cudaStream_t s, uint32_t* active_out) {
  using namespace b2az;
  if (!active_out) return 0;
  uint32_t host[2] = {0, 0};
  CUDA_TRY(cudaMemsetAsync(sp->active_dev, 0, 8, s));
  k_sp_count_active<<<148, 128, 0, s>>>(sp->forest->view, sp->view, sp->active_dev);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(host, sp->active_dev, 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  *active_out = host[0];
  if (host[1] & (1u | 4u | 16u))
    return fail(B2AZ_ENOMEM, "tafl self-play: a tree's node slab is full (the search went on on a truncated tree): raise "
                             "b2az_forest_params.words_per_tree");
  if (host[1] & 0x100u) return fail(B2AZ_ENOMEM, "tafl self-play: the training-sample ring is full: drain more often or raise hist_capacity");
  if (host[1] & 2u) return fail(B2AZ_ESTATE, "tafl self-play: selection path longer than the path buffer");
  if (host[1] & 8u) return fail(B2AZ_EMOVE, "tafl self-play: update_root could not find the move");
  return 0;
}
int b2az_tafl_selfplay_play(b2az_tafl_selfplay* sp, void* stream, uint32_t n_moves, uint32_t* active_out) {
  using namespace b2az;
  if (!sp) return fail(B2AZ_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  b2az_forest* f = sp->forest;
  for (uint32_t m = 0; m < n_moves; ++m) {
    if (sp_lockstep(f->view.game)) {
      const unsigned ctas = std::max(1u, std::min((sp->view.n_games + 15u) / 16u, 148u * 2u));
      FOREST_DISPATCH(f, (k_sp_search_lock<G_><<<ctas, 512, 0, s>>>(f->view, sp->view, sp->view.visits)));
    } else {
      FOREST_DISPATCH(f, (k_sp_search<G_><<<SP_CTAS(sp), 128, 0, s>>>(f->view, sp->view, sp->view.visits)));
    }
    FOREST_DISPATCH(f, (k_sp_move<G_><<<SP_CTAS(sp), 128, 0, s>>>(f->view, sp->view)));
  }
  CUDA_TRY(cudaGetLastError());
  return sp_active(sp, s, active_out);
}
int b2az_tafl_selfplay_find_leaf(b2az_tafl_selfplay* sp, void* stream, const float** canon_dev) {
  using namespace b2az;
  if (!sp || !canon_dev) return fail(B2AZ_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  b2az_forest* f = sp->forest;
  if (!sp->leaf_canon)
    if (int rc = dev_alloc(&sp->leaf_canon, (size_t)sp->view.n_games * f->canon)) return rc;
  FOREST_DISPATCH(f, (k_sp_find_leaf<G_><<<SP_CTAS(sp), 128, 0, s>>>(f->view, sp->view, sp->leaf_canon)));
  CUDA_TRY(cudaGetLastError());
  *canon_dev = sp->leaf_canon;
  return 0;
}
int b2az_tafl_selfplay_process_result(b2az_tafl_selfplay* sp, void* stream, const float* v, const float* pi, int host_pointers,
                                      uint32_t* active_out) {
  using namespace b2az;
  if (!sp || !v || !pi) return fail(B2AZ_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  b2az_forest* f = sp->forest;
  const size_t G = sp->view.n_games;
  if (host_pointers) {
    if (!sp->ev_v) {
      if (int rc = dev_alloc(&sp->ev_v, G * 3)) return rc;
      if (int rc = dev_alloc(&sp->ev_pi, G * f->actions)) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(sp->ev_v, v, G * 3 * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(sp->ev_pi, pi, G * f->actions * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));  // b2az.h: calls taking host buffers synchronise before returning (pinned buffers!)
    v = sp->ev_v; pi = sp->ev_pi;
  }
  ++sp->view.cache.epoch;  // one epoch per insert launch (sp_cache_insert)
  FOREST_DISPATCH(f, (k_sp_process_result<G_><<<SP_CTAS(sp), 128, 0, s>>>(f->view, sp->view, v, pi)));
  FOREST_DISPATCH(f, (k_sp_move<G_><<<SP_CTAS(sp), 128, 0, s>>>(f->view, sp->view)));
  CUDA_TRY(cudaGetLastError());
  return sp_active(sp, s, active_out);
}
int b2az_tafl_selfplay_drain_history(b2az_tafl_selfplay* sp, void* stream, uint32_t max_rows, float* canon_host, float* v_host,
                                     float* pi_host, uint32_t* slot_host, uint32_t* n_out) {
  using namespace b2az;
  if (!sp || !n_out) return fail(B2AZ_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t count = 0;
  CUDA_TRY(cudaMemcpyAsync(&count, sp->view.out_count, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  count = std::min(count, sp->view.out_cap);
  if (!sp->view.history_enabled) count = 0;
  const uint32_t head = std::min(sp->hist_head, count);
  const uint32_t n = std::min(count - head, max_rows);  // oldest rows first (history_ is a FIFO)
  const size_t A = sp->forest->actions, C = sp->forest->canon;
  if (n) {
    if (canon_host) CUDA_TRY(cudaMemcpyAsync(canon_host, sp->view.out_canon + (size_t)head * C, (size_t)n * C * 4, cudaMemcpyDeviceToHost, s));
    if (v_host) CUDA_TRY(cudaMemcpyAsync(v_host, sp->view.out_v + (size_t)head * 3, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    if (pi_host) CUDA_TRY(cudaMemcpyAsync(pi_host, sp->view.out_pi + (size_t)head * A, (size_t)n * A * 4, cudaMemcpyDeviceToHost, s));
    if (slot_host) CUDA_TRY(cudaMemcpyAsync(slot_host, sp->view.out_slot + head, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  }
  sp->hist_head = head + n;
  if (sp->hist_head == count) {  // emptied: the ring restarts at row 0 (no launch is in flight: the stream was synchronised)
    CUDA_TRY(cudaMemsetAsync(sp->view.out_count, 0, 4, s));
    sp->hist_head = 0;
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  *n_out = n;
  return 0;
}
// b2az_get_stats for this engine: PlayManager's getters (play_manager.h:173-180, 288-316) over all slots
int b2az_tafl_selfplay_get_stats(b2az_tafl_selfplay* sp, void* stream, b2az_stats* out) {
  using namespace b2az;
  if (!sp || !out) return fail(B2AZ_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t G = sp->view.n_games;
  sp->h_slots.resize(G);
  uint32_t count = 0;
  CUDA_TRY(cudaMemcpyAsync(sp->h_slots.data(), sp->view.slots, (size_t)G * sizeof(SpSlot), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(&count, sp->view.out_count, 4, cudaMemcpyDeviceToHost, s));
  sp->h_tree_err.resize(sp->forest->view.n_trees);
  CUDA_TRY(cudaMemcpy2DAsync(sp->h_tree_err.data(), 4, &sp->forest->view.trees[0].error, sizeof(ForestTree), 4,
                             (size_t)sp->forest->view.n_trees, cudaMemcpyDeviceToHost, s));
  unsigned long long ctr[4] = {0, 0, 0, 0};
  if (sp->view.cache.sets) CUDA_TRY(cudaMemcpyAsync(ctr, sp->view.cache.ctr, sizeof(ctr), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  memset(out, 0, sizeof(*out));
  out->cache_hits = ctr[0]; out->cache_misses = ctr[1]; out->cache_evictions = ctr[3];
  out->cache_size = ctr[2] - ctr[3]; out->cache_max_size = (uint64_t)sp->view.cache.sets * 4u;
  for (uint32_t e : sp->h_tree_err) {  // the trees' sticky bits (az_forest.h ForestTree::error)
    if (e & (1u | 4u | 16u)) out->device_error |= B2AZ_DEVERR_POOL;
    if (e & 2u) out->device_error |= B2AZ_DEVERR_DEPTH;
    if (e & 8u) out->device_error |= B2AZ_DEVERR_MOVE;
  }
  double leaf_depth = 0, entropy = 0, valid = 0, fast_depth = 0, fast_entropy = 0;
  uint64_t moves = 0, full = 0, length = 0, fast = 0;
  for (const SpSlot& g : sp->h_slots) {
    for (int i = 0; i < 3; ++i) out->resign_scores[i] += g.resign_scores[i];
    fast_depth += g.fast_leaf_depth; fast_entropy += g.fast_entropy; fast += g.total_fast_move_count;
    out->simulations += g.simulations;
    out->games_completed += g.games_completed;
    out->games_started += g.games_started;
    out->active_games += g.active ? 1u : 0u;
    for (int i = 0; i < 3; ++i) out->scores[i] += g.scores[i];
    leaf_depth += g.leaf_depth; entropy += g.entropy; valid += g.valid_moves;
    moves += g.total_move_count; full += g.total_full_move_count; length += g.game_length;
    if (g.error & 1u) out->device_error |= B2AZ_DEVERR_HIST;
  }
  out->moves = moves;
  out->hist_count = sp->view.history_enabled ? std::min(count, sp->view.out_cap) - std::min(sp->hist_head, count) : 0u;
  out->avg_game_length = (float)length / (float)out->games_completed;
  if (full) {
    out->avg_leaf_depth = (float)(leaf_depth / (double)full);
    out->avg_search_entropy = (float)(entropy / (double)full);
  }
  if (fast) {
    out->fast_avg_leaf_depth = (float)(fast_depth / (double)fast);
    out->fast_avg_search_entropy = (float)(fast_entropy / (double)fast);
  }
  out->fast_move_count = fast; out->fast_sum_leaf_depth = fast_depth; out->fast_sum_search_entropy = fast_entropy;
  if (length) out->avg_moves_per_turn = (float)moves / (float)length;
  if (moves) out->avg_valid_moves = (float)(valid / (double)moves);
  out->sum_game_length = length;
  out->total_move_count = moves; out->full_move_count = full;
  out->sum_leaf_depth = leaf_depth; out->sum_search_entropy = entropy; out->sum_valid_moves = valid;
  return 0;
}
// PlayManager's per-variant tables summed over the slots: out[4] (num_tracked_variants() == 4 for StarGambitUnifiedGS)
int b2az_tafl_selfplay_variant_stats(b2az_tafl_selfplay* sp, void* stream, b2az_variant_stats* out4) {
  using namespace b2az;
  if (!sp || !out4) return fail(B2AZ_EINVAL, "null argument");
  static_assert(sizeof(b2az_variant_stats) == sizeof(SpVariantAcc), "variant stats layout");
  memset(out4, 0, 4 * sizeof(b2az_variant_stats));
  if (!sp->view.variants) return fail(B2AZ_ESTATE, "this game has no variants");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)sp->view.n_games * 4u;
  std::vector<SpVariantAcc> h(n);
  CUDA_TRY(cudaMemcpyAsync(h.data(), sp->view.variants, n * sizeof(SpVariantAcc), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i) {
    const SpVariantAcc& a = h[i];
    b2az_variant_stats& o = out4[i & 3u];
    for (int j = 0; j < 3; ++j) o.scores[j] += a.scores[j];
    o.games_completed += a.games_completed; o.game_length += a.game_length; o.total_move_count += a.total_move_count;
    o.full_move_count += a.full_move_count; o.fast_move_count += a.fast_move_count;
    o.leaf_depth += a.leaf_depth; o.entropy += a.entropy; o.valid_moves += a.valid_moves;
    o.fast_leaf_depth += a.fast_leaf_depth; o.fast_entropy += a.fast_entropy;
  }
  return 0;
}
// The reference-API flavour of one simulation (build_batch / update_inferences with HOST buffers, py_wrapper.cc:449-504,
// play_manager.cc:619-642): find_leaf for every active slot, then the leaves' canonical planes compacted into
// canon_host float32[n][P][S][S] with the slot ids in ids_host[n] (ascending) ...
int b2az_tafl_selfplay_leaf_batch_host(b2az_tafl_selfplay* sp, void* stream, uint32_t max_rows, float* canon_host,
                                       uint32_t* ids_host, uint32_t* n_out) {
  using namespace b2az;
  if (!sp || !canon_host || !ids_host || !n_out) return fail(B2AZ_EINVAL, "null argument");
  const float* dev = nullptr;
  if (int rc = b2az_tafl_selfplay_find_leaf(sp, stream, &dev)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t G = sp->view.n_games;
  const size_t C = sp->forest->canon;
  sp->h_canon.resize((size_t)G * C);
  sp->h_slots.resize(G);
  CUDA_TRY(cudaMemcpyAsync(sp->h_canon.data(), dev, (size_t)G * C * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(sp->h_slots.data(), sp->view.slots, (size_t)G * sizeof(SpSlot), cudaMemcpyDeviceToHost, s));
  if (sp->view.wait) {
    sp->h_wait.resize(G);
    CUDA_TRY(cudaMemcpyAsync(sp->h_wait.data(), sp->view.wait, (size_t)G * 4, cudaMemcpyDeviceToHost, s));
  }
  sp->h_group.clear();
  if (sp->view.leaf_group) {
    sp->h_group_all.resize(G);
    CUDA_TRY(cudaMemcpyAsync(sp->h_group_all.data(), sp->view.leaf_group, (size_t)G, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  uint32_t n = 0;
  for (uint32_t g = 0; g < G; ++g) {
    if (!sp->h_slots[g].active) continue;
    if (sp->view.wait && !sp->h_wait[g]) continue;  // (cache) its leaves hit: nothing for the evaluator
    if (n >= max_rows) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay_leaf_batch_host: more active slots than max_rows");
    memcpy(canon_host + (size_t)n * C, sp->h_canon.data() + (size_t)g * C, C * 4);
    sp->h_group.push_back(sp->view.leaf_group ? sp->h_group_all[g] : (uint8_t)0);
    ids_host[n++] = g;
  }
  *n_out = n;
  return 0;
}
// ... and the answers (v float32[n][3], pi float32[n][A], row i for slot ids[i]; every active slot must be answered):
// process_result + the move of every slot whose search is complete.
int b2az_tafl_selfplay_submit_eval_host(b2az_tafl_selfplay* sp, void* stream, const uint32_t* ids, const float* v, const float* pi,
                                        uint32_t n) {
  using namespace b2az;
  if (!sp || (n && (!ids || !v || !pi))) return fail(B2AZ_EINVAL, "null argument");
  const uint32_t G = sp->view.n_games;
  const size_t A = sp->forest->actions;
  sp->h_v.assign((size_t)G * 3, 0.0f);
  sp->h_pi.resize((size_t)G * A);
  for (uint32_t i = 0; i < n; ++i) {
    if (ids[i] >= G) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay_submit_eval_host: slot id out of range");
    memcpy(&sp->h_v[(size_t)ids[i] * 3], v + (size_t)i * 3, 12);
    memcpy(&sp->h_pi[(size_t)ids[i] * A], pi + (size_t)i * A, A * 4);
  }
  return b2az_tafl_selfplay_process_result(sp, stream, sp->h_v.data(), sp->h_pi.data(), 1, nullptr);
}
int b2az_tafl_selfplay_root_state(b2az_tafl_selfplay* sp, uint32_t slot, void* state, uint32_t state_bytes, void* hist,
                                  uint32_t hist_cap, uint32_t* hist_count) {
  using namespace b2az;
  if (!sp) return fail(B2AZ_EINVAL, "null argument");
  if (slot >= sp->view.n_games) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay_root_state: slot out of range");
  return b2az_forest_get_root(sp->forest, 2u * slot, state, state_bytes, hist, hist_cap, hist_count);
}
int b2az_tafl_selfplay_leaf_groups_host(b2az_tafl_selfplay* sp, uint8_t* groups_host, uint32_t n) {
  using namespace b2az;
  if (!sp || (n && !groups_host)) return fail(B2AZ_EINVAL, "null argument");
  if (n > sp->h_group.size()) return fail(B2AZ_EINVAL, "b2az_tafl_selfplay_leaf_groups_host: more rows than the last leaf batch had");
  if (n) memcpy(groups_host, sp->h_group.data(), n);
  return 0;
}
// perm_scores_ / variant_perm_scores_ (play_manager.cc:205-255, 466-474): slot g plays permutation g % n_perms
int b2az_tafl_selfplay_perm_stats(b2az_tafl_selfplay* sp, void* stream, b2az_perm_stats* out8, uint32_t* n_perms_out) {
  using namespace b2az;
  if (!sp || !out8) return fail(B2AZ_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t G = sp->view.n_games, P = sp->view.n_perms;
  memset(out8, 0, P * sizeof(b2az_perm_stats));
  sp->h_slots.resize(G);
  CUDA_TRY(cudaMemcpyAsync(sp->h_slots.data(), sp->view.slots, (size_t)G * sizeof(SpSlot), cudaMemcpyDeviceToHost, s));
  std::vector<SpVariantAcc> hv;
  if (sp->view.variants) {
    hv.resize((size_t)G * 4u);
    CUDA_TRY(cudaMemcpyAsync(hv.data(), sp->view.variants, hv.size() * sizeof(SpVariantAcc), cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  for (uint32_t g = 0; g < G; ++g) {
    b2az_perm_stats& o = out8[g % P];
    for (int i = 0; i < 3; ++i) o.scores[i] += sp->h_slots[g].scores[i];
    o.games_completed += sp->h_slots[g].games_completed;
    if (!hv.empty())
      for (uint32_t v = 0; v < 4u; ++v) {
        const SpVariantAcc& a = hv[(size_t)g * 4u + v];
        for (int i = 0; i < 3; ++i) o.variant_scores[v][i] += a.scores[i];
        o.variant_games_completed[v] += a.games_completed;
      }
  }
  if (n_perms_out) *n_perms_out = P;
  return 0;
}
int b2az_tafl_selfplay_slots(b2az_tafl_selfplay* sp, void* stream, b2az_tafl_selfplay_slot* slots_host, uint32_t* tree_errors_host) {
  using namespace b2az;
  if (!sp) return fail(B2AZ_EINVAL, "null argument");
  static_assert(sizeof(b2az_tafl_selfplay_slot) == sizeof(SpSlot), "slot layout");
  CUDA_TRY(cudaSetDevice(sp->forest->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (slots_host)
    CUDA_TRY(cudaMemcpyAsync(slots_host, sp->view.slots, (size_t)sp->view.n_games * sizeof(SpSlot), cudaMemcpyDeviceToHost, s));
  if (tree_errors_host) {
    // sticky error bits of the 2 * n_games trees (slab full, path too long, ...)
    CUDA_TRY(cudaMemcpy2DAsync(tree_errors_host, 4, &sp->forest->view.trees[0].error, sizeof(ForestTree), 4,
                               (size_t)sp->forest->view.n_trees, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}
#endif

}  // extern "C"
