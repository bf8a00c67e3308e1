// az_engine_types.h — device-resident data layout of the self-play pool.
//
// Layout in HBM (DESIGN.md "Data layout"):
//   block pool    the unit of tree storage is the CHILD BLOCK of one expanded node: 160 B = five 32 B
//                 sectors holding, for its (<= 7) children, one 16 B record {n, q, pol, d} each, the
//                 children's own block indices fc[7], and a 16 B header (moves as nibbles, terminal codes,
//                 and the OWNING node's v / k / player). A PUCT step is therefore ONE independent burst of
//                 ten 16 B loads (everything the selection, the move replay, the next hop AND the later
//                 backprop need), instead of the pointer chase node -> children vector -> child -> its
//                 vector of the reference (mcts.h:14-48); backprop dirties one sector per level.
//   pages         the pool is cut into pages of 64 blocks (10 KB); a tree owns a chain of pages and
//                 bump-allocates blocks in its last page. Free pages sit in one ticket ring (two
//                 atomicAdd tickets, no CAS loop, no fence).
//   trees         one 64 B TreeHdr per (game slot, seat): the root node's scalars live here
//                 (reference: MCTS::root_, mcts.h:161) plus the pending leaf and the arena cursor.
//                 Re-rooting (MCTS::update_root, mcts.cc:151-173) just re-points the header at the
//                 chosen child's block; the garbage left behind is squeezed out by a Cheney copy
//                 only when the arena has outgrown its budget (amortised over several moves).
//   games         GameSlot (64 B, hot) + GameCold (64 B, metric accumulators) per concurrent game
//                 (reference: GameData, play_manager.h:33-58).
#pragma once

#include "az_common.h"
#include "az_connect4.h"
#include "az_rng.h"

namespace b2az {

constexpr u32 kNil = 0xFFFFFFFFu;
constexpr int kPageLog2 = 6;           // 64 blocks per page
constexpr u32 kPageBlocks = 1u << kPageLog2;
constexpr int kMaxPath = 44;           // Connect4: at most 42 plies below any root
constexpr int kMaxHist = 42;           // recorded moves per game
constexpr int kA = 7;                  // Connect4 action count
constexpr int kP = 2;                  // players
constexpr int kKMax = 7;               // children per block (= Connect4 actions)
constexpr int kQGames = 448;           // game slots per persistent CTA (shared memory: 272 B each) == per pool region

// Every field is a 32-bit word (floats are stored as their bit patterns) so that scalar accesses and
// the 16 B vector accesses are the same type for the compiler's alias analysis.
struct __attribute__((aligned(32))) Block {   // 160 B = five 32 B sectors, ten 16 B vectors
  u32 rec[kKMax][4];  // +0    per child: n | q | pol | d  (Node::n, ::q, ::policy, ::d; f32 bit patterns).
                      //       One 16 B vector per child: selection reads it, backprop rewrites it whole.
  u32 fc[kKMax];      // +112  the child's OWN block, kNil while unexpanded / terminal
  u32 pad_;           // +140
  u32 mv;             // +144  move of child j in nibble j                       (Node::move)
  u32 term;           // +148  terminal code of child j in bits 2j..2j+1: 0 = not terminal / unknown, else
                      //       1 + one-hot score index                                (Node::scores)
  u32 v;              // +152  f32 v of the node that OWNS this block (Node::v, set on its first visit)
  u32 kp;             // +156  k | player << 8 of that node
};
static_assert(sizeof(Block) == 160, "Block must stay five 32 B sectors");

struct __attribute__((aligned(16))) TreeHdr {
  // root node scalars (Node fields, mcts.h:18-26)
  float q, d, v, policy;
  u32 n;
  u32 fc;           // the root's child block
  u16 move;
  u8 k, player, term;
  // the pending leaf (MCTS::current_ / path_): what process_result needs, recorded by find_leaf
  u8 leaf_term, leaf_k, leaf_player;
  u32 leaf_blk;     // the leaf's own (fresh) block, kNil for a terminal / childless leaf
  u16 path_len;
  u16 pages_used;
  // MCTS members (mcts.h:159-160)
  u32 depth;            // depth_: simulations finished in this search
  u32 total_leaf_depth; // total_leaf_depth_
  // arena
  u32 first_page, cur_page;
  u32 bump;             // next free block offset inside cur_page
  u32 region;           // the pool region this tree allocates from (fixed at creation: slot / region_games)
};
static_assert(sizeof(TreeHdr) == 64, "TreeHdr must stay one 64 B record");

struct __attribute__((aligned(16))) HistEntry {  // one training sample, compact (48 B)
  u64 p0, p1;       // root position the search started from (PlayHistory::canonical, rebuilt on drain)
  float pi[kA];     // policy target
  u8 player;        // side to move at that position
  u8 result;        // 1 + one-hot index of the final score, filled at game end (PlayHistory::v)
  u8 pad_[2];
};
static_assert(sizeof(HistEntry) == 48, "HistEntry layout");

struct __attribute__((aligned(16))) GameSlot {  // hot: touched every simulation
  u64 p0, p1;
  u32 turn;
  u8 player;
  u8 initialized;  // GameData::initialized
  u8 capped;       // GameData::capped
  u8 active;       // 0 once the slot retired (play_manager.cc:506-509)
  u32 eval_row;    // row of this game's leaf in the evaluation batch
  u32 hist_n;      // entries in partial_history
  u32 move_count, full_move_count, fast_move_count;
  u8 playthrough;  // GameData::playthrough: set once, never cleared — like the reference (play_manager.cc:328-329)
  u8 pad_[3];
  Pcg32 rng;       // per-game stream (B2AZ_RNG_PER_GAME)
};
static_assert(sizeof(GameSlot) == 64, "GameSlot must stay one 64 B record");

struct __attribute__((aligned(16))) GameCold {  // touched once per move / per launch
  double total_avg_leaf_depth, total_search_entropy;
  double fast_total_avg_leaf_depth, fast_total_search_entropy;
  double total_valid_moves;
  unsigned long long sims;    // simulations finished in this slot (summed on demand; no global atomic per sim)
  unsigned long long nmoves;  // moves played in this slot
  u32 games_done;             // games finished in this slot (slot_quota)
  u32 pad_;
};
static_assert(sizeof(GameCold) == 64, "GameCold layout");

// Position cache (replaces S3FIFOCache / ShardedS3FIFOCache, s3fifo_cache.h): 4-way set-associative, keyed by
// an exact 64-bit encoding of the position (equality class of Connect4GS::hash, connect4_gs.cc:33-37).
constexpr int kCacheWays = 4;
struct __attribute__((aligned(8))) CacheVal {  // what a hit returns: pi[A] and v[P+1] (s3fifo_cache.h:41-60)
  float pi[kA];
  float v[kP + 1];
};
static_assert(sizeof(CacheVal) == 40, "CacheVal layout");

// Gumbel root-search state of one tree (MCTS::gumbel_* members, mcts.h:163-176). Connect4: <= 7 root
// children, so at most ceil(log2 7) = 3 halving phases.
constexpr int kMaxPhases = 4;
struct __attribute__((aligned(16))) GumbelState {
  float g[kKMax];             // gumbel_g_: one draw per root child, child order
  u32 survivors;              // gumbel_survivors_: child indices as nibbles, in rank order
  u32 phase_numc[kMaxPhases]; // gumbel_phases_[i].first
  u32 phase_vper[kMaxPhases]; // gumbel_phases_[i].second
  u32 num_sims_target;        // gumbel_num_sims_target_
  u32 sims_in_phase;          // gumbel_sims_in_phase_
  u8 n_surv, n_phases, phase_idx, initialized;
  u8 effective_m, pad_[3];
};
static_assert(sizeof(GumbelState) == 80, "GumbelState layout");

constexpr int kMaxPerms = 8;
struct Globals {
  unsigned long long simulations, moves, game_length;
  unsigned long long wins[3], resign_wins[3];
  unsigned long long total_move_count, full_move_count, fast_move_count;
  unsigned long long hist_written, hist_read;
  unsigned long long cache_hits, cache_misses, cache_evictions, cache_reinserts, cache_size;
  unsigned long long compactions, pages_popped;
  double total_avg_leaf_depth, total_search_entropy, fast_total_avg_leaf_depth, fast_total_search_entropy;
  double total_valid_moves;
  u32 games_completed, games_started, active_games, error;
  u32 leaf_count;
  u32 pad_;
  unsigned long long pad2_[2];
  Pcg32 global_rng;  // B2AZ_RNG_GLOBAL
  unsigned long long perm_wins[kMaxPerms][3];  // perm_scores_ (play_manager.cc:205-211, 466-467)
};

// slot g plays permutation g % n_perms in every one of its games (the reference's round-robin hand-out when the slots
// finish in order; G is a multiple of n_perms)
struct PermTables {
  u32 n_perms;
  u32 random_groups;   // NN mode (eval_type 0): bit i = model group i is EvalType::RANDOM, its searches run dumb_eval inline
  u32 visits[kMaxPerms][2], cap_visits[kMaxPerms][2];  // seat_visits_ / seat_cap_visits_[perm][seat]
  u8 seat_group[kMaxPerms][2];                         // seat_perms_[perm][seat]
};

struct EngineView {
  // ---- parameters (PlayParams subset, play_manager.h:60-154)
  u32 G, games_to_play;
  u32 visits[2], cap_visits[2];  // one seating (the usual case); several: EngineView::perms
  float cpuct, fpu_reduction, epsilon, root_temp;
  float start_temp, final_temp, half_life, playout_cap_percent;
  float resign_percent, resign_playthrough_percent;
  float gumbel_c_visit, gumbel_c_scale;
  u32 gumbel_m;
  u8 gumbel_enabled, gumbel_full, fast_search_uses_gumbel, pad0_;
  u8 history_enabled, tree_reuse, root_fpu_zero, shaped_dirichlet;
  u8 policy_target_pruning, playout_cap, eval_type, rng_mode;
  u32 num_pages, hist_capacity;
  u32 compact_pages;   // a tree is compacted at a move once its arena holds more pages than this
  u32 hit_cap;         // position-cache hits a game may chain inside ONE launch before it has to emit a leaf row
  u32 slot_quota;      // > 0: every slot retires after this many games (deterministic: b2az_params.per_slot_quota)
  // ---- block pool
  Block* blocks;       // [num_pages * kPageBlocks]
  u32* page_next;      // [num_pages] chain links
  u32* ring;           // [num_pages] free pages: region r owns ring[r * region_pages .. +region_pages) (kNil = empty slot)
  // The pool is cut into REGIONS of region_pages consecutive pages; the slots [r * region_games, (r+1) * region_games)
  // allocate from and free into region r only. A region is what one CTA of the persistent step kernels owns, so an
  // SM's tree traffic stays inside one contiguous ~250 MB stretch of HBM (about 120 2 MB pages: its TLB holds them).
  // Measured (tools/micro/tlb_probe.cu, profiles/r2i_tlb_probe.jsonl): dependent random 160 B block reads run at
  // 7.2 G/s when every thread roams the whole pool and at 24 G/s when an SM's threads stay inside such a stretch.
  unsigned long long* ring_tickets;  // [n_regions][2]  pop / push ticket counters of each region's ring
  u32 region_games, region_pages, n_regions, pad3_;
  // ---- per tree / per game
  TreeHdr* trees;      // [G * kP]
  GumbelState* gum;    // [G * kP], NULL unless gumbel_enabled
  GameSlot* games;     // [G]
  GameCold* cold;      // [G]
  u32* path;           // [G][kMaxPath]  block index of every selected edge
  u8* pslot;           // [G][kMaxPath]  child slot (bits 0-3) | parent's player (bits 4-7)
  // ---- evaluation in (batch-row order) and leaf batch out
  const float* ev_v;   // [rows][kP + 1]
  const float* ev_pi;  // [rows][kA]
  u64* leaf_p0;
  u64* leaf_p1;
  u8* leaf_player;
  u32* leaf_game;
  u8* leaf_seat;       // [rows] bits 0-3: the SEARCHING seat (the slot's side to move); bits 4-7: its model group, which evaluates the leaf
  u8 seat_group[kP];   // model group of each seat (PlayParams::model_groups, play_manager.cc:24-31) with one seating
  u8 pad4_[2];
  // Seat permutations and a RANDOM group next to an NN one (play_manager.cc:46-90, 213-221, 577-587): a table in device
  // memory, NULL for the usual one-seating run — kept out of this structure, which every kernel takes by value and hands
  // on to its out-of-line helpers (a larger view costs the step kernel 9 %: 656 B more stack per thread, profiles/r4b)
  const PermTables* perms;
  // ---- position cache (NULL / 0 when max_cache_size == 0)
  u64* cache_keys;        // [buckets][kCacheWays]  0 = empty  (one 32 B sector per bucket)
  u32* cache_meta;        // [buckets]  per way one byte: freq (bits 0-1) | main-queue flag (bit 2)
  u32* cache_lock;        // [buckets]  insert-side spin lock
  CacheVal* cache_vals;   // [buckets * kCacheWays]
  u32* cache_ghost;       // [ghost_slots] fingerprints of keys evicted from the small queue
  u32 cache_buckets, cache_ghost_slots;
  u64* leaf_key;          // [G] cache key of every leaf row (insert after the evaluation)
  u32* hit_val;           // [G] per game: index into cache_vals of a pending cache hit, kNil if none
  // ---- history
  HistEntry* hist_partial;  // [G][kMaxHist]
  HistEntry* hist_out;      // ring of hist_capacity
  Globals* glob;
};

}  // namespace b2az
