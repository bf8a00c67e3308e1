// az_engine_types.h — device-resident data layout of the self-play pool.
//
// Layout in HBM (DESIGN.md "Data layout"):
//   node pool     structure-of-arrays over a global node index: q[], pol[], n[] (the three fields a
//                 PUCT scan reads), mv[] (u16 move of the edge into the node) and rec[] (16 B:
//                 first-child index, child count, side to move, terminal code, v, d). The children
//                 of a node form ONE contiguous block, 8-node aligned and padded to a multiple of
//                 8, so q/pol/n of a Connect4 block are exactly one 32 B sector each.
//   pages         the pool is cut into pages of 2^kPageLog2 nodes; a tree owns a chain of pages
//                 and bump-allocates blocks inside its current page. Free pages sit on sharded
//                 lock-free stacks. Re-rooting (MCTS::update_root, mcts.cc:151-173) Cheney-copies
//                 the kept subtree into fresh pages in BFS order and frees the old chain, which
//                 is the device equivalent of the reference's move + recursive ~Node.
//   trees         one TreeHdr per (game slot, seat): the root node's scalars live here, not in
//                 the pool (reference: MCTS::root_ member, mcts.h:161).
//   games         one GameSlot per concurrent game (reference: GameData, play_manager.h:33-58).
#pragma once

#include "az_common.h"
#include "az_connect4.h"
#include "az_rng.h"

namespace b2az {

constexpr u32 kNil = 0xFFFFFFFFu;
constexpr u32 kRootRef = 0xFFFFFFFEu;  // "the leaf / current node is the tree's root"
constexpr int kPageLog2 = 8;           // 256 nodes per page
constexpr u32 kPageNodes = 1u << kPageLog2;
constexpr int kMaxPath = 44;           // Connect4: at most 42 plies below any root
constexpr int kMaxHist = 42;           // recorded moves per game
constexpr int kNumStacks = 64;         // free-page stack shards
constexpr int kA = 7;                  // Connect4 action count
constexpr int kP = 2;                  // players
constexpr int kKMax = 8;               // padded child-block size for Connect4

struct __attribute__((aligned(16))) NodeRec {
  u32 fc;     // first child (global node index) — valid once the node is expanded
  u16 k;      // number of children (0: terminal or no legal move)
  u8 player;  // side to move AT this node (mcts.cc:491)
  u8 term;    // 0 = not terminal, else 1 + index of the one-hot score (mcts.cc:492-494)
  float v;    // node value from its own player's perspective, set on first visit (mcts.cc:538-542)
  float d;    // running mean of the draw share (mcts.cc:535-537)
};

struct __attribute__((aligned(16))) TreeHdr {
  // root node scalars (Node fields, mcts.h:18-26)
  float q, d, v, policy;
  u32 n;
  u32 fc;
  u16 k;
  u8 player;
  u8 term;
  u16 move;
  u16 path_len;     // length of the stored selection path (MCTS::path_)
  // MCTS members (mcts.h:159-160)
  u32 depth;            // depth_: simulations finished in this search
  u32 total_leaf_depth; // total_leaf_depth_
  // arena
  u32 first_page, cur_page;
  u32 bump;             // next free node offset inside cur_page
  u32 leaf;             // current_: node index of the pending leaf, or kRootRef
  u32 pad_[2];
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

struct __attribute__((aligned(16))) GameSlot {
  u64 p0, p1;
  u32 turn;
  u8 player;
  u8 initialized;  // GameData::initialized
  u8 capped;       // GameData::capped
  u8 active;       // 0 once the slot retired (play_manager.cc:506-509)
  u32 eval_row;    // row of this game's leaf in the evaluation batch
  u32 hist_n;      // entries in partial_history
  u32 move_count, full_move_count, fast_move_count;
  u32 leaf_k;      // legal-move count at the pending leaf (RANDOM eval: dumb_eval needs only this)
  double total_avg_leaf_depth, total_search_entropy;
  double fast_total_avg_leaf_depth, fast_total_search_entropy;
  double total_valid_moves;
  Pcg32 rng;       // per-game stream (B2AZ_RNG_PER_GAME)
  unsigned long long sims;    // simulations finished in this slot (summed on demand; no global atomic per sim)
  unsigned long long nmoves;  // moves played in this slot
  u32 pad_[2];
};
static_assert(sizeof(GameSlot) == 128, "GameSlot must stay one 128 B line");

struct Globals {
  unsigned long long simulations, moves, game_length;
  unsigned long long wins[3], resign_wins[3];
  unsigned long long total_move_count, full_move_count, fast_move_count;
  unsigned long long hist_written, hist_read;
  unsigned long long cache_hits, cache_misses, cache_evictions, cache_reinserts, cache_size;
  double total_avg_leaf_depth, total_search_entropy, fast_total_avg_leaf_depth, fast_total_search_entropy;
  double total_valid_moves;
  u32 games_completed, games_started, active_games, error;
  u32 leaf_count;
  u32 pad_;
  Pcg32 global_rng;  // B2AZ_RNG_GLOBAL
};

struct EngineView {
  // ---- parameters (PlayParams subset, play_manager.h:60-154)
  u32 G, games_to_play;
  u32 visits[2], cap_visits[2];
  float cpuct, fpu_reduction, epsilon, root_temp;
  float start_temp, final_temp, half_life, playout_cap_percent;
  u8 history_enabled, tree_reuse, root_fpu_zero, shaped_dirichlet;
  u8 policy_target_pruning, playout_cap, eval_type, rng_mode;
  u32 num_pages, hist_capacity;
  // ---- node pool
  float* q;
  float* pol;
  u32* n;
  u16* mv;
  NodeRec* rec;
  u32* page_next;
  u32* page_fill;
  unsigned long long* stack_head;  // [kNumStacks] (tag << 32 | top page)
  // ---- per tree / per game
  TreeHdr* trees;    // [G * kP]
  GameSlot* games;   // [G]
  u32* path;         // [G][kMaxPath]
  // ---- evaluation in (batch-row order) and leaf batch out
  const float* ev_v;   // [rows][kP + 1]
  const float* ev_pi;  // [rows][kA]
  u64* leaf_p0;
  u64* leaf_p1;
  u8* leaf_player;
  u32* leaf_game;
  // ---- history
  HistEntry* hist_partial;  // [G][kMaxHist]
  HistEntry* hist_out;      // ring of hist_capacity
  Globals* glob;
};

}  // namespace b2az
