// az_forest.h — batched wide-tree MCTS ("forest") over the tafl games: MCTS::find_leaf / process_result /
// update_root / counts (mcts.cc:93-173, 462-573) for MANY trees at once, ONE WARP PER TREE. Included at the end
// of az_engine.cu (same translation unit: it uses that file's error / launch helpers).
//
// This is the wide-game counterpart of the Connect4 step kernel (az_engine_logic.h, one thread per game): a tafl
// node has 30-120 children, so the per-node work is spread over the 32 lanes of a warp —
//   selection   Node::best_child (mcts.cc:130-149): the children's {n, q, policy} are read as coalesced arrays, 32
//               children per pass; `seen_policy` is summed IN CHILD ORDER (float addition is not associative: the
//               lanes hand their values round with shuffles and every lane accumulates the same sequence), the PUCT
//               scores are computed one child per lane and reduced to the FIRST maximum (ties -> lowest index)
//   expansion   Node::add_children (mcts.cc:93-101): legal moves in ascending id order from the row/column
//               occupancy lines (az_tafl.h slides_lines), written with a warp prefix sum, then std::shuffle with the
//               tree's own pcg32 stream (lane 0; the draws are sequential by definition)
//   priors      Node::set_policy_normalized (mcts.cc:109-121): gather pi[move] per lane, sequential sum in child
//               order as above, one division per lane
//   backprop    lane 0 walks the (short) path
// The game state is replayed along the path on every lane's register copy (Tafl<GAME>::play); repetition counts
// come from the root's history (HBM) plus the keys of the current path, scanned lane-parallel.
//
// Node storage: a fixed slab of 32-bit words per tree; an expanded node owns one block [k | n[k] q[k] policy[k]
// d[k] v[k] first_child[k] move|player|terminal[k]] (structure of arrays: Node::n/q/policy/d/v/children/move/
// player/scores, mcts.h:14-48), bump-allocated in one half of the slab. Re-rooting (MCTS::update_root,
// mcts.cc:151-173) copies the chosen child's subtree breadth first into the other half (a Cheney copy by the
// whole warp), which is what frees the discarded siblings.
//
// Scope of this version: PUCT selection, or Gumbel root search (sequential halving over the top-m root children,
// improved-policy targets, final action; mcts.cc:175-283, 336-401) with PUCT below the root; root policy
// temperature and (shaped) Dirichlet noise (mcts.cc:403-460), gumbel_full (pi'-matching at interior nodes), relative_values.
#pragma once

#include "az_rng.h"
#include "az_tafl.h"
#include "az_connect4.h"
#include "az_stargambit_kernels.h"

// Template id of the Star Gambit instantiation of the forest: ONE instantiation serves the eight game ids
// (B2AZ_SG_GAME / B2AZ_SG_UNIFIED, ForestView::game carries the id: variant and frame are run-time values)
#define B2AZ_FOREST_SG 10
// Connect4 under the same search (game id 30): what the `MCTS` class of the Python module needs for Connect4GS — the
// batched PlayManager engine (az_engine_logic.h) remains the fast path for Connect4 self-play
#define B2AZ_FOREST_C4 30

namespace b2az {

constexpr int kFPath = 96;      // longest selection path kept (a tafl game is at most max_turns plies deep)
constexpr int kFMaxK = 512;     // most legal moves of one position (11x11: 36 pieces x <= 20 targets in theory)

struct ForestLeaf {             // a pending leaf and the path to it (MCTS::path_/current_, or one InFlightLeaf, mcts.h:129-136)
  u32 path_len;
  u32 leaf_blk, leaf_k, leaf_term, leaf_player, leaf_new;
  u32 path_blk[kFPath];         // block that holds the child selected at level i
  u16 path_slot[kFPath];
  u8 path_player[kFPath];       // player of the node the selection was made AT (the parent of that child)
};
struct ForestTree {             // one per tree, HBM
  TaflState state;              // the root position (GameState of the caller in the reference)
  u32 n;                        // root_.n
  float v, d;                   // root_.v, root_.d
  u32 blk, k;                   // root_'s children block (0 = none) and their number
  u32 player, term;             // root_.player, root_.scores (0 none, else 1 + winner index / 3 = draw)
  u32 depth, total_leaf_depth;  // MCTS::depth_, total_leaf_depth_
  u32 bump;                     // next free word of the slab (word 0 is reserved: 0 = "no block")
  u32 half;                     // which half of the slab is in use (the other one receives the next compaction)
  u32 hist_len;                 // repetition history of the root position (keys since the last capture)
  Pcg32 rng;
  u32 error;                    // sticky: 1 slab full, 2 path too long, 4 too many legal moves, 8 unknown move,
                                //         16 too many in-flight leaves
  u32 nif;                      // root_.n_in_flight (WU-UCT)
  u32 expanded;                 // root_ has had add_children() called (its children may be unstored: terminal root)
  u32 in_flight;                // MCTS::in_flight_.size()
  ForestLeaf leaf;              // MCTS::path_ / current_: the pending leaf of find_leaf
};

constexpr int kFMaxM = 64;       // most Gumbel candidates kept at the root (PlayParams::gumbel_m, default 16)
constexpr int kFMaxPhases = 8;   // ceil(log2 kFMaxM) sequential-halving phases + slack
struct ForestGumbel {            // MCTS::gumbel_* members (mcts.h:163-176), one per tree
  u32 num_sims_target, sims_in_phase;
  u32 n_surv, n_phases, phase_idx, initialized, effective_m, pad_;
  u32 phase_numc[kFMaxPhases], phase_vper[kFMaxPhases];
  u16 survivors[kFMaxM];         // child indices in rank order
};

// The search settings PlayManager::make_mcts hands every seat's MCTS object (play_manager.cc:602-617: seat_epsilon_,
// seat_mcts_root_temp_, seat_root_fpu_zero_, seat_gumbel_*_ [perm][seat]). Set 0 = the forest's own parameters; the
// self-play engine fills one set per (seat permutation, seat): tree t = 2 * slot + seat, slot plays permutation slot % n.
struct SeatSearch {
  float epsilon, root_policy_temp, gumbel_c_visit, gumbel_c_scale;
  u32 gumbel_m;
  u8 root_fpu_zero, gumbel_enabled, gumbel_full, pad_;
};
constexpr int kFSeatSets = 16;
struct ForestView {
  // The search settings (epsilon, root temperature, root FPU, Gumbel) live in device memory, one record per tree. Measured
  // alternatives, all slower: the (permutation, seat) table inside the view (+392 B, which every thread copies to its
  // stack: the helpers take the view by reference), a 24 B in-view default next to a table, and a 16-entry table indexed
  // with ((t >> 1) % n) * 2 + (t & 1) at every use (profiles/r4e_forest_view_ab.jsonl, r4i_seat_table_ab.txt,
  // r4q_per_tree_settings_ab.txt). Against the build before per-seat settings existed the searches keep a loss of 2 %.
  const SeatSearch* seat;  // [n_trees]: one record per tree, like the members of the reference's MCTS objects (no index arithmetic
                           // in the search: the kernels are instruction-fetch bound, profiles/r4p_k_sp_search_ncu_summary.json)
  u32 n_trees, words_per_tree, max_turns, game;
  float cpuct, fpu_reduction;
  u32 shaped_dirichlet;
  u32 serial_shuffle;  // diagnostics: always take the sequential std::shuffle path
  ForestGumbel* gum;   // [n_trees], null unless gumbel_enabled
  float* gum_g;        // [n_trees][2 * kFMaxK]: gumbel_g_ per root child, then scratch scores
  float* noise;        // [n_trees][kFMaxK] Dirichlet draws, null unless epsilon > 0
  ForestTree* trees;
  u32* pool;           // [n_trees][words_per_tree]
  TaflKey* hist;       // [n_trees][max_turns + 2]
  TaflKey* pkeys;      // [n_trees][kFPath + 2] keys of the positions along the current path
  float* leaf_canon;   // [max(1, max_in_flight)][n_trees][CANON]
  ForestLeaf* inflight; // [n_trees][max_in_flight] (WU-UCT), null when max_in_flight == 0
  u32 max_in_flight;
  u32 relative_values;  // MCTS::relative_values_: evaluations arrive in the leaf mover's frame (mcts.cc:522-524)
  u32 actions, canon;   // num_moves and canonical floats of the game (compile-time constants for the tafl games)
  SGState* sg_state;    // Star Gambit: [n_trees] root positions (ForestTree::state is the tafl record)
  u64* sg_hist;         // [n_trees][sg_hist_cap] position keys of the root since the last deploy
  u64* sg_pkeys;        // [n_trees][kFPath + 2] keys appended along the current path
  u32 sg_hist_cap;
  float sg_probs[4];    // game 24 (StarGambitUnifiedGS with the variant mix): the variants' weights
  u32 rng_pair;         // self-play: trees 2g and 2g+1 (the two seats' MCTS objects of game g, play_manager.h game.mcts[])
                        // draw from ONE generator (the reference's thread-local one), kept in tree 2g
};
// the generator tree t draws from
// the search settings of tree t (SeatSearch): the forest's own, or the (permutation, seat) record of a self-play engine
AZ_HD SeatSearch fseat(const ForestView& F, u32 t) { return F.seat[t]; }
#define FSEAT(F, t) fseat((F), (t))
// the three flag bytes of the record (root_fpu_zero | gumbel_enabled << 8 | gumbel_full << 16) in one load: what a descent reads
AZ_HD u32 fseat_flags(const ForestView& F, u32 t) {
  const SeatSearch* p = F.seat + t;
  return (u32)p->root_fpu_zero | ((u32)p->gumbel_enabled << 8) | ((u32)p->gumbel_full << 16);
}
// The forest kernels take their views by value and their out-of-line helpers by reference, i.e. every thread keeps a copy
// on its stack. __grid_constant__ parameters (no copy, helpers read the constant bank through a generic pointer) measured
// 5 % slower on k_forest_simulate and equal on the self-play kernels (profiles/r4e_forest_view_ab.jsonl): not used here.
#ifdef B2AZ_EXP_GC_FOREST  /* experiment build */
#define AZ_GC_F AZ_GRID_CONSTANT
#else
#define AZ_GC_F
#endif
#define FOREST_RNG(F, t) ((F).trees[(F).rng_pair ? ((t) & ~1u) : (t)].rng)

#ifndef B2AZ_HOST_EMU
// block field offsets (words) for a block of k children at word b: [b] = k, then eight arrays of k words
__device__ __forceinline__ u32 fb_n(u32 b, u32 k) { (void)k; return b + 1u; }
__device__ __forceinline__ u32 fb_q(u32 b, u32 k) { return b + 1u + k; }
__device__ __forceinline__ u32 fb_pol(u32 b, u32 k) { return b + 1u + 2u * k; }
__device__ __forceinline__ u32 fb_d(u32 b, u32 k) { return b + 1u + 3u * k; }
__device__ __forceinline__ u32 fb_v(u32 b, u32 k) { return b + 1u + 4u * k; }
__device__ __forceinline__ u32 fb_fc(u32 b, u32 k) { return b + 1u + 5u * k; }
__device__ __forceinline__ u32 fb_mv(u32 b, u32 k) { return b + 1u + 6u * k; }  // move | player << 16 | term << 20 | expanded << 22
__device__ __forceinline__ u32 fb_nif(u32 b, u32 k) { return b + 1u + 7u * k; } // Node::n_in_flight (WU-UCT)
__device__ __forceinline__ u32 fb_words(u32 k) { return 1u + 8u * k; }

template <int GAME>
struct ForestSmem {   // per warp
  u32 lines[2 * Tafl<GAME>::S + 2];
  u16 moves[kFMaxK];
  u32 draws[kFMaxK / 2 + 2];   // the shuffle's uniform draws, generated lane-parallel (pcg32 jump-ahead)
};

template <>
struct ForestSmem<B2AZ_FOREST_C4> {
  u16 moves[8];
  u32 draws[8];
};
template <>
struct ForestSmem<B2AZ_FOREST_SG> : SGWarpSmem {  // map / moves / cell_unit of the Star Gambit warp functions
  u32 draws[kSGMaxK / 2 + 2];
};

// pcg32 is an LCG underneath: state_d = A_d * state_0 + C_d * inc with A_d = M^d, C_d = 1 + M + ... + M^(d-1)
// (mod 2^64). With the table every lane can produce "the d-th output from now" on its own, so the k/2 draws of a
// std::shuffle are generated 32 at a time instead of one after the other on lane 0.
struct PcgJump {
  u64 a[kFMaxK / 2 + 2], c[kFMaxK / 2 + 2];
};
constexpr PcgJump make_pcg_jump() {
  PcgJump t{};
  u64 a = 1, c = 0;
  for (int i = 0; i < kFMaxK / 2 + 2; ++i) {
    t.a[i] = a;
    t.c[i] = c;
    a = a * AZ_PCG_MULT;
    c = c * AZ_PCG_MULT + 1ULL;
  }
  return t;
}
__device__ constexpr PcgJump kPcgJump = make_pcg_jump();
__device__ __forceinline__ u32 pcg32_output(u64 old) {  // XSH-RR of a state (pcg32_next without the advance)
  const u32 xorshifted = (u32)(((old >> 18) ^ old) >> 27);
  const u32 rot = (u32)(old >> 59);
  return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
}
// std::shuffle(moves, moves + k, rng) (pairwise path, stl_algo.h:3766-3799) with the draws generated by all lanes.
// Lemire's method rejects (and redraws) with probability range / 2^32 per draw: any draw that MIGHT reject sends
// the whole shuffle down the sequential path (rng_shuffle on lane 0), so the result is always the reference's.
__device__ __forceinline__ void forest_shuffle(Pcg32& rng, u16* a, u32* draws, u32 k, u32 lane, bool force_serial) {
  if (k < 2u) return;
  const u32 even = (k & 1u) ? 0u : 1u, i0 = even ? 2u : 1u;
  const u32 pairs = k > i0 ? (k - i0 + 1u) / 2u : 0u, D = even + pairs;
  // measured: below ~64 children the table loads and 64-bit multiplies cost more than lane 0's serial draws
  // (Brandubh, k ~ 33: 136 -> 129 M sims/s with the parallel path; OpenTafl, k ~ 113: 47.6 -> 48.6 M)
  bool risky = force_serial || k < 64u;
  if (!risky)
  for (u32 d = lane; d < D; d += 32u) {
    const u64 st = kPcgJump.a[d] * rng.state + kPcgJump.c[d] * rng.inc;
    u32 range = 2u;
    if (!(even && d == 0u)) {
      const u32 i = i0 + 2u * (d - even);
      range = (i + 1u) * (i + 2u);
    }
    const u64 product = (u64)pcg32_output(st) * (u64)range;
    risky |= (u32)product < range;
    draws[d] = (u32)(product >> 32);
  }
  if (__any_sync(0xFFFFFFFFu, risky)) {
    if (lane == 0) rng_shuffle<u16>(rng, a, k);
    rng.state = __shfl_sync(0xFFFFFFFFu, rng.state, 0);
    __syncwarp();
    return;
  }
  __syncwarp();
  if (lane == 0) {
    u32 d = 0, i = 1;
    if (even) {
      const u32 j = draws[d++];
      const u16 t = a[1]; a[1] = a[j]; a[j] = t;
      i = 2;
    }
    for (; i < k; i += 2u) {
      u32 p0, p1;
      small_divmod(draws[d++], i + 2u, p0, p1);
      u16 t = a[i]; a[i] = a[p0]; a[p0] = t;
      t = a[i + 1u]; a[i + 1u] = a[p1]; a[p1] = t;
    }
  }
  rng.state = kPcgJump.a[D] * rng.state + kPcgJump.c[D] * rng.inc;
  __syncwarp();
}

// In-order float sums over a chunk of 32 children held one per lane (float addition is not associative and the
// children are in shuffled order, so the order of the reference's loop must be kept): the lanes hand their value
// round with shuffles and every lane accumulates the same sequence.
__device__ __forceinline__ float seq_sum_masked(float acc, float val, bool take) {  // only lanes with `take`, ascending
  u32 m = __ballot_sync(0xFFFFFFFFu, take);
  while (m) {
    const int t = __ffs((int)m) - 1;
    m &= m - 1u;
    acc = fadd(acc, __shfl_sync(0xFFFFFFFFu, val, t));
  }
  return acc;
}
__device__ __forceinline__ float seq_sum_all(float acc, float val, u32 count) {  // lanes 0 .. count-1
  for (u32 t = 0; t < count; ++t) acc = fadd(acc, __shfl_sync(0xFFFFFFFFu, val, (int)t));
  return acc;
}

// play_move along the selection path, repetition count from the root's history + the path's own keys
// (BrandubhGS::play_move & co; az_tafl.h play_hist restated for a read-only root history)
template <int GAME>
__device__ __forceinline__ bool forest_play(TaflState& s, u32 mv, const TaflKey* hist, u32& base_len, TaflKey* pkeys,
                                            u32& pk_len, u32 lane) {
  typedef Tafl<GAME> T;
  if (s.turn == 0) {  // the start position enters the (copy's) table with the first move
    base_len = 0;
    if (lane == 0) pkeys[0] = T::key(s);
    pk_len = 1;
    __syncwarp();
  }
  bool cap;
  if (!T::play(s, mv, &cap)) return false;
  if (cap) { base_len = 0; pk_len = 0; }
  const TaflKey key = T::key(s);
  u32 same = 0;
  for (u32 i = lane; i < base_len; i += 32u) same += T::key_eq(hist[i], key) ? 1u : 0u;
  for (u32 i = lane; i < pk_len; i += 32u) same += T::key_eq(pkeys[i], key) ? 1u : 0u;
  same = warp_sum(same) + 1u;
  if (lane == 0) pkeys[pk_len] = key;
  ++pk_len;
  __syncwarp();
  s.rep = (u8)(same > 255u ? 255u : same);
  return true;
}

// Node::add_children(valid_moves()) into sm.moves (ascending ids, then std::shuffle); returns k
template <int GAME>
__device__ __forceinline__ u32 forest_legal_moves(const TaflState& s, ForestSmem<GAME>& sm, Pcg32& rng, u32 lane, u32* err,
                                                  bool serial_shuffle = false) {
  typedef Tafl<GAME> T;
  constexpr int S = T::S, CELLS = T::CELLS, CHUNKS = (CELLS + 31) / 32;
  const B128 occ = s.king | s.def | s.atk;
  if (lane < (u32)S) {
    sm.lines[lane] = b128_bits(occ, S * (int)lane, S);
  } else if (lane < 2u * S) {
    const int c = (int)lane - S;
    u32 v = 0;
#pragma unroll
    for (int h = 0; h < S; ++h) v |= (b128_test(occ, S * h + c) ? 1u : 0u) << h;
    sm.lines[lane] = v;
  }
  __syncwarp();
  const B128 mine = T::own(s);
  u32 base = 0;
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j) {
    const u32 c = 32u * j + lane;
    u32 r = 0, cl = 0;
    if (c < (u32)CELLS && ((b128_word(mine, j) >> lane) & 1u)) {
      const int h = (int)(c / (u32)S), w = (int)(c % (u32)S);
      T::slides_lines(((b128_word(s.king, j) >> lane) & 1u) != 0, h, w, sm.lines[h], sm.lines[S + w], r, cl);
    }
    const u32 mycnt = (u32)__popc(r) + (u32)__popc(cl);
    u32 incl = mycnt;  // inclusive warp prefix sum
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((int)lane >= o) incl += up;
    }
    u32 off = base + incl - mycnt;
    const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (base + total <= (u32)kFMaxK) {
      for (u32 m = r; m; m &= m - 1u) sm.moves[off++] = (u16)(c * (u32)(2 * S) + (u32)(__ffs((int)m) - 1));
      for (u32 m = cl; m; m &= m - 1u) sm.moves[off++] = (u16)(c * (u32)(2 * S) + (u32)S + (u32)(__ffs((int)m) - 1));
    }
    base += total;
  }
  __syncwarp();
  u32 k = base;
  if (k > (u32)kFMaxK) { *err |= 4u; k = 0; }
  forest_shuffle(rng, sm.moves, sm.draws, k, lane, serial_shuffle);  // every lane ends with the same generator state
  return k;
}

// ---- The game under the search. FGame<GAME>::Pos is the working position of one descent (every lane holds a copy);
// the tafl games keep three bitboards, Star Gambit a unit list (az_stargambit_kernels.h).
template <int GAME>
struct FGame {  // Brandubh / OpenTafl / Tawlbwrdd
  typedef Tafl<GAME> T;
  struct Pos {
    TaflState s;
    const TaflKey* hist;
    TaflKey* pkeys;
    u32 base_len, pk_len;
  };
  static __device__ __forceinline__ u32 actions(const ForestView&) { return (u32)T::A; }
  static __device__ __forceinline__ u32 canon(const ForestView&) { return (u32)T::CANON; }
  static __device__ __forceinline__ void open(const ForestView& F, u32 t, u32, ForestSmem<GAME>&, Pos& P) {
    P.s = F.trees[t].state;
    P.hist = F.hist + (size_t)t * (F.max_turns + 2u);
    P.pkeys = F.pkeys + (size_t)t * (kFPath + 2);
    P.base_len = F.trees[t].hist_len;
    P.pk_len = 0;
  }
  static __device__ __forceinline__ bool play(Pos& P, u32 mv, u32 lane) {
    return forest_play<GAME>(P.s, mv, P.hist, P.base_len, P.pkeys, P.pk_len, lane);
  }
  static __device__ __forceinline__ u32 player(const Pos& P) { return P.s.player; }
  static __device__ __forceinline__ u32 legal(const Pos& P, ForestSmem<GAME>& sm, Pcg32& rng, u32 lane, u32* err, bool serial) {
    return forest_legal_moves<GAME>(P.s, sm, rng, lane, err, serial);
  }
  static __device__ __forceinline__ u32 terminal(const Pos& P, u32 k) {
    const u32 pre = T::terminal_pre(P.s);
    return pre ? pre : T::terminal_post(P.s, k != 0);
  }
  static __device__ __forceinline__ void emit_canon(const Pos& P, ForestSmem<GAME>&, float* out, u32 lane) {
    const TaflState& s = P.s;
    constexpr int CELLS = T::CELLS, CHUNKS = (CELLS + 31) / 32;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) {
      const u32 c = 32u * j + lane;
      if (c < (u32)CELLS) {
        out[c] = (float)((b128_word(s.king, j) >> lane) & 1u);
        out[CELLS + c] = (float)((b128_word(s.def, j) >> lane) & 1u);
        out[2 * CELLS + c] = (float)((b128_word(s.atk, j) >> lane) & 1u);
      }
    }
#pragma unroll
    for (int pl = 3; pl < T::PLANES; ++pl) {
      const float v = T::canon_elem(s, (u32)(pl * CELLS));
#pragma unroll
      for (int j = 0; j < CHUNKS; ++j) {
        const u32 c = 32u * j + lane;
        if (c < (u32)CELLS) out[pl * CELLS + c] = v;
      }
    }
  }
  // hash_game_state's equality class (brandubh_gs.cc:105-109: board + side to move + repetition count; OpenTafl adds the
  // turn, opentafl_gs.cc:102-107) as a 64-bit key of the position cache; the VALUE is free, 0 is reserved
  static __device__ __forceinline__ u64 state_key(const Pos& P, const ForestView&) {
    const TaflState& s = P.s;
    u64 h = 0x9E3779B97F4A7C15ULL + (u64)GAME;
    const u64 w[7] = {s.king.lo, s.king.hi, s.def.lo, s.def.hi, s.atk.lo, s.atk.hi,
                      (u64)s.player | ((u64)s.rep << 8) | (GAME == B2AZ_TAFL_OPENTAFL ? (u64)s.turn << 16 : 0ULL)};
#pragma unroll
    for (int i = 0; i < 7; ++i) { h = (h ^ w[i]) * 0xBF58476D1CE4E5B9ULL; h ^= h >> 29; }
    h = (h ^ (h >> 32)) * 0x94D049BB133111EBULL;
    h ^= h >> 31;
    return h ? h : 1ULL;
  }
  // a training sample's position while its game is still running: the canonical planes themselves (a few hundred floats)
  static __device__ __forceinline__ u32 stage_floats(const ForestView&) { return (u32)T::CANON; }
  static __device__ __forceinline__ void stage(const Pos& P, ForestSmem<GAME>& sm, float* row, u32 lane) { emit_canon(P, sm, row, lane); }
  static __device__ __forceinline__ void unstage(const ForestView&, const float* row, ForestSmem<GAME>&, float* out, u32 lane) {
    for (u32 e = lane; e < (u32)T::CANON; e += 32u) out[e] = row[e];
  }
  // gs.play_move(move) on the tree's root position with its persistent repetition history (update_root's caller)
  static __device__ __forceinline__ u32 root_play(const ForestView& F, u32 t, u32 move, ForestSmem<GAME>&, u32 lane) {
    ForestTree& R = F.trees[t];
    TaflKey* hist = F.hist + (size_t)t * (F.max_turns + 2u);
    TaflState s = R.state;
    u32 hist_len = R.hist_len, err = 0;
    if (s.turn == 0) {
      if (lane == 0) hist[0] = T::key(s);
      hist_len = 1;
      __syncwarp();
    }
    bool cap;
    if (!T::play(s, move, &cap)) return 8u;
    if (cap) hist_len = 0;
    const TaflKey key = T::key(s);
    u32 same = 0;
    for (u32 i = lane; i < hist_len; i += 32u) same += T::key_eq(hist[i], key) ? 1u : 0u;
    same = warp_sum(same) + 1u;
    if (hist_len + 1u > F.max_turns + 2u) {
      err |= 2u;
    } else {
      if (lane == 0) hist[hist_len] = key;
      ++hist_len;
    }
    s.rep = (u8)(same > 255u ? 255u : same);
    if (lane == 0 && !err) { R.state = s; R.hist_len = hist_len; }
    return err;
  }
  static __device__ __forceinline__ void init(const ForestView& F, u32 t, int = -1) {
    T::init(F.trees[t].state, F.max_turns);
    F.trees[t].hist_len = 0;
  }
  static __device__ __forceinline__ void info(const ForestView& F, u32 t, u32* turn, u32* rep, u32* player) {
    *turn = F.trees[t].state.turn; *rep = F.trees[t].state.rep; *player = F.trees[t].state.player;
  }
  // the root position as the scheduler sees it (GameData::gs): current_player / current_turn / scores / variant
  static __device__ __forceinline__ u32 root_player(const ForestView& F, u32 t) { return F.trees[t].state.player; }
  static __device__ __forceinline__ u32 root_turn(const ForestView& F, u32 t) { return F.trees[t].state.turn; }
  static __device__ __forceinline__ u32 root_terminal(const ForestView& F, u32 t) { return T::terminal(F.trees[t].state); }
  static __device__ __forceinline__ int root_variant(const ForestView&, u32) { return -1; }
  static __device__ __forceinline__ int pick_variant(const ForestView&, Pcg32&) { return -1; }
};
template <>
struct FGame<B2AZ_FOREST_C4> {  // Connect4 (connect4_gs.cc): two 64-bit stone sets; the root lives in ForestTree::state's
  struct Pos { C4State s; };     // first words (king.lo / king.hi = the players' stones)
  static __device__ __forceinline__ C4State load(const TaflState& t) {
    C4State s; s.p[0] = t.king.lo; s.p[1] = t.king.hi; s.turn = t.turn; s.player = t.player; return s;
  }
  static __device__ __forceinline__ void store(TaflState& t, const C4State& s) {
    t.king.lo = s.p[0]; t.king.hi = s.p[1]; t.turn = s.turn; t.player = (u8)s.player;
  }
  static __device__ __forceinline__ u32 actions(const ForestView&) { return 7u; }
  static __device__ __forceinline__ u32 canon(const ForestView&) { return (u32)C4_CANON; }
  static __device__ __forceinline__ void open(const ForestView& F, u32 t, u32, ForestSmem<B2AZ_FOREST_C4>&, Pos& P) { P.s = load(F.trees[t].state); }
  static __device__ __forceinline__ bool play(Pos& P, u32 mv, u32) { return mv < 7u && c4_play(P.s, mv); }
  static __device__ __forceinline__ u32 player(const Pos& P) { return P.s.player; }
  static __device__ __forceinline__ u32 legal(const Pos& P, ForestSmem<B2AZ_FOREST_C4>& sm, Pcg32& rng, u32 lane, u32*, bool serial) {
    const u32 m = c4_valid_mask(P.s), k = (u32)__popc(m);
    if (lane < 7u && ((m >> lane) & 1u)) sm.moves[__popc(m & ((1u << lane) - 1u))] = (u16)lane;
    __syncwarp();
    forest_shuffle(rng, sm.moves, sm.draws, k, lane, serial);
    return k;
  }
  static __device__ __forceinline__ u32 terminal(const Pos& P, u32) { return c4_terminal(P.s); }
  static __device__ __forceinline__ void emit_canon(const Pos& P, ForestSmem<B2AZ_FOREST_C4>&, float* out, u32 lane) {
    for (u32 e = lane; e < (u32)C4_CANON; e += 32u) out[e] = c4_canon_elem(P.s.p[0], P.s.p[1], P.s.player, e);
  }
  static __device__ __forceinline__ u64 state_key(const Pos& P, const ForestView&) { const u64 h = c4_hash(P.s); return h ? h : 1ULL; }
  static __device__ __forceinline__ u32 stage_floats(const ForestView&) { return (u32)C4_CANON; }
  static __device__ __forceinline__ void stage(const Pos& P, ForestSmem<B2AZ_FOREST_C4>& sm, float* row, u32 lane) { emit_canon(P, sm, row, lane); }
  static __device__ __forceinline__ void unstage(const ForestView&, const float* row, ForestSmem<B2AZ_FOREST_C4>&, float* out, u32 lane) {
    for (u32 e = lane; e < (u32)C4_CANON; e += 32u) out[e] = row[e];
  }
  static __device__ __forceinline__ u32 root_play(const ForestView& F, u32 t, u32 move, ForestSmem<B2AZ_FOREST_C4>&, u32 lane) {
    C4State s = load(F.trees[t].state);
    if (move >= 7u || !c4_play(s, move)) return 8u;
    if (lane == 0) store(F.trees[t].state, s);
    return 0;
  }
  static __device__ __forceinline__ void init(const ForestView& F, u32 t, int = -1) {
    C4State s;
    c4_init(s);
    store(F.trees[t].state, s);
    F.trees[t].hist_len = 0;
  }
  static __device__ __forceinline__ void info(const ForestView& F, u32 t, u32* turn, u32* rep, u32* player) {
    *turn = F.trees[t].state.turn; *rep = 0; *player = F.trees[t].state.player;
  }
  static __device__ __forceinline__ u32 root_player(const ForestView& F, u32 t) { return F.trees[t].state.player; }
  static __device__ __forceinline__ u32 root_turn(const ForestView& F, u32 t) { return F.trees[t].state.turn; }
  static __device__ __forceinline__ u32 root_terminal(const ForestView& F, u32 t) { return c4_terminal(load(F.trees[t].state)); }
  static __device__ __forceinline__ int root_variant(const ForestView&, u32) { return -1; }
  static __device__ __forceinline__ int pick_variant(const ForestView&, Pcg32&) { return -1; }
};
template <>
struct FGame<B2AZ_FOREST_SG> {  // Star Gambit: the variants' own classes and the Unified view
  struct Pos {
    SGState* sp_;  // the warp's working position (shared memory, SGWarpSmem::st)
    SGHistWarp hist;
    SGSpace sp;
    bool unified;
  };
  static __device__ __forceinline__ u32 actions(const ForestView& F) { return F.actions; }
  static __device__ __forceinline__ u32 canon(const ForestView& F) { return F.canon; }
  static __device__ __forceinline__ void open(const ForestView& F, u32 t, u32 lane, ForestSmem<B2AZ_FOREST_SG>& sm, Pos& P) {
    P.sp_ = &sm.st;
    sg_warp_load(sm.st, F.sg_state + t, lane);
    P.unified = sg_game_unified(F.game);
    P.sp = sg_space(sm.st.variant, P.unified);
    P.hist.base = F.sg_hist + (size_t)t * F.sg_hist_cap;
    P.hist.base_len = F.trees[t].hist_len;
    P.hist.keys = F.sg_pkeys + (size_t)t * (kFPath + 2);
    P.hist.len = 0; P.hist.cap = kFPath + 2; P.hist.lane = lane; P.hist.overflow = false;
  }
  static __device__ __forceinline__ bool play(Pos& P, u32 mv, u32 lane) {
    const bool ok = sg_play(*P.sp_, P.hist, P.sp, mv, SGAnyValidWarp{lane}) && !P.hist.overflow;
    __syncwarp();
    return ok;
  }
  static __device__ __forceinline__ u32 player(const Pos& P) { return P.sp_->player; }
  static __device__ __forceinline__ u32 legal(const Pos& P, ForestSmem<B2AZ_FOREST_SG>& sm, Pcg32& rng, u32 lane, u32* err, bool serial) {
    u32 k = sg_warp_legal(*P.sp_, P.sp, sm, lane);
    if (k > (u32)kSGMaxK) { *err |= 4u; k = 0; }
    forest_shuffle(rng, sm.moves, sm.draws, k, lane, serial);
    return k;
  }
  static __device__ __forceinline__ u32 terminal(const Pos& P, u32) { const u32 t = sg_terminal(*P.sp_); return t > 3u ? 3u : t; }
  static __device__ __forceinline__ void emit_canon(const Pos& P, ForestSmem<B2AZ_FOREST_SG>& sm, float* out, u32 lane) {
    sg_warp_canon(*P.sp_, P.hist.count(sg_position_key(*P.sp_)), P.sp, P.unified, sm, lane, out);
  }
  // hash_game_state's equality class (star_gambit_gs.cc:321-338: side to move, has_taken_action, every unit's nine
  // fields, the reserves; the Unified view adds the variant, 2399-2402) as a 64-bit key; 0 is reserved
  static __device__ __forceinline__ u64 state_key(const Pos& P, const ForestView&) {
    const SGState& s = *P.sp_;
    u64 h = 0x9E3779B97F4A7C15ULL ^ ((u64)s.player | ((u64)s.acted << 8) | ((u64)(P.unified ? s.variant + 1u : 0u) << 16));
    const u8* b = reinterpret_cast<const u8*>(s.units);
    const int nb = (int)s.n_units * (int)sizeof(SGUnit);
    for (int i = 0; i < nb; i += 8) {
      u64 w = 0;
      for (int j = 0; j < 8 && i + j < nb; ++j) w |= (u64)b[i + j] << (8 * j);
      h = (h ^ w) * 0xBF58476D1CE4E5B9ULL; h ^= h >> 29;
    }
    u64 r = 0;
    for (int i = 0; i < 8; ++i) r |= (u64)(&s.reserves[0][0])[i] << (8 * i);
    h = (h ^ r) * 0xBF58476D1CE4E5B9ULL; h ^= h >> 29;
    h = (h ^ (h >> 32)) * 0x94D049BB133111EBULL;
    h ^= h >> 31;
    return h ? h : 1ULL;
  }
  // A training sample's position while its game is still running (PlayManager's partial_history): the 200-byte
  // position record + its repetition count instead of the 24 KB of canonical planes, which are written once, into the
  // output ring, when the game ends.
  static __device__ __forceinline__ u32 stage_floats(const ForestView&) { return (u32)(sizeof(SGState) / 4u + 1u); }
  static __device__ __forceinline__ void stage(const Pos& P, ForestSmem<B2AZ_FOREST_SG>&, float* row, u32 lane) {
    const int rc = P.hist.count(sg_position_key(*P.sp_));
    u32* w = reinterpret_cast<u32*>(row);
    const u32* src = reinterpret_cast<const u32*>(P.sp_);
    for (u32 i = lane; i < (u32)(sizeof(SGState) / 4u); i += 32u) w[i] = src[i];
    if (lane == 0) w[sizeof(SGState) / 4u] = (u32)rc;
  }
  static __device__ __forceinline__ void unstage(const ForestView& F, const float* row, ForestSmem<B2AZ_FOREST_SG>& sm, float* out, u32 lane) {
    const u32* w = reinterpret_cast<const u32*>(row);
    SGState& s = sm.st;
    sg_warp_load(s, reinterpret_cast<const SGState*>(row), lane);
    const bool unified = sg_game_unified(F.game);
    sg_warp_canon(s, (int)w[sizeof(SGState) / 4u], sg_space(s.variant, unified), unified, sm, lane, out);
  }
  static __device__ __forceinline__ u32 root_play(const ForestView& F, u32 t, u32 move, ForestSmem<B2AZ_FOREST_SG>& sm, u32 lane) {
    ForestTree& R = F.trees[t];
    SGState& s = sm.st;
    sg_warp_load(s, F.sg_state + t, lane);
    const bool unified = sg_game_unified(F.game);
    const SGSpace sp = sg_space(s.variant, unified);
    SGHistWarp h;
    h.base = nullptr; h.base_len = 0; h.keys = F.sg_hist + (size_t)t * F.sg_hist_cap; h.len = R.hist_len; h.cap = F.sg_hist_cap;
    h.lane = lane; h.overflow = false;
    const bool ok = sg_play(s, h, sp, move, SGAnyValidWarp{lane});
    __syncwarp();
    if (!ok) return 8u;
    if (h.overflow) return 2u;
    sg_warp_store(F.sg_state + t, s, lane);
    if (lane == 0) R.hist_len = h.len;
    return 0;
  }
  // (one thread per tree) `variant` < 0: the game id's own; game 24 (the variant mix) without a pick yet: by tree pair
  static __device__ __forceinline__ void init(const ForestView& F, u32 t, int variant = -1) {
    SGState s;
    if (variant < 0) variant = F.game == 24u ? (int)((t >> 1) & 3u) : sg_game_variant(F.game);
    sg_init(s, variant);
    F.sg_state[t] = s;
    F.sg_hist[(size_t)t * F.sg_hist_cap] = sg_position_key(s);
    F.trees[t].hist_len = 1;
  }
  static __device__ __forceinline__ void info(const ForestView& F, u32 t, u32* turn, u32* rep, u32* player) {
    *turn = F.sg_state[t].turn; *rep = 0; *player = F.sg_state[t].player;
  }
  static __device__ __forceinline__ u32 root_player(const ForestView& F, u32 t) { return F.sg_state[t].player; }
  static __device__ __forceinline__ u32 root_turn(const ForestView& F, u32 t) { return F.sg_state[t].turn; }
  static __device__ __forceinline__ u32 root_terminal(const ForestView& F, u32 t) {
    const SGState& s = F.sg_state[t];
    if (!s.over) return 0u;
    return s.winner == 0 ? 1u : s.winner == 1 ? 2u : 3u;
  }
  static __device__ __forceinline__ int root_variant(const ForestView& F, u32 t) {  // get_variant_id(): Unified only
    return sg_game_unified(F.game) ? (int)F.sg_state[t].variant : -1;
  }
  // StarGambitUnifiedGS::randomize_start (star_gambit_gs.cc:2421-2425): a variant drawn from the weights. The reference
  // draws from an unseedable mt19937 (2357-2362); here the slot's coin stream decides. -1: the game id fixes the variant.
  static __device__ __forceinline__ int pick_variant(const ForestView& F, Pcg32& coin) {
    if (F.game != 24u) return -1;
    const float total = fadd(fadd(F.sg_probs[0], F.sg_probs[1]), fadd(F.sg_probs[2], F.sg_probs[3]));
    const float u = fmul(rng_uniform01(coin), total);
    float acc = 0.0f;
    for (int v = 0; v < 3; ++v) {
      acc = fadd(acc, F.sg_probs[v]);
      if (u < acc) return v;
    }
    return 3;
  }
};

// ---- Gumbel root search over a wide root (mcts.cc:28-66, 175-283, 336-401). Root-only and a few hundred scalar
// steps per move, so it runs on lane 0; the formulas and their float order are those of the Connect4 engine's
// (az_engine_logic.h gumbel_*), with arrays in HBM instead of nibble-packed registers.
#define FG_LOG_FLOOR 1e-20f
__device__ __forceinline__ void fg_reset(ForestGumbel& G) {  // reset_gumbel_state (mcts.cc:180-188)
  G.initialized = 0; G.effective_m = 0; G.n_surv = 0; G.n_phases = 0; G.phase_idx = 0; G.sims_in_phase = 0;
}
__device__ __noinline__ void fg_phase_plan(ForestGumbel& G, u32 m, u32 n) {  // seq_halving_phase_plan (mcts.cc:28-66)
  G.n_phases = 0;
  if (m <= 1) { G.phase_numc[0] = 1; G.phase_vper[0] = n; G.n_phases = 1; return; }
  u32 log2m = 0;
  for (u32 v = m - 1; v > 0; v >>= 1) ++log2m;
  if (log2m == 0) log2m = 1;
  u32 base_v = n / (log2m * m);
  if (base_v < 1u) base_v = 1u;
  u32 sims_used = 0, num_c = m;
  for (u32 ph = 0; ph < log2m && G.n_phases < (u32)kFMaxPhases; ++ph) {
    if (sims_used >= n) break;
    const u32 remaining = n - sims_used;
    const bool is_final = (ph == log2m - 1);
    u32 v_per = is_final ? (remaining / num_c < 1u ? 1u : remaining / num_c) : base_v * (1u << ph);
    if (num_c * v_per > remaining) {
      v_per = remaining / num_c;
      if (v_per == 0) { num_c = remaining; v_per = 1; }
    }
    G.phase_numc[G.n_phases] = num_c;
    G.phase_vper[G.n_phases] = v_per;
    ++G.n_phases;
    sims_used += num_c * v_per;
    num_c = num_c / 2 < 1u ? 1u : num_c / 2;
  }
}
// Top-`take` of `cnt` scores, descending (std::partial_sort in the reference: with continuous Gumbel noise in every
// score ties have probability zero, so a plain selection gives the same ranking). `ids[i]` names score i.
__device__ __noinline__ void fg_rank_top(const float* score, const u16* ids, u32 cnt, u32 take, u16* out) {
  u32 used[kFMaxK / 32];
  for (int i = 0; i < kFMaxK / 32; ++i) used[i] = 0;
  for (u32 r = 0; r < take; ++r) {
    int best = -1;
    float bs = 0.0f;
    for (u32 i = 0; i < cnt; ++i) {
      if ((used[i >> 5] >> (i & 31u)) & 1u) continue;
      const float sc = score[i];
      if (best < 0 || sc > bs) { best = (int)i; bs = sc; }
    }
    used[best >> 5] |= 1u << (best & 31);
    out[r] = ids ? ids[best] : (u16)best;
  }
}
__device__ __forceinline__ float fg_sigma_scale(const ForestView& F, u32 t, u32 max_visit) {
  return fmul(fadd(FSEAT(F, t).gumbel_c_visit, (float)max_visit), FSEAT(F, t).gumbel_c_scale);
}
// init_gumbel_state (mcts.cc:190-227); lane 0 only
__device__ __noinline__ void fg_init(const ForestView& F, u32 t, ForestTree& R, ForestGumbel& G, const u32* pool) {
  const u32 num_legal = R.k, b = R.blk;
  if (num_legal == 0 || b == 0) return;
  const u32 remaining = R.depth < G.num_sims_target ? G.num_sims_target - R.depth : 0u;
  if (remaining == 0) return;
  u32 m = FSEAT(F, t).gumbel_m < num_legal ? FSEAT(F, t).gumbel_m : num_legal;
  if (remaining < m) m = remaining;
  if (m < 1u) m = 1u;
  G.effective_m = m;
  float* g = F.gum_g + (size_t)t * (2 * kFMaxK);
  float* score = g + kFMaxK;
  Pcg32 rng = FOREST_RNG(F, t);
  for (u32 i = 0; i < num_legal; ++i) g[i] = rng_gumbel(rng);
  FOREST_RNG(F, t) = rng;
  for (u32 i = 0; i < num_legal; ++i) score[i] = fadd(g[i], az_logf(fadd(u2f(pool[fb_pol(b, num_legal) + i]), FG_LOG_FLOOR)));
  fg_rank_top(score, nullptr, num_legal, m, G.survivors);
  G.n_surv = m;
  fg_phase_plan(G, m, remaining);
  G.phase_idx = 0;
  G.sims_in_phase = 0;
  G.initialized = 1;
}
// gumbel_advance_phase (mcts.cc:229-264)
__device__ __noinline__ void fg_advance_phase(const ForestView& F, u32 t, const ForestTree& R, ForestGumbel& G, const u32* pool) {
  if (G.phase_idx + 1u >= G.n_phases) return;
  const u32 next_num_c = G.phase_numc[G.phase_idx + 1];
  if (next_num_c >= G.n_surv) { ++G.phase_idx; G.sims_in_phase = 0; return; }
  const u32 b = R.blk, k = R.k;
  const float* g = F.gum_g + (size_t)t * (2 * kFMaxK);
  u32 max_visit = 0;
  for (u32 r = 0; r < G.n_surv; ++r) {
    const u32 n = pool[fb_n(b, k) + G.survivors[r]];
    if (n > max_visit) max_visit = n;
  }
  const float sigma_scale = fg_sigma_scale(F, t, max_visit);
  float score[kFMaxM];
  u16 ids[kFMaxM];
  for (u32 r = 0; r < G.n_surv; ++r) {
    const u32 c = G.survivors[r];
    const float logit = az_logf(fadd(u2f(pool[fb_pol(b, k) + c]), FG_LOG_FLOOR));
    const float q_hat = pool[fb_n(b, k) + c] > 0 ? u2f(pool[fb_q(b, k) + c]) : 0.0f;
    score[r] = fadd(fadd(g[c], logit), fmul(sigma_scale, q_hat));
    ids[r] = (u16)c;
  }
  u16 next[kFMaxM];
  fg_rank_top(score, ids, G.n_surv, next_num_c, next);
  for (u32 r = 0; r < next_num_c; ++r) G.survivors[r] = next[r];
  G.n_surv = next_num_c;
  ++G.phase_idx;
  G.sims_in_phase = 0;
}
// gumbel_next_root_child (mcts.cc:266-283)
__device__ __noinline__ u32 fg_next_root_child(const ForestView& F, u32 t, const ForestTree& R, ForestGumbel& G, const u32* pool) {
  if (G.phase_idx < G.n_phases) {
    if (G.sims_in_phase >= G.phase_numc[G.phase_idx] * G.phase_vper[G.phase_idx]) fg_advance_phase(F, t, R, G, pool);
  }
  if (G.n_surv == 0) return 0;
  const u32 pick = G.sims_in_phase % G.n_surv;
  ++G.sims_in_phase;
  return G.survivors[pick];
}
// gumbel_final_action (mcts.cc:375-401) when the search initialised; 0xFFFFFFFF otherwise (the reference then falls
// back to pick_move(probs(0)), which the caller does from the counts)
__device__ __noinline__ u32 fg_final_action(const ForestView& F, u32 t, const ForestTree& R, const ForestGumbel& G, const u32* pool) {
  if (!G.initialized || G.n_surv == 0 || R.blk == 0) return 0xFFFFFFFFu;
  const u32 b = R.blk, k = R.k;
  const float* g = F.gum_g + (size_t)t * (2 * kFMaxK);
  u32 max_visit = 0;
  for (u32 i = 0; i < k; ++i) {
    const u32 n = pool[fb_n(b, k) + i];
    if (n > max_visit) max_visit = n;
  }
  const float sigma_scale = fg_sigma_scale(F, t, max_visit);
  u32 best = G.survivors[0];
  float best_score = -INFINITY;
  for (u32 r = 0; r < G.n_surv; ++r) {
    const u32 c = G.survivors[r];
    const float logit = az_logf(fadd(u2f(pool[fb_pol(b, k) + c]), FG_LOG_FLOOR));
    const float q_hat = pool[fb_n(b, k) + c] > 0 ? u2f(pool[fb_q(b, k) + c]) : 0.0f;
    const float score = fadd(fadd(g[c], logit), fmul(sigma_scale, q_hat));
    if (score > best_score) { best_score = score; best = c; }
  }
  return pool[fb_mv(b, k) + best] & 0xFFFFu;
}
// gumbel_improved_policy (mcts.cc:336-373) incl. compute_v_mix_from_children (mcts.cc:71-89): pi'[move] for the
// root's children into out[A] (already zeroed); z lives in the tree's scratch row. lane 0 only.
__device__ __noinline__ void fg_improved_policy(const ForestView& F, u32 t, const ForestTree& R, const u32* pool, float* out) {
  const u32 b = R.blk, k = R.k;
  if (k == 0 || b == 0) return;
  float* z = F.gum_g + (size_t)t * (2 * kFMaxK) + kFMaxK;
  u32 max_visit = 0;
  float sum_visits = 0.0f, sum_priors_visited = 0.0f, weighted_num = 0.0f;
  for (u32 i = 0; i < k; ++i) {
    const u32 n = pool[fb_n(b, k) + i];
    if (n > max_visit) max_visit = n;
    sum_visits = fadd(sum_visits, (float)n);
    if (n > 0) {
      const float p = u2f(pool[fb_pol(b, k) + i]);
      sum_priors_visited = fadd(sum_priors_visited, p);
      weighted_num = fadd(weighted_num, fmul(p, u2f(pool[fb_q(b, k) + i])));
    }
  }
  float v_mix = R.v;
  if (sum_priors_visited > 0.0f) {
    const float weighted_q = fdiv(weighted_num, sum_priors_visited);
    v_mix = fdiv(fadd(R.v, fmul(sum_visits, weighted_q)), fadd(sum_visits, 1.0f));
  }
  const float sigma_scale = fg_sigma_scale(F, t, max_visit);
  float z_max = -INFINITY;
  for (u32 i = 0; i < k; ++i) {
    const float completed_q = pool[fb_n(b, k) + i] > 0 ? u2f(pool[fb_q(b, k) + i]) : v_mix;
    z[i] = fadd(az_logf(fadd(u2f(pool[fb_pol(b, k) + i]), FG_LOG_FLOOR)), fmul(sigma_scale, completed_q));
    if (z[i] > z_max) z_max = z[i];
  }
  float z_sum = 0.0f;
  for (u32 i = 0; i < k; ++i) {
    z[i] = az_expf(fsub(z[i], z_max));
    z_sum = fadd(z_sum, z[i]);
  }
  if (z_sum <= 0.0f) return;
  for (u32 i = 0; i < k; ++i) out[pool[fb_mv(b, k) + i] & 0xFFFFu] = fdiv(z[i], z_sum);
}

// MCTS::gumbel_interior_select (mcts.cc:285-334; gumbel_full): argmax_a [ pi'(a) - N(a) / (1 + sum N) ] with
// pi' = softmax(log prior + sigma * completedQ) at an INTERIOR node whose children live in block b; `node_v` is the
// node's stored value. Sequential float order of the reference; lane 0; the tree's Gumbel scratch row holds z.
__device__ __noinline__ u32 fg_interior_select(const ForestView& F, u32 t, const u32* pool, u32 b, u32 k, float node_v) {
  float* z = F.gum_g + (size_t)t * (2 * kFMaxK) + kFMaxK;
  u32 max_visit = 0, sum_n = 0;
  float sum_visits = 0.0f, sum_priors_visited = 0.0f, weighted_num = 0.0f;
  for (u32 i = 0; i < k; ++i) {
    const u32 n = pool[fb_n(b, k) + i];
    if (n > max_visit) max_visit = n;
    sum_n += n;
    sum_visits = fadd(sum_visits, (float)n);
    if (n > 0) {
      const float p = u2f(pool[fb_pol(b, k) + i]);
      sum_priors_visited = fadd(sum_priors_visited, p);
      weighted_num = fadd(weighted_num, fmul(p, u2f(pool[fb_q(b, k) + i])));
    }
  }
  float v_mix = node_v;
  if (sum_priors_visited > 0.0f) {
    const float weighted_q = fdiv(weighted_num, sum_priors_visited);
    v_mix = fdiv(fadd(node_v, fmul(sum_visits, weighted_q)), fadd(sum_visits, 1.0f));
  }
  const float sigma_scale = fg_sigma_scale(F, t, max_visit);
  float z_max = -INFINITY;
  for (u32 i = 0; i < k; ++i) {
    const float completed_q = pool[fb_n(b, k) + i] > 0 ? u2f(pool[fb_q(b, k) + i]) : v_mix;
    z[i] = fadd(az_logf(fadd(u2f(pool[fb_pol(b, k) + i]), FG_LOG_FLOOR)), fmul(sigma_scale, completed_q));
    if (z[i] > z_max) z_max = z[i];
  }
  float z_sum = 0.0f;
  for (u32 i = 0; i < k; ++i) {
    z[i] = az_expf(fsub(z[i], z_max));
    z_sum = fadd(z_sum, z[i]);
  }
  const float inv = z_sum > 0.0f ? fdiv(1.0f, z_sum) : 0.0f;
  const float denom = fadd(1.0f, (float)sum_n);
  u32 best = 0;
  float best_score = -INFINITY;
  for (u32 i = 0; i < k; ++i) {
    const float score = fsub(fmul(z[i], inv), fdiv((float)pool[fb_n(b, k) + i], denom));
    if (score > best_score) { best_score = score; best = i; }
  }
  return best;
}

// ---- root policy temperature and Dirichlet noise over a wide root (mcts.cc:403-460); lane 0, policy array in HBM.
// Same formulas / float order as the Connect4 engine's add_root_noise (az_engine_logic.h); the noise values live
// in the tree's Gumbel scratch row (never needed at the same time: Gumbel replaces the noise, mcts.cc:514-518).
__device__ __noinline__ void fr_apply_root_policy_temp(const ForestView& F, u32 t, u32* pool, u32 b, u32 k) {  // mcts.cc:448-460
  const float root_temp = FSEAT(F, t).root_policy_temp;
  if (root_temp == 1.0f || b == 0) return;
  const float e = fdiv(1.0f, root_temp);
  float sum = 0.0f;
  for (u32 j = 0; j < k; ++j) {
    const float p = az_powf(u2f(pool[fb_pol(b, k) + j]), e);
    pool[fb_pol(b, k) + j] = f2u(p);
    sum = fadd(sum, p);
  }
  if (sum > 0.0f)
    for (u32 j = 0; j < k; ++j) pool[fb_pol(b, k) + j] = f2u(fdiv(u2f(pool[fb_pol(b, k) + j]), sum));
}
__device__ __noinline__ void fr_add_root_noise(const ForestView& F, u32 t, Pcg32& rng, u32* pool, u32 b, u32 k) {  // mcts.cc:403-446
  if (b == 0 || k == 0) return;
  float* noise = F.noise + (size_t)t * kFMaxK;
  double sum = 0.0;
  if (F.shaped_dirichlet && k > 1) {
    const float N = (float)k;
    float log_sum = 0.0f;
    for (u32 j = 0; j < k; ++j) log_sum = fadd(log_sum, az_logf(fadd(std_min(u2f(pool[fb_pol(b, k) + j]), 0.01f), 1e-20f)));
    const float log_mean = fdiv(log_sum, N);
    float shaped_sum = 0.0f;
    for (u32 j = 0; j < k; ++j) {
      const float lp = az_logf(fadd(std_min(u2f(pool[fb_pol(b, k) + j]), 0.01f), 1e-20f));
      shaped_sum = fadd(shaped_sum, std_max(0.0f, fsub(lp, log_mean)));
    }
    const float uniform = fdiv(1.0f, N);
    for (u32 j = 0; j < k; ++j) {
      const float lp = az_logf(fadd(std_min(u2f(pool[fb_pol(b, k) + j]), 0.01f), 1e-20f));
      const float shaped = std_max(0.0f, fsub(lp, log_mean));
      float alpha_prop = (shaped_sum > 0.0f) ? fmul(0.5f, fadd(fdiv(shaped, shaped_sum), uniform)) : uniform;
      alpha_prop = std_max(alpha_prop, 1e-6f);
      GammaDist gd;
      gamma_init(gd, fmul(10.83f, alpha_prop));
      noise[j] = gamma_draw(rng, gd);
      sum = dadd(sum, (double)noise[j]);
    }
  } else {
    GammaDist gd;
    gamma_init(gd, fdiv(10.83f, (float)k));
    for (u32 j = 0; j < k; ++j) {
      noise[j] = gamma_draw(rng, gd);
      sum = dadd(sum, (double)noise[j]);
    }
  }
  const float fsum = (float)sum;
  const float eps = FSEAT(F, t).epsilon;
  const float keep = fsub(1.0f, eps);
  for (u32 j = 0; j < k; ++j)
    pool[fb_pol(b, k) + j] = f2u(fadd(fmul(u2f(pool[fb_pol(b, k) + j]), keep), fdiv(fmul(eps, noise[j]), fsum)));
}

// MCTS::find_leaf (mcts.cc:462-498) for tree t
// BATCHED = MCTS::find_leaf_batched (mcts.cc:752-789, WU-UCT): the descent may pass nodes that a previous in-flight
// call expanded but that have no visit yet, every node on the path (and the leaf) gets ++n_in_flight AFTER the
// selection made at it, and the leaf goes to the tree's in-flight list instead of current_/path_.
// LOCK: the CTA's warps walk their trees in lock step — one CTA-wide vote per tree level, so that all of them run the
// same stretch of code at the same time (the kernels are 110-200 KB of SASS against a 32 KB instruction cache per SM:
// warps scattered over the code starve on instruction fetch, profiles/r3j_k_sp_search_sg_ncu_summary.json). A warp
// without a simulation to do (`live` false) only takes part in the votes.
template <int GAME, bool BATCHED, bool LOCK = false>
__device__ void forest_find_leaf(const ForestView& F, u32 t, ForestSmem<GAME>& sm, u32 lane, bool emit_canon,
                                 ForestLeaf& Lf, float* canon_out, bool live = true, u64* key_out = nullptr) {
  typedef FGame<GAME> G;
  if (LOCK && !live) {
    while (__syncthreads_or(0)) {}
    return;
  }
  ForestTree& R = F.trees[t];
  u32* pool = F.pool + (size_t)t * F.words_per_tree;
  typename G::Pos pos;
  G::open(F, t, lane, sm, pos);
  u32 err = 0;
  // current_ = &root_
  u32 cur_n = R.n, cur_term = R.term, cur_blk = R.blk, cur_k = R.k, cur_player = R.player;
  u32 cur_nif = R.nif, cur_expanded = R.expanded;
  float cur_v = R.v;
  u32 par_blk = 0, par_slot = 0, par_k = 0;  // where the current node's own fields live (0 = it is the root)
  u32 plen = 0;
  bool at_root = true;
  // lazy Gumbel init (mcts.cc:468-472): once the root is expanded and a sims target is set
  bool gumbel_on = false;
  const u32 sflags = fseat_flags(F, t);  // once per descent
  if (sflags & 0xFF00u) {
    ForestGumbel& G = F.gum[t];
    if (lane == 0 && !G.initialized && G.num_sims_target > 0 && R.n > 0 && R.k > 0 && R.blk != 0) fg_init(F, t, R, G, pool);
    __syncwarp();
    gumbel_on = G.initialized != 0;
  }
  bool descending = true;
  for (;;) {
    bool go = descending && (BATCHED ? ((cur_n > 0 || cur_nif > 0) && cur_blk != 0) : (cur_n > 0)) && cur_term == 0;
    if (go && (plen >= (u32)kFPath || cur_blk == 0)) { err |= 2u; go = false; }
    if (LOCK) { if (!__syncthreads_or(go ? 1 : 0)) break; } else if (!go) break;
    if (!go) { descending = false; continue; }
    const u32 b = cur_blk, k = cur_k;
    u32 best_j = 0xFFFFFFFFu;
    if (gumbel_on && at_root) {  // the root child comes from the sequential-halving schedule (mcts.cc:476-478)
      u32 forced = 0;
      if (lane == 0) forced = fg_next_root_child(F, t, R, F.gum[t], pool);
      best_j = __shfl_sync(0xFFFFFFFFu, forced, 0);
    } else if (gumbel_on && (sflags & 0xFF0000u)) {  // pi'-matching below the root as well (mcts.cc:479-481)
      u32 sel = 0;
      if (lane == 0) sel = fg_interior_select(F, t, pool, b, k, cur_v);
      best_j = __shfl_sync(0xFFFFFFFFu, sel, 0);
    } else {
    // Node::best_child (mcts.cc:130-149)
    float seen = 0.0f;
    for (u32 c0 = 0; c0 < k; c0 += 32u) {
      const u32 j = c0 + lane;
      const u32 nj = j < k ? pool[fb_n(b, k) + j] : 0u;
      const float pj = j < k ? u2f(pool[fb_pol(b, k) + j]) : 0.0f;
      seen = seq_sum_masked(seen, pj, j < k && nj > 0);  // usually a handful of visited children
    }
    const float fpu = (at_root && (sflags & 0xFFu)) ? 0.0f : F.fpu_reduction;
    const float fpu_value = fsub(cur_v, fmul(fpu, fsqrt(seen)));
    const float sqrt_n = fsqrt((float)(cur_n + (BATCHED ? cur_nif : 0u)));  // sqrt(n + n_in_flight) (mcts.cc:138)
    float best_u = 0.0f;
    for (u32 c0 = 0; c0 < k; c0 += 32u) {
      const u32 j = c0 + lane;
      if (j < k) {
        const u32 nj = pool[fb_n(b, k) + j];
        const u32 fj = BATCHED ? pool[fb_nif(b, k) + j] : 0u;  // Node::uct: n + n_in_flight + 1 (mcts.cc:125-127)
        const float qj = u2f(pool[fb_q(b, k) + j]), pj = u2f(pool[fb_pol(b, k) + j]);
        const float u = fadd(nj == 0 ? fpu_value : qj, fdiv(fmul(fmul(F.cpuct, pj), sqrt_n), (float)(nj + fj + 1u)));
        if (best_j == 0xFFFFFFFFu || u > best_u) { best_u = u; best_j = j; }  // per lane: ascending j, strict >
      }
    }
    // first maximum over the lanes: larger u wins, equal u -> lower index (child 0 is the reference's start value)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ou = __shfl_xor_sync(0xFFFFFFFFu, best_u, o);
      const u32 oj = __shfl_xor_sync(0xFFFFFFFFu, best_j, o);
      const bool take = oj != 0xFFFFFFFFu && (best_j == 0xFFFFFFFFu || ou > best_u || (ou == best_u && oj < best_j));
      if (take) { best_u = ou; best_j = oj; }
    }
    if (best_j == 0xFFFFFFFFu) best_j = 0;  // (all scores NaN: the reference keeps child 0)
    }
    // NaN scores lose every comparison in the reference loop as well, except at index 0, which is only replaced by
    // a strictly greater score; with finite scores both orders agree.
    const u32 mvw = pool[fb_mv(b, k) + best_j];
    if (lane == 0) {
      Lf.path_blk[plen] = b;
      Lf.path_slot[plen] = (u16)best_j;
      Lf.path_player[plen] = (u8)cur_player;
      if (BATCHED) {  // ++cur->n_in_flight, after the selection made at cur
        if (at_root) R.nif = cur_nif + 1u;
        else pool[fb_nif(par_blk, par_k) + par_slot] = cur_nif + 1u;
      }
    }
    ++plen;
    if (!G::play(pos, mvw & 0xFFFFu, lane)) { err |= 8u; descending = false; continue; }
    par_blk = b; par_slot = best_j; par_k = k;
    cur_n = pool[fb_n(b, k) + best_j];
    cur_v = u2f(pool[fb_v(b, k) + best_j]);
    cur_blk = pool[fb_fc(b, k) + best_j];
    cur_k = cur_blk ? pool[cur_blk] : 0u;
    cur_player = (mvw >> 16) & 3u;
    cur_term = (mvw >> 20) & 3u;
    cur_expanded = (mvw >> 22) & 1u;
    cur_nif = BATCHED ? pool[fb_nif(b, k) + best_j] : 0u;
    at_root = false;
  }
  if (BATCHED && lane == 0 && !err) {  // ++cur->n_in_flight on the leaf itself
    if (plen == 0) R.nif = cur_nif + 1u;
    else pool[fb_nif(par_blk, par_k) + par_slot] = cur_nif + 1u;
  }
  u32 leaf_new = 0, leaf_blk = cur_blk, leaf_k = cur_k, leaf_term = cur_term, leaf_player = cur_player;
  // (batched: only if a prior in-flight call has not expanded it already, mcts.cc:776-783)
  if (cur_n == 0 && !err && !(BATCHED && cur_expanded)) {
    // current_->player, scores, add_children(valid_moves) incl. the shuffle (mcts.cc:490-496)
    leaf_new = 1;
    leaf_player = G::player(pos);
    Pcg32 rng = FOREST_RNG(F, t);
    const u32 k = G::legal(pos, sm, rng, lane, &err, F.serial_shuffle != 0);
    leaf_term = G::terminal(pos, k);
    // children of a terminal node are never visited: their draws are consumed above, their storage is skipped
    leaf_k = leaf_term ? 0u : k;
    leaf_blk = 0;
    if (leaf_k) {
      const u32 need = fb_words(leaf_k);
      const u32 b = R.bump;
      if (b + need > (R.half ? F.words_per_tree : F.words_per_tree / 2u)) {
        err |= 1u;
        leaf_k = 0;
      } else {
        leaf_blk = b;
        if (lane == 0) pool[b] = leaf_k;
        for (u32 i = lane; i < 6u * leaf_k; i += 32u) pool[b + 1u + i] = 0u;      // n q policy d v first_child
        for (u32 j = lane; j < leaf_k; j += 32u) pool[fb_nif(b, leaf_k) + j] = 0u;
        for (u32 j = lane; j < leaf_k; j += 32u) pool[fb_mv(b, leaf_k) + j] = sm.moves[j];
        if (lane == 0) R.bump = b + need;
      }
    }
    if (lane == 0) {
      FOREST_RNG(F, t) = rng;
      if (par_blk == 0 && plen == 0) {
        R.player = leaf_player; R.term = leaf_term; R.blk = leaf_blk; R.k = leaf_k; R.expanded = 1;
      } else {
        pool[fb_fc(par_blk, par_k) + par_slot] = leaf_blk;
        pool[fb_mv(par_blk, par_k) + par_slot] |= (leaf_player << 16) | (leaf_term << 20) | (1u << 22);
      }
    }
  }
  if (emit_canon) G::emit_canon(pos, sm, canon_out, lane);
  if (key_out) *key_out = G::state_key(pos, F);  // the leaf position's cache key (hash_game_state)
  if (lane == 0) {
    R.total_leaf_depth += plen;
    Lf.path_len = plen;
    Lf.leaf_blk = leaf_blk; Lf.leaf_k = leaf_k; Lf.leaf_term = leaf_term; Lf.leaf_player = leaf_player; Lf.leaf_new = leaf_new;
    if (err) R.error |= err;
  }
  __syncwarp();
}

// MCTS::process_result (mcts.cc:500-555) for tree t. RANDOM = dumb_eval (game_state.h:160-173) instead of (v, pi).
// BATCHED = MCTS::process_result_batched (mcts.cc:791-845): the same with --n_in_flight along the path.
template <int GAME, bool RANDOM, bool BATCHED>
__device__ void forest_process_result(const ForestView& F, u32 t, const float* ev_v, const float* ev_pi, u32 lane,
                                      bool root_noise_enabled, ForestLeaf& Lf, u32 row = 0xFFFFFFFFu) {
  const size_t er = row == 0xFFFFFFFFu ? (size_t)t : (size_t)row;  // the evaluator's row of this tree (self-play: the game slot)
  const u32 A = FGame<GAME>::actions(F);
  ForestTree& R = F.trees[t];
  u32* pool = F.pool + (size_t)t * F.words_per_tree;
  float val0, val1, vald;
  const u32 lterm = Lf.leaf_term, lk = Lf.leaf_k, lblk = Lf.leaf_blk, lplayer = Lf.leaf_player, plen = Lf.path_len;
  if (lterm != 0) {
    val0 = lterm == 1 ? 1.0f : 0.0f; val1 = lterm == 2 ? 1.0f : 0.0f; vald = lterm == 3 ? 1.0f : 0.0f;
  } else {
    if (RANDOM) {
      val0 = val1 = vald = (float)(1.0 / 3.0);
    } else {
      val0 = ev_v[er * 3 + 0]; val1 = ev_v[er * 3 + 1]; vald = ev_v[er * 3 + 2];
      // relative_to_absolute(value, current_->player, 2) (mcts.cc:522-524, game_state.h:37-48): seat 1's answer swaps
      if (F.relative_values && lplayer == 1u) { const float sw = val0; val0 = val1; val1 = sw; }
    }
    if (lk > 0 && lblk != 0) {
      // set_policy_normalized (mcts.cc:109-121); the same code for the root and interior nodes here
      // (root_policy_temp == 1, no noise)
      float rp = 0.0f;
      if (RANDOM) {  // valids.cast<float>() / float(Vector<uint8_t>::sum()) — the uint8 sum wraps mod 256
        const u32 s8 = lk & 255u;
        rp = s8 ? fdiv(1.0f, (float)s8) : 0.0f;
      }
      float sum = 0.0f;
      for (u32 c0 = 0; c0 < lk; c0 += 32u) {
        const u32 j = c0 + lane;
        float p = 0.0f;
        if (j < lk) {
          p = RANDOM ? rp : ev_pi[er * A + (pool[fb_mv(lblk, lk) + j] & 0xFFFFu)];
          pool[fb_pol(lblk, lk) + j] = f2u(p);
        }
        const u32 cnt = lk - c0 < 32u ? lk - c0 : 32u;
        if (RANDOM) {  // all priors equal: the same in-order sum without the shuffles
          for (u32 q = 0; q < cnt; ++q) sum = fadd(sum, rp);
        } else {
          sum = seq_sum_all(sum, p, cnt);
        }
      }
      __syncwarp();
      const bool is_root = plen == 0;
      if (is_root && FSEAT(F, t).root_policy_temp != 1.0f) {
        // set_policy_normalized(pi, apply_temp = true, 1 / T): every prior is raised to 1/T BEFORE the in-order sum
        // (mcts.cc:111-120); az_powf is out of line and scalar: lane 0 redoes the root's priors serially
        if (lane == 0) {
          const float e = fdiv(1.0f, FSEAT(F, t).root_policy_temp);
          float tsum = 0.0f;
          for (u32 j = 0; j < lk; ++j) {
            const float p = az_powf(u2f(pool[fb_pol(lblk, lk) + j]), e);
            pool[fb_pol(lblk, lk) + j] = f2u(p);
            tsum = fadd(tsum, p);
          }
          for (u32 j = 0; j < lk; ++j) pool[fb_pol(lblk, lk) + j] = f2u(fdiv(u2f(pool[fb_pol(lblk, lk) + j]), tsum));
        }
      } else {
        for (u32 j = lane; j < lk; j += 32u) pool[fb_pol(lblk, lk) + j] = f2u(fdiv(u2f(pool[fb_pol(lblk, lk) + j]), sum));
      }
      __syncwarp();
      // Gumbel replaces Dirichlet noise (mcts.cc:514-518)
      if (is_root && root_noise_enabled && FSEAT(F, t).epsilon > 0.0f && !FSEAT(F, t).gumbel_enabled && lane == 0) {
        Pcg32 rng = FOREST_RNG(F, t);
        fr_add_root_noise(F, t, rng, pool, lblk, lk);
        FOREST_RNG(F, t) = rng;
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    const float dshare = fdiv(vald, 2.0f);
    if (BATCHED) {  // --current_->n_in_flight on the leaf, --parent->n_in_flight on every node of the path
      if (plen == 0) --R.nif;
      else { const u32 b = Lf.path_blk[plen - 1u], k = pool[b]; --pool[fb_nif(b, k) + Lf.path_slot[plen - 1u]]; }
      for (u32 i = 0; i < plen; ++i) {
        if (i == 0) --R.nif;
        else { const u32 b = Lf.path_blk[i - 1u], k = pool[b]; --pool[fb_nif(b, k) + Lf.path_slot[i - 1u]]; }
      }
    }
    for (u32 i = plen; i-- > 0;) {
      const u32 b = Lf.path_blk[i], sl = Lf.path_slot[i], pp = Lf.path_player[i], k = pool[b];
      const float v = fadd(pp == 0 ? val0 : val1, dshare);
      const u32 n0 = pool[fb_n(b, k) + sl];
      const float q0 = u2f(pool[fb_q(b, k) + sl]), d0 = u2f(pool[fb_d(b, k) + sl]);
      pool[fb_q(b, k) + sl] = f2u(fdiv(fadd(fmul(q0, (float)n0), v), (float)(n0 + 1u)));
      pool[fb_d(b, k) + sl] = f2u(fdiv(fadd(fmul(d0, (float)n0), vald), (float)(n0 + 1u)));
      if (n0 == 0) pool[fb_v(b, k) + sl] = f2u(fadd(lplayer == 0 ? val0 : val1, dshare));  // only the leaf can be new
      pool[fb_n(b, k) + sl] = n0 + 1u;
    }
    if (R.n == 0) {
      R.v = fadd(R.player == 0 ? val0 : val1, dshare);
      R.d = vald;
    }
    ++R.depth;
    ++R.n;
    Lf.path_len = 0;
  }
  __syncwarp();
}

// MCTS::update_root(gs, move) (mcts.cc:151-173) followed by gs.play_move(move) on the tree's root position
template <int GAME>
__device__ void forest_update_root(const ForestView& F, u32 t, u32 move, ForestSmem<GAME>& sm, u32 lane) {
  typedef FGame<GAME> G;
  ForestTree& R = F.trees[t];
  u32* pool = F.pool + (size_t)t * F.words_per_tree;
  u32 err = 0;
  u32 blk = R.blk, k = R.k;
  const u32 root_term = R.term;
  if (blk == 0) {
    // root_.children.empty(): add_children(gs.valid_moves()) — the shuffle draws happen; the block is only stored
    // when it can be descended into later (a terminal root keeps no children here, see find_leaf)
    Pcg32 rng = FOREST_RNG(F, t);
    {
      typename G::Pos pos;
      G::open(F, t, lane, sm, pos);
      k = G::legal(pos, sm, rng, lane, &err, F.serial_shuffle != 0);
    }
    if (lane == 0) FOREST_RNG(F, t) = rng;
    // the chosen child is a fresh node whatever its slot: only membership matters
    bool found = false;
    for (u32 j = lane; j < k; j += 32u) found |= sm.moves[j] == move;
    found = __any_sync(0xFFFFFFFFu, found);
    if (!found) err |= 8u;
    if (lane == 0) { R.n = 0; R.v = 0.0f; R.d = 0.0f; R.blk = 0; R.k = 0; R.player = 0; R.term = 0; R.nif = 0; R.expanded = 0; }
  } else {
    u32 slot = 0xFFFFFFFFu;
    for (u32 j = lane; j < k; j += 32u)
      if ((pool[fb_mv(blk, k) + j] & 0xFFFFu) == move) slot = j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u32 os = __shfl_xor_sync(0xFFFFFFFFu, slot, o);
      slot = os < slot ? os : slot;
    }
    if (slot == 0xFFFFFFFFu) {
      err |= 8u;
    } else if (lane == 0) {
      const u32 mvw = pool[fb_mv(blk, k) + slot];
      const u32 nb = pool[fb_fc(blk, k) + slot];
      R.n = pool[fb_n(blk, k) + slot];
      R.v = u2f(pool[fb_v(blk, k) + slot]);
      R.d = u2f(pool[fb_d(blk, k) + slot]);
      R.blk = nb;
      R.k = nb ? pool[nb] : 0u;
      R.player = (mvw >> 16) & 3u;
      R.term = (mvw >> 20) & 3u;
      R.expanded = (mvw >> 22) & 1u;
      R.nif = pool[fb_nif(blk, k) + slot];
    }
  }
  (void)root_term;
  __syncwarp();
  // The siblings of the chosen child are garbage now (the reference frees them: `root_ = std::move(tmp)` + ~Node).
  // Cheney copy of the kept subtree, breadth first, into the idle half of the slab; block indices are opaque to the
  // search and the children keep their order inside a block, so results do not depend on it.
  if (!err) {
    const u32 half_words = F.words_per_tree / 2u;
    const u32 to_base = R.half ? 1u : half_words, to_end = R.half ? half_words : F.words_per_tree;
    u32 to = to_base;
    const u32 old_root = R.blk;
    u32 new_root = 0;
    if (old_root) {
      const u32 rk = pool[old_root], rw = fb_words(rk);
      if (to + rw > to_end) {
        err |= 1u;
      } else {
        for (u32 i = lane; i < rw; i += 32u) pool[to + i] = pool[old_root + i];
        new_root = to;
        to += rw;
        __syncwarp();
        for (u32 scan = new_root; scan < to && !err;) {
          const u32 sk = pool[scan];
          for (u32 c0 = 0; c0 < sk && !err; c0 += 32u) {
            const u32 j = c0 + lane;
            const u32 fcj = j < sk ? pool[fb_fc(scan, sk) + j] : 0u;
            u32 live = __ballot_sync(0xFFFFFFFFu, fcj != 0u);
            while (live) {
              const int src_lane = __ffs((int)live) - 1;
              live &= live - 1u;
              const u32 src = __shfl_sync(0xFFFFFFFFu, fcj, src_lane);
              const u32 ck = pool[src], cw = fb_words(ck);
              if (to + cw > to_end) { err |= 1u; break; }
              for (u32 i = lane; i < cw; i += 32u) pool[to + i] = pool[src + i];
              if (lane == 0) pool[fb_fc(scan, sk) + c0 + (u32)src_lane] = to;
              to += cw;
            }
          }
          __syncwarp();
          scan += fb_words(sk);
        }
      }
    }
    if (lane == 0 && !err) { R.blk = new_root; R.bump = to; R.half ^= 1u; }
    __syncwarp();
  }
  // gs.play_move(move) with the persistent repetition history
  if (!err) err |= G::root_play(F, t, move, sm, lane);
  if (lane == 0) {
    if (F.gum) fg_reset(F.gum[t]);  // update_root ends with reset_gumbel_state() (mcts.cc:172)
    R.depth = 0;
    R.total_leaf_depth = 0;
    R.leaf.path_len = 0;
    R.in_flight = 0;
    if (err) R.error |= err;
  }
  __syncwarp();
}

// ---- kernels: one warp per tree, 4 warps per CTA
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_find_leaf(const AZ_GC_F ForestView F) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5)
    forest_find_leaf<GAME, false>(F, t, sm[wib], lane, true, F.trees[t].leaf, F.leaf_canon + (size_t)t * FGame<GAME>::canon(F));
}
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_process_result(const AZ_GC_F ForestView F, const float* ev_v, const float* ev_pi,
                                                               u32 root_noise_enabled) {
  const u32 lane = threadIdx.x & 31u;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5)
    forest_process_result<GAME, false, false>(F, t, ev_v, ev_pi, lane, root_noise_enabled != 0, F.trees[t].leaf);
}
// PlayManager's step after a move under tree reuse (play_manager.cc:546-553): the reused root gets the root
// temperature again and fresh noise — MCTS::apply_root_policy_temp() then add_root_noise() for trees with root_n > 0
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_root_noise(const AZ_GC_F ForestView F, u32 add_noise) {
  const u32 lane = threadIdx.x & 31u;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    ForestTree& R = F.trees[t];
    if (lane == 0 && R.n > 0 && R.blk != 0) {
      u32* pool = F.pool + (size_t)t * F.words_per_tree;
      fr_apply_root_policy_temp(F, t, pool, R.blk, R.k);
      if (add_noise && FSEAT(F, t).epsilon > 0.0f) {
        Pcg32 rng = FOREST_RNG(F, t);
        fr_add_root_noise(F, t, rng, pool, R.blk, R.k);
        FOREST_RNG(F, t) = rng;
      }
    }
  }
}
// n_sims x (find_leaf + dumb_eval + process_result) fused: the RANDOM-evaluator search (EvalType::RANDOM)
#ifndef B2AZ_FOREST_MINB
#define B2AZ_FOREST_MINB 8  /* 64 registers, 32 warps per SM: 100 -> 140 M sims/s (Brandubh, Gumbel) */
#endif
template <int GAME>
__global__ void __launch_bounds__(128, GAME == B2AZ_FOREST_SG ? 4 : B2AZ_FOREST_MINB) k_forest_simulate(const AZ_GC_F ForestView F, u32 n_sims, u32 root_noise_enabled) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5)
    for (u32 i = 0; i < n_sims; ++i) {
      forest_find_leaf<GAME, false>(F, t, sm[wib], lane, false, F.trees[t].leaf, nullptr);
      forest_process_result<GAME, true, false>(F, t, nullptr, nullptr, lane, root_noise_enabled != 0, F.trees[t].leaf);
    }
}
// ---- WU-UCT: MCTS::find_leaf_batched / process_result_batched / reset_batch (mcts.cc:752-851)
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_find_leaf_batched(const AZ_GC_F ForestView F) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    ForestTree& R = F.trees[t];
    const u32 li = R.in_flight;
    if (li >= F.max_in_flight) {
      if (lane == 0) R.error |= 16u;
      continue;
    }
    forest_find_leaf<GAME, true>(F, t, sm[wib], lane, true, F.inflight[(size_t)t * F.max_in_flight + li],
                                 F.leaf_canon + ((size_t)li * F.n_trees + t) * FGame<GAME>::canon(F));
    if (lane == 0) R.in_flight = li + 1u;
    __syncwarp();
  }
}
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_process_result_batched(const AZ_GC_F ForestView F, u32 leaf_index, const float* ev_v,
                                                                       const float* ev_pi, u32 root_noise_enabled) {
  const u32 lane = threadIdx.x & 31u;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    if (leaf_index >= F.trees[t].in_flight) continue;  // in_flight_.at(leaf_index) throws in the reference
    forest_process_result<GAME, false, true>(F, t, ev_v, ev_pi, lane, root_noise_enabled != 0,
                                             F.inflight[(size_t)t * F.max_in_flight + leaf_index]);
  }
}
// n_rounds x (width x find_leaf_batched, then width x process_result_batched with dumb_eval, then reset_batch) fused
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_simulate_batched(const AZ_GC_F ForestView F, u32 n_rounds, u32 width) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    ForestLeaf* fl = F.inflight + (size_t)t * F.max_in_flight;
    for (u32 r = 0; r < n_rounds; ++r) {
      for (u32 i = 0; i < width; ++i) forest_find_leaf<GAME, true>(F, t, sm[wib], lane, false, fl[i], nullptr);
      for (u32 i = 0; i < width; ++i) forest_process_result<GAME, true, true>(F, t, nullptr, nullptr, lane, false, fl[i]);
    }
    if (lane == 0) F.trees[t].in_flight = 0;
  }
}
__global__ void k_forest_reset_batch(const AZ_GC_F ForestView F) {
  for (u32 t = GLOBAL_TID; t < F.n_trees; t += GLOBAL_NT) F.trees[t].in_flight = 0;
}
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_update_root(const AZ_GC_F ForestView F, const u32* moves) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5)
    if (moves[t] != 0xFFFFFFFFu) forest_update_root<GAME>(F, t, moves[t], sm[wib], lane);
}
// Greedy self-play step entirely on the device: every tree that is not over plays its most visited root move
// (lowest move id on ties — argmax of MCTS::counts()) through update_root + play_move. Used by the throughput tool.
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_advance(const AZ_GC_F ForestView F) {
  __shared__ ForestSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    const ForestTree& R = F.trees[t];
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    const u32 b = R.blk, k = R.k;
    if (b == 0 || R.term != 0) continue;
    u32 best_n = 0, best_mv = 0xFFFFFFFFu;
    for (u32 j = lane; j < k; j += 32u) {
      const u32 nj = pool[fb_n(b, k) + j], mv = pool[fb_mv(b, k) + j] & 0xFFFFu;
      if (best_mv == 0xFFFFFFFFu || nj > best_n || (nj == best_n && mv < best_mv)) { best_n = nj; best_mv = mv; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u32 on = __shfl_xor_sync(0xFFFFFFFFu, best_n, o), om = __shfl_xor_sync(0xFFFFFFFFu, best_mv, o);
      if (om != 0xFFFFFFFFu && (best_mv == 0xFFFFFFFFu || on > best_n || (on == best_n && om < best_mv))) { best_n = on; best_mv = om; }
    }
    if (best_mv != 0xFFFFFFFFu) forest_update_root<GAME>(F, t, best_mv, sm[wib], lane);
  }
}
// MCTS::probs(temp) (mcts.cc:575-618) into out[A] and, with `pick`, MCTS::pick_move(probs) (mcts.cc:717-735: one
// uniform draw, first move whose running sum exceeds it). The reference works on dense num_moves-long vectors and
// sums them in MOVE order: the counts (or priors) are scattered into the dense row by all lanes, the sums, pow() and
// the cumulative pick run on lane 0 over the A entries (once per move: not a hot path).
template <int GAME>
__device__ void forest_probs(const ForestView& F, u32 t, float temp, float* out, u32* picked_out, u32 pick, u32 pruned, u32 lane) {
  const u32 A = FGame<GAME>::actions(F);
  {
    ForestTree& R = F.trees[t];
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    const u32 b = R.blk, k = b ? R.k : 0u;
    for (u32 m = lane; m < A; m += 32u) out[m] = 0.0f;
    __syncwarp();
    u32 total = 0;
    for (u32 j = lane; j < k; j += 32u) total += pool[fb_n(b, k) + j];
    total = warp_sum(total);  // counts.cast<float>().sum() == 0 <=> no child has a visit
    // MCTS::probs_pruned (mcts.cc:620-674, KataGo's policy-target pruning by PUCT inversion) when asked for and the
    // root has more than one visit: reduced visit counts instead of the raw ones; falls back to probs() when
    // nothing survives
    bool use_pruned = false;
    if (pruned && R.n > 1 && k > 0) {
      const float es = fmul(F.cpuct, fsqrt((float)R.n));
      float best_sel = -1e30f;
      for (u32 j = lane; j < k; j += 32u) {
        const u32 nj = pool[fb_n(b, k) + j];
        if (nj == 0) continue;
        const float sel = fadd(u2f(pool[fb_q(b, k) + j]), fdiv(fmul(es, u2f(pool[fb_pol(b, k) + j])), (float)(nj + 1u)));
        if (sel > best_sel) best_sel = sel;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // a maximum: order independent
        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best_sel, o);
        if (ob > best_sel) best_sel = ob;
      }
      for (u32 j = lane; j < k; j += 32u) {
        const u32 nj = pool[fb_n(b, k) + j];
        if (nj == 0) continue;
        const float qj = u2f(pool[fb_q(b, k) + j]);
        const float gap = fsub(best_sel, qj);
        const float desired = gap <= 0.0f ? (float)nj : fsub(fdiv(fmul(es, u2f(pool[fb_pol(b, k) + j])), gap), 1.0f);
        out[pool[fb_mv(b, k) + j] & 0xFFFFu] = std_min((float)nj, std_max(0.0f, desired));
      }
      __syncwarp();
      float ptotal = 0.0f;  // pruned.sum() over the dense vector in move order (zero entries do not change a sum)
      for (u32 c0 = 0; c0 < A; c0 += 32u) {
        const float val = c0 + lane < A ? out[c0 + lane] : 0.0f;
        ptotal = seq_sum_masked(ptotal, val, val != 0.0f);
      }
      use_pruned = ptotal != 0.0f;
      if (!use_pruned) {
        for (u32 m = lane; m < A; m += 32u) out[m] = 0.0f;
        __syncwarp();
      }
    }
    if (!use_pruned)
      for (u32 j = lane; j < k; j += 32u) {
        const u32 mv = pool[fb_mv(b, k) + j] & 0xFFFFu;
        out[mv] = total == 0 ? u2f(pool[fb_pol(b, k) + j]) : (float)pool[fb_n(b, k) + j];
      }
    __syncwarp();
    // The reference works on the dense num_moves-long vector in MOVE order. Adding a zero entry never changes a float
    // sum and pow(0, e > 0) == 0, so only the non-zero entries (at most k of the A) carry the arithmetic: every chunk of
    // 32 moves is loaded one per lane and its non-zero values are folded in ascending move order with ballots + shuffles
    // (every lane accumulates the same sequence); the element-wise steps run one move per lane.
    auto dense_sum = [&]() {
      float acc = 0.0f;
      for (u32 c0 = 0; c0 < A; c0 += 32u) {
        const float val = c0 + lane < A ? out[c0 + lane] : 0.0f;
        acc = seq_sum_masked(acc, val, val != 0.0f);
      }
      return acc;
    };
    auto dense_pow = [&](float e, float divide_by, bool divide_first) {  // out = pow(divide_first ? out / d : out, e), then Σ
      for (u32 m = lane; m < A; m += 32u) {
        float val = out[m];
        if (divide_first) val = fdiv(val, divide_by);
        if (val != 0.0f || !(e > 0.0f)) val = az_powf(val, e);
        out[m] = val;
      }
      __syncwarp();
      return dense_sum();
    };
    auto dense_div = [&](float d) {
      for (u32 m = lane; m < A; m += 32u) out[m] = fdiv(out[m], d);
      __syncwarp();
    };
    auto dense_argmax_share = [&](bool double_share) {  // uniform over the largest entries (order independent)
      float best = -INFINITY;
      for (u32 m = lane; m < A; m += 32u) best = out[m] > best ? out[m] : best;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        best = ob > best ? ob : best;
      }
      u32 ties = 0;
      for (u32 m = lane; m < A; m += 32u) ties += out[m] == best ? 1u : 0u;
      ties = warp_sum(ties);
      const float share = double_share ? (float)(1.0 / (double)ties) : fdiv(1.0f, (float)ties);
      for (u32 m = lane; m < A; m += 32u) out[m] = out[m] == best ? share : 0.0f;
      __syncwarp();
    };
    if (use_pruned) {
      const float tot = dense_sum();
      if (temp == 0.0f) {
        dense_argmax_share(false);
      } else {
        dense_div(tot);
        if (temp != 1.0f) {
          const float sum2 = dense_pow(fdiv(1.0f, temp), 1.0f, false);
          dense_div(sum2);
        }
      }
    } else if (total == 0) {  // the prior policy (raw-policy mode), tempered
      float sum;
      if (temp != 0.0f) sum = dense_pow(fdiv(1.0f, temp), 1.0f, false);
      else sum = dense_sum();
      dense_div(sum);
    } else if (temp == 0.0f) {  // uniform over the most visited moves
      dense_argmax_share(true);
    } else {
      const float sum = dense_sum();
      const float sum2 = dense_pow(fdiv(1.0f, temp), sum, true);  // `1 / temp`: int / float
      dense_div(sum2);
    }
    if (pick) {  // MCTS::pick_move: one uniform draw, first move whose running sum exceeds it, else the last positive one
      float choice = 0.0f;
      if (lane == 0) {
        Pcg32 rng = FOREST_RNG(F, t);
        choice = rng_uniform01(rng);
        FOREST_RNG(F, t) = rng;
      }
      choice = __shfl_sync(0xFFFFFFFFu, choice, 0);
      u32 mvp = 0xFFFFFFFFu, last_pos = 0xFFFFFFFFu;
      float sum = 0.0f;
      for (u32 c0 = 0; c0 < A && mvp == 0xFFFFFFFFu; c0 += 32u) {
        const float val = c0 + lane < A ? out[c0 + lane] : 0.0f;
        u32 nz = __ballot_sync(0xFFFFFFFFu, val != 0.0f);
        const u32 pos = __ballot_sync(0xFFFFFFFFu, val > 0.0f);
        if (pos) last_pos = c0 + 31u - (u32)__clz((int)pos);
        while (nz) {
          const int src = __ffs((int)nz) - 1;
          nz &= nz - 1u;
          sum = fadd(sum, __shfl_sync(0xFFFFFFFFu, val, src));
          if (sum > choice) { mvp = c0 + (u32)src; break; }
        }
      }
      if (mvp == 0xFFFFFFFFu) mvp = last_pos;  // (the scan ran to the end: last_pos covers the whole row)
      if (lane == 0) *picked_out = mvp;
    }
    __syncwarp();
  }
}
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_probs(const AZ_GC_F ForestView F, float temp, float* probs, u32* picked, u32 pick, u32 pruned) {
  const u32 lane = threadIdx.x & 31u;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5)
    forest_probs<GAME>(F, t, temp, probs + (size_t)t * FGame<GAME>::actions(F), picked + t, pick, pruned, lane);
}
// MCTS::counts / root_q_values (mcts.cc:557-573) + a few scalars per tree
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_counts(const AZ_GC_F ForestView F, u32* counts, float* q, u32* info) {
  const u32 A = FGame<GAME>::actions(F);
  const u32 lane = threadIdx.x & 31u;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    const ForestTree& R = F.trees[t];
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    for (u32 m = lane; m < A; m += 32u) {
      if (counts) counts[(size_t)t * A + m] = 0u;
      if (q) q[(size_t)t * A + m] = 0.0f;
    }
    __syncwarp();
    const u32 b = R.blk, k = R.k;
    if (b)
      for (u32 j = lane; j < k; j += 32u) {
        const u32 mv = pool[fb_mv(b, k) + j] & 0xFFFFu;
        if (counts) counts[(size_t)t * A + mv] = pool[fb_n(b, k) + j];
        if (q) q[(size_t)t * A + mv] = u2f(pool[fb_q(b, k) + j]);
      }
    if (info && lane == 0) {
      u32* o = info + (size_t)t * 16;
      o[0] = R.depth; o[1] = R.n; o[2] = R.k; o[3] = R.term; o[4] = R.player;
      FGame<GAME>::info(F, t, &o[5], &o[6], &o[11]);
      o[7] = R.error; o[8] = R.bump; o[9] = f2u(R.v); o[10] = R.total_leaf_depth;
      {  // MCTS::root_value (mcts.h:78-100): win / loss / draw from the best-Q visited child, else the root's own value
        float qv = 0.0f, dv = 0.0f;
        bool found = false;
        for (u32 j = 0; b && j < k; ++j) {
          const float qj = u2f(pool[fb_q(b, k) + j]);
          if (pool[fb_n(b, k) + j] > 0 && qj > qv) { qv = qj; dv = u2f(pool[fb_d(b, k) + j]); found = true; }
        }
        if (!found && R.n > 0) { qv = R.v; dv = R.d; }
        const float w = fsub(qv, fdiv(dv, 2.0f));
        o[12] = f2u(w); o[13] = f2u((float)dsub(dsub(1.0, (double)w), (double)dv)); o[14] = f2u(dv); o[15] = R.in_flight;
      }
    }
  }
}
// MCTS::set_gumbel_num_sims(n) (mcts.cc:175-178) for every tree
__global__ void k_forest_gumbel_arm(const AZ_GC_F ForestView F, u32 n) {
  for (u32 t = GLOBAL_TID; t < F.n_trees; t += GLOBAL_NT) {
    F.gum[t].num_sims_target = n;
    fg_reset(F.gum[t]);
  }
}
// gumbel_final_action + gumbel_improved_policy per tree
template <int GAME>
__global__ void __launch_bounds__(128) k_forest_gumbel_result(const AZ_GC_F ForestView F, u32* action, float* policy) {
  const u32 A = FGame<GAME>::actions(F);
  const u32 lane = threadIdx.x & 31u;
  for (u32 t = GLOBAL_TID >> 5; t < F.n_trees; t += GLOBAL_NT >> 5) {
    const ForestTree& R = F.trees[t];
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    if (policy) {
      for (u32 m = lane; m < A; m += 32u) policy[(size_t)t * A + m] = 0.0f;
      __syncwarp();
      if (lane == 0) fg_improved_policy(F, t, R, pool, policy + (size_t)t * A);
    }
    if (action && lane == 0) action[t] = fg_final_action(F, t, R, F.gum[t], pool);
    __syncwarp();
  }
}
// MCTS::principal_variation(depth) (mcts.cc:676-715): the most-visited line from the root (first maximum in child
// order; the root step is the Gumbel final action when Gumbel is active). out [n_trees][depth], len [n_trees].
__global__ void k_forest_pv(const AZ_GC_F ForestView F, u32 depth, u32* out, u32* len) {
  for (u32 t = GLOBAL_TID; t < F.n_trees; t += GLOBAL_NT) {
    const ForestTree& R = F.trees[t];
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    u32 b = R.blk, k = b ? R.k : 0u, n = 0;
    for (u32 i = 0; i < depth && b != 0 && k != 0; ++i) {
      u32 best = 0xFFFFFFFFu;
      if (i == 0 && FSEAT(F, t).gumbel_enabled) {
        const u32 mv = fg_final_action(F, t, R, F.gum[t], pool);
        for (u32 j = 0; j < k && best == 0xFFFFFFFFu; ++j)
          if ((pool[fb_mv(b, k) + j] & 0xFFFFu) == mv) best = j;
      }
      if (best == 0xFFFFFFFFu) {
        u32 best_n = 0;
        for (u32 j = 0; j < k; ++j) {
          const u32 nj = pool[fb_n(b, k) + j];
          if (nj > best_n) { best_n = nj; best = j; }
        }
      }
      if (best == 0xFFFFFFFFu || pool[fb_n(b, k) + best] == 0) break;
      out[(size_t)t * depth + n++] = pool[fb_mv(b, k) + best] & 0xFFFFu;
      const u32 nb = pool[fb_fc(b, k) + best];
      b = nb; k = nb ? pool[nb] : 0u;
    }
    len[t] = n;
  }
}
// the moves along the path of a pending leaf (MCTS::path_ / one InFlightLeaf): what the caller replays on its own copy
// of the root position to obtain the leaf GameState find_leaf returns. slot < 0: the plain find_leaf's leaf.
__global__ void k_forest_leaf_path(const AZ_GC_F ForestView F, int slot, u32* out, u32* len) {
  for (u32 t = GLOBAL_TID; t < F.n_trees; t += GLOBAL_NT) {
    const ForestLeaf& L = slot < 0 ? F.trees[t].leaf : F.inflight[(size_t)t * F.max_in_flight + (u32)slot];
    const u32* pool = F.pool + (size_t)t * F.words_per_tree;
    for (u32 i = 0; i < L.path_len; ++i) {
      const u32 b = L.path_blk[i], k = pool[b];
      out[(size_t)t * kFPath + i] = pool[fb_mv(b, k) + L.path_slot[i]] & 0xFFFFu;
    }
    len[t] = L.path_len;
  }
}
// MCTS::apply_root_policy_temp (mcts.cc:448-460) and / or MCTS::add_root_noise (mcts.cc:403-446), separately callable
__global__ void k_forest_root_ops(const AZ_GC_F ForestView F, u32 apply_temp, u32 add_noise) {
  for (u32 t = GLOBAL_TID; t < F.n_trees; t += GLOBAL_NT) {
    ForestTree& R = F.trees[t];
    if (R.blk == 0 || R.k == 0) continue;
    u32* pool = F.pool + (size_t)t * F.words_per_tree;
    if (apply_temp) fr_apply_root_policy_temp(F, t, pool, R.blk, R.k);
    if (add_noise && F.noise) {
      Pcg32 rng = FOREST_RNG(F, t);
      fr_add_root_noise(F, t, rng, pool, R.blk, R.k);
      FOREST_RNG(F, t) = rng;
    }
  }
}
template <int GAME>
__global__ void k_forest_init(const AZ_GC_F ForestView F, unsigned long long seed) {
  for (u32 t = GLOBAL_TID; t < F.n_trees; t += GLOBAL_NT) {
    ForestTree& R = F.trees[t];
    memset(&R, 0, sizeof(ForestTree));
    FGame<GAME>::init(F, t);
    R.bump = 1;
    pcg32_seed(R.rng, seed + t);  // tree t == a reference MCTS driven after MCTS::seed_thread_rng(seed + t)
  }
}
#endif  // !B2AZ_HOST_EMU

}  // namespace b2az

struct b2az_forest {
  b2az::ForestView view;
  int device = 0;
  uint32_t actions = 0, canon = 0;
  uint32_t* moves_dev = nullptr;
  float *ev_v = nullptr, *ev_pi = nullptr;
};

#ifndef B2AZ_HOST_EMU
#define FOREST_DISPATCH(F, CALL)                                                          \
  switch ((F)->view.game) {                                                               \
    case B2AZ_TAFL_BRANDUBH: { constexpr int G_ = B2AZ_TAFL_BRANDUBH; CALL; } break;      \
    case B2AZ_TAFL_OPENTAFL: { constexpr int G_ = B2AZ_TAFL_OPENTAFL; CALL; } break;      \
    case B2AZ_TAFL_TAWLBWRDD: { constexpr int G_ = B2AZ_TAFL_TAWLBWRDD; CALL; } break;    \
    case B2AZ_FOREST_C4: { constexpr int G_ = B2AZ_FOREST_C4; CALL; } break;              \
    default: { constexpr int G_ = B2AZ_FOREST_SG; CALL; } break;                          \
  }
static inline unsigned forest_ctas(const b2az_forest* f) { return std::max(1u, std::min((f->view.n_trees + 3u) / 4u, 148u * 8u)); }
#endif

extern "C" {

int b2az_forest_create(const b2az_forest_params* p, int device, b2az_forest** out) {
  using namespace b2az;
  if (!p || !out) return fail(B2AZ_EINVAL, "null argument");
  const bool is_sg = (p->game >= 10 && p->game <= 13) || (p->game >= 20 && p->game <= 24);  // 24: Unified, variant mix
  const bool is_c4 = p->game == B2AZ_FOREST_C4;
  if (p->game > B2AZ_TAFL_TAWLBWRDD && !is_sg && !is_c4) return fail(B2AZ_EINVAL, "b2az_forest: unknown game");
  if (p->n_trees == 0 || p->max_turns == 0 || p->max_turns > 65535u) return fail(B2AZ_EINVAL, "b2az_forest: bad n_trees / max_turns");
  if (!(p->root_policy_temp > 0.0f)) return fail(B2AZ_EINVAL, "b2az_forest: root_policy_temp must be positive (1 = off)");
  if (p->epsilon < 0.0f || p->epsilon > 1.0f) return fail(B2AZ_EINVAL, "b2az_forest: epsilon must be in [0, 1]");
  if (p->gumbel_enabled && (p->gumbel_m == 0 || p->gumbel_m > (uint32_t)kFMaxM))
    return fail(B2AZ_EINVAL, "b2az_forest: gumbel_m must be in [1, 64]");
#ifdef B2AZ_HOST_EMU
  (void)device;
  return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  b2az_forest* f = new b2az_forest();
  f->device = device;
  ForestView& V = f->view;
  memset(&V, 0, sizeof(V));
  V.n_trees = p->n_trees; V.max_turns = p->max_turns; V.game = p->game;
  V.cpuct = p->cpuct; V.fpu_reduction = p->fpu_reduction;
  if (is_sg) {
    const SGSpace sp = sg_space(p->game == 24u ? B2AZ_SG_BATTLE : (int)(p->game % 10u), p->game >= 20u);
    for (int i = 0; i < 4; ++i) V.sg_probs[i] = 0.25f;
    f->actions = (uint32_t)sp.num_moves();
    f->canon = (uint32_t)(sp.planes(p->game >= 20u) * sp.udim * sp.udim);
  } else if (is_c4) {
    f->actions = 7u;
    f->canon = (uint32_t)C4_CANON;
  } else {
    const uint32_t S = p->game == B2AZ_TAFL_BRANDUBH ? 7u : 11u, planes = p->game == B2AZ_TAFL_OPENTAFL ? 8u : 7u;
    f->actions = 2u * S * S * S;
    f->canon = planes * S * S;
  }
  V.actions = f->actions; V.canon = f->canon; V.relative_values = p->relative_values ? 1u : 0u;
  V.words_per_tree = p->words_per_tree ? p->words_per_tree : (1u << 20);
  auto bail = [&](int rc) { b2az_forest_destroy(f); return rc; };
  if (int rc = dev_alloc(&V.trees, (size_t)V.n_trees)) return bail(rc);
  if (int rc = dev_alloc_raw(&V.pool, (size_t)V.n_trees * V.words_per_tree)) return bail(rc);
  if (is_sg) {  // position keys since the last deploy (a turn is several actions: the bound is generous, overflow is reported)
    V.sg_hist_cap = 2048u;
    if (int rc = dev_alloc(&V.sg_state, (size_t)V.n_trees)) return bail(rc);
    if (int rc = dev_alloc(&V.sg_hist, (size_t)V.n_trees * V.sg_hist_cap)) return bail(rc);
    if (int rc = dev_alloc(&V.sg_pkeys, (size_t)V.n_trees * (kFPath + 2))) return bail(rc);
  } else {
    if (int rc = dev_alloc(&V.hist, (size_t)V.n_trees * (V.max_turns + 2u))) return bail(rc);
    if (int rc = dev_alloc(&V.pkeys, (size_t)V.n_trees * (kFPath + 2))) return bail(rc);
  }
  V.max_in_flight = p->max_in_flight;
  if (V.max_in_flight > 64u) return bail(fail(B2AZ_EINVAL, "b2az_forest: max_in_flight must be <= 64"));
  if (int rc = dev_alloc(&V.leaf_canon, (size_t)V.n_trees * f->canon * std::max(1u, V.max_in_flight))) return bail(rc);
  if (V.max_in_flight)
    if (int rc = dev_alloc(&V.inflight, (size_t)V.n_trees * V.max_in_flight)) return bail(rc);
  if (int rc = dev_alloc(&f->moves_dev, (size_t)V.n_trees)) return bail(rc);
  V.shaped_dirichlet = p->shaped_dirichlet ? 1u : 0u;
  V.serial_shuffle = p->debug_serial_shuffle ? 1u : 0u;
  if (p->epsilon > 0.0f)
    if (int rc = dev_alloc(&V.noise, (size_t)V.n_trees * kFMaxK)) return bail(rc);
  if (p->gumbel_enabled) {
    if (int rc = dev_alloc(&V.gum, (size_t)V.n_trees)) return bail(rc);
    if (int rc = dev_alloc(&V.gum_g, (size_t)V.n_trees * 2 * kFMaxK)) return bail(rc);
  }
  {
    SeatSearch* sets = nullptr;
    if (int rc = dev_alloc(&sets, (size_t)V.n_trees)) return bail(rc);
    V.seat = sets;
    const SeatSearch s0{p->epsilon, p->root_policy_temp, p->gumbel_c_visit, p->gumbel_c_scale, p->gumbel_m, (u8)(p->root_fpu_zero ? 1 : 0),
                        (u8)(p->gumbel_enabled ? 1 : 0), (u8)(p->gumbel_enabled && p->gumbel_full ? 1 : 0), 0};
    std::vector<SeatSearch> h((size_t)V.n_trees, s0);  // every tree: the forest's own settings
    CUDA_TRY(cudaMemcpy(sets, h.data(), h.size() * sizeof(SeatSearch), cudaMemcpyHostToDevice));
  }
  FOREST_DISPATCH(f, (k_forest_init<G_><<<148, 128>>>(V, p->seed)));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  *out = f;
  return 0;
#endif
}

int b2az_forest_destroy(b2az_forest* f) {
  using namespace b2az;
  if (!f) return 0;
  dev_free(f->view.trees); dev_free(f->view.pool); dev_free(f->view.hist); dev_free(f->view.pkeys);
  dev_free(f->view.leaf_canon); dev_free(f->moves_dev); dev_free(f->ev_v); dev_free(f->ev_pi);
  dev_free(f->view.gum); dev_free(f->view.gum_g); dev_free(f->view.noise); dev_free(f->view.inflight);
  dev_free(const_cast<SeatSearch*>(f->view.seat));
  dev_free(f->view.sg_state); dev_free(f->view.sg_hist); dev_free(f->view.sg_pkeys);
  delete f;
  return 0;
}

#ifdef B2AZ_HOST_EMU
#define FOREST_NO_CUDA(...) { return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback"); }
int b2az_forest_find_leaf(b2az_forest*, void*, const float**) FOREST_NO_CUDA()
int b2az_forest_leaf_canon_host(b2az_forest*, void*, float*) FOREST_NO_CUDA()
int b2az_forest_process_result(b2az_forest*, void*, const float*, const float*, int) FOREST_NO_CUDA()
int b2az_forest_process_result_host(b2az_forest*, void*, const float*, const float*, int) FOREST_NO_CUDA()
int b2az_forest_simulate(b2az_forest*, void*, uint32_t, int) FOREST_NO_CUDA()
int b2az_forest_root_noise(b2az_forest*, void*, int) FOREST_NO_CUDA()
int b2az_forest_find_leaf_batched(b2az_forest*, void*, const float**) FOREST_NO_CUDA()
int b2az_forest_process_result_batched(b2az_forest*, void*, uint32_t, const float*, const float*, int, int) FOREST_NO_CUDA()
int b2az_forest_simulate_batched(b2az_forest*, void*, uint32_t, uint32_t) FOREST_NO_CUDA()
int b2az_forest_reset_batch(b2az_forest*, void*) FOREST_NO_CUDA()
int b2az_forest_probs(b2az_forest*, void*, float, int, int, float*, uint32_t*) FOREST_NO_CUDA()
int b2az_forest_advance(b2az_forest*, void*) FOREST_NO_CUDA()
int b2az_forest_set_gumbel_num_sims(b2az_forest*, void*, uint32_t) FOREST_NO_CUDA()
int b2az_forest_gumbel_result(b2az_forest*, void*, uint32_t*, float*) FOREST_NO_CUDA()
int b2az_forest_update_root(b2az_forest*, void*, const uint32_t*) FOREST_NO_CUDA()
int b2az_forest_counts(b2az_forest*, void*, uint32_t*, float*, uint32_t*) FOREST_NO_CUDA()
int b2az_forest_principal_variation(b2az_forest*, void*, uint32_t, uint32_t*, uint32_t*) FOREST_NO_CUDA()
int b2az_forest_leaf_path(b2az_forest*, void*, int, uint32_t*, uint32_t*) FOREST_NO_CUDA()
int b2az_forest_root_ops(b2az_forest*, void*, int, int) FOREST_NO_CUDA()
int b2az_forest_set_root(b2az_forest*, uint32_t, const void*, uint32_t, const void*, uint32_t) FOREST_NO_CUDA()
int b2az_forest_get_root(b2az_forest*, uint32_t, void*, uint32_t, void*, uint32_t, uint32_t*) FOREST_NO_CUDA()
#else
int b2az_forest_find_leaf(b2az_forest* f, void* stream, const float** canon_dev) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_find_leaf<G_><<<forest_ctas(f), 128, 0, s>>>(f->view)));
  CUDA_TRY(cudaGetLastError());
  if (canon_dev) *canon_dev = f->view.leaf_canon;
  return 0;
}
int b2az_forest_leaf_canon_host(b2az_forest* f, void* stream, float* canon_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f || !canon_host) return fail(B2AZ_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // every in-flight slot's rows (one slot for the plain find_leaf)
  const size_t slots = std::max(1u, f->view.max_in_flight);
  if (int rc = copy_d2h(canon_host, f->view.leaf_canon, slots * f->view.n_trees * f->canon * 4, s)) return rc;
  return stream_sync(s);
}
int b2az_forest_process_result(b2az_forest* f, void* stream, const float* v_dev, const float* pi_dev,
                               int root_noise_enabled) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f || !v_dev || !pi_dev) return fail(B2AZ_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_process_result<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, v_dev, pi_dev, root_noise_enabled ? 1u : 0u)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_process_result_host(b2az_forest* f, void* stream, const float* v_host, const float* pi_host,
                                    int root_noise_enabled) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f || !v_host || !pi_host) return fail(B2AZ_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees;
  if (!f->ev_v) {
    if (int rc = dev_alloc(&f->ev_v, n * 3)) return rc;
    if (int rc = dev_alloc(&f->ev_pi, n * f->actions)) return rc;
  }
  if (int rc = copy_h2d(f->ev_v, v_host, n * 3 * 4, s)) return rc;
  if (int rc = copy_h2d(f->ev_pi, pi_host, n * f->actions * 4, s)) return rc;
  if (int rc = b2az_forest_process_result(f, stream, f->ev_v, f->ev_pi, root_noise_enabled)) return rc;
  return stream_sync(s);
}
int b2az_forest_find_leaf_batched(b2az_forest* f, void* stream, const float** canon_dev) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (!f->view.max_in_flight) return fail(B2AZ_ESTATE, "b2az_forest: created with max_in_flight == 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_find_leaf_batched<G_><<<forest_ctas(f), 128, 0, s>>>(f->view)));
  CUDA_TRY(cudaGetLastError());
  if (canon_dev) *canon_dev = f->view.leaf_canon;
  return 0;
}
int b2az_forest_process_result_batched(b2az_forest* f, void* stream, uint32_t leaf_index, const float* v, const float* pi,
                                       int root_noise_enabled, int host_pointers) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f || !v || !pi) return fail(B2AZ_EINVAL, "null argument");
  if (leaf_index >= f->view.max_in_flight) return fail(B2AZ_EINVAL, "b2az_forest: leaf_index out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees;
  if (host_pointers) {
    if (!f->ev_v) {
      if (int rc = dev_alloc(&f->ev_v, n * 3)) return rc;
      if (int rc = dev_alloc(&f->ev_pi, n * f->actions)) return rc;
    }
    if (int rc = copy_h2d(f->ev_v, v, n * 3 * 4, s)) return rc;
    if (int rc = copy_h2d(f->ev_pi, pi, n * f->actions * 4, s)) return rc;
    v = f->ev_v; pi = f->ev_pi;
  }
  FOREST_DISPATCH(f, (k_forest_process_result_batched<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, leaf_index, v, pi,
                                                                                        root_noise_enabled ? 1u : 0u)));
  CUDA_TRY(cudaGetLastError());
  return host_pointers ? stream_sync(s) : 0;
}
int b2az_forest_simulate_batched(b2az_forest* f, void* stream, uint32_t n_rounds, uint32_t width) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (width == 0 || width > f->view.max_in_flight) return fail(B2AZ_EINVAL, "b2az_forest: width must be in [1, max_in_flight]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_simulate_batched<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, n_rounds, width)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_reset_batch(b2az_forest* f, void* stream) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  k_forest_reset_batch<<<148, 128, 0, static_cast<cudaStream_t>(stream)>>>(f->view);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_probs(b2az_forest* f, void* stream, float temp, int pruned, int pick_move, float* probs_host,
                      uint32_t* moves_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (pick_move && !moves_host) return fail(B2AZ_EINVAL, "b2az_forest_probs: pick_move needs moves_host");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees, A = f->actions;
  float* dp = nullptr;
  u32* dm = nullptr;
  int rc = dev_alloc(&dp, n * A);
  if (!rc) rc = dev_alloc(&dm, n);
  if (!rc) {
    FOREST_DISPATCH(f, (k_forest_probs<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, temp, dp, dm, pick_move ? 1u : 0u, pruned ? 1u : 0u)));
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_forest_probs launch failed");
  }
  if (!rc && probs_host) rc = copy_d2h(probs_host, dp, n * A * 4, s);
  if (!rc && pick_move) rc = copy_d2h(moves_host, dm, n * 4, s);
  if (!rc) rc = stream_sync(s);
  dev_free(dp); dev_free(dm);
  return rc;
}
int b2az_forest_root_noise(b2az_forest* f, void* stream, int add_noise) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_root_noise<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, add_noise ? 1u : 0u)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_simulate(b2az_forest* f, void* stream, uint32_t n_sims, int root_noise_enabled) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (n_sims == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_simulate<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, n_sims, root_noise_enabled ? 1u : 0u)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_set_gumbel_num_sims(b2az_forest* f, void* stream, uint32_t n) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (!f->view.gum) return fail(B2AZ_ESTATE, "b2az_forest: created without gumbel_enabled");
  k_forest_gumbel_arm<<<148, 128, 0, static_cast<cudaStream_t>(stream)>>>(f->view, n);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_gumbel_result(b2az_forest* f, void* stream, uint32_t* action_host, float* policy_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (!f->view.gum) return fail(B2AZ_ESTATE, "b2az_forest: created without gumbel_enabled");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees, A = f->actions;
  u32* da = nullptr;
  float* dp = nullptr;
  int rc = 0;
  if (action_host) rc = rc ? rc : dev_alloc(&da, n);
  if (policy_host) rc = rc ? rc : dev_alloc(&dp, n * A);
  if (!rc) {
    FOREST_DISPATCH(f, (k_forest_gumbel_result<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, da, dp)));
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_forest_gumbel_result launch failed");
  }
  if (!rc && action_host) rc = copy_d2h(action_host, da, n * 4, s);
  if (!rc && policy_host) rc = copy_d2h(policy_host, dp, n * A * 4, s);
  if (!rc) rc = stream_sync(s);
  dev_free(da); dev_free(dp);
  return rc;
}
int b2az_forest_advance(b2az_forest* f, void* stream) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOREST_DISPATCH(f, (k_forest_advance<G_><<<forest_ctas(f), 128, 0, s>>>(f->view)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_update_root(b2az_forest* f, void* stream, const uint32_t* moves_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f || !moves_host) return fail(B2AZ_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = copy_h2d(f->moves_dev, moves_host, (size_t)f->view.n_trees * 4, s)) return rc;
  FOREST_DISPATCH(f, (k_forest_update_root<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, f->moves_dev)));
  CUDA_TRY(cudaGetLastError());
  return stream_sync(s);
}
int b2az_forest_principal_variation(b2az_forest* f, void* stream, uint32_t depth, uint32_t* moves_host, uint32_t* len_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));
  using namespace b2az;
  if (!f || !moves_host || !len_host || depth == 0) return fail(B2AZ_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees;
  u32 *dm = nullptr, *dl = nullptr;
  int rc = dev_alloc(&dm, n * depth);
  if (!rc) rc = dev_alloc(&dl, n);
  if (!rc) {
    k_forest_pv<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(f->view, depth, dm, dl);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_forest_pv launch failed");
  }
  if (!rc) rc = copy_d2h(moves_host, dm, n * depth * 4, s);
  if (!rc) rc = copy_d2h(len_host, dl, n * 4, s);
  if (!rc) rc = stream_sync(s);
  dev_free(dm); dev_free(dl);
  return rc;
}
int b2az_forest_leaf_path(b2az_forest* f, void* stream, int slot, uint32_t* moves_host, uint32_t* len_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));
  using namespace b2az;
  if (!f || !moves_host || !len_host) return fail(B2AZ_EINVAL, "null argument");
  if (slot >= (int)f->view.max_in_flight) return fail(B2AZ_EINVAL, "b2az_forest_leaf_path: slot out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees;
  u32 *dm = nullptr, *dl = nullptr;
  int rc = dev_alloc(&dm, n * kFPath);
  if (!rc) rc = dev_alloc(&dl, n);
  if (!rc) {
    k_forest_leaf_path<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(f->view, slot, dm, dl);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_forest_leaf_path launch failed");
  }
  if (!rc) rc = copy_d2h(moves_host, dm, n * kFPath * 4, s);
  if (!rc) rc = copy_d2h(len_host, dl, n * 4, s);
  if (!rc) rc = stream_sync(s);
  dev_free(dm); dev_free(dl);
  return rc;
}
int b2az_forest_root_ops(b2az_forest* f, void* stream, int apply_temp, int add_noise) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  if (add_noise && !f->view.noise) return fail(B2AZ_ESTATE, "b2az_forest: created with epsilon == 0 (no noise buffers)");
  k_forest_root_ops<<<(f->view.n_trees + 127u) / 128u, 128, 0, static_cast<cudaStream_t>(stream)>>>(f->view, apply_temp ? 1u : 0u, add_noise ? 1u : 0u);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int b2az_forest_set_root(b2az_forest* f, uint32_t tree, const void* state, uint32_t state_bytes, const void* hist,
                         uint32_t hist_count) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));
  using namespace b2az;
  if (!f || !state) return fail(B2AZ_EINVAL, "null argument");
  if (tree >= f->view.n_trees) return fail(B2AZ_EINVAL, "b2az_forest_set_root: tree out of range");
  const bool sg = f->view.sg_state != nullptr;
  const size_t want = sg ? sizeof(SGState) : sizeof(TaflState), key = sg ? sizeof(u64) : sizeof(TaflKey);
  const size_t cap = sg ? f->view.sg_hist_cap : (size_t)f->view.max_turns + 2u;
  if (state_bytes != want) return fail(B2AZ_EINVAL, "b2az_forest_set_root: state record of the wrong size");
  if (hist_count > cap || (hist_count && !hist)) return fail(B2AZ_EINVAL, "b2az_forest_set_root: history too long");
  ForestTree h;
  CUDA_TRY(cudaMemcpy(&h, f->view.trees + tree, sizeof(h), cudaMemcpyDeviceToHost));
  if (h.n != 0 || h.blk != 0 || h.expanded != 0) return fail(B2AZ_ESTATE, "b2az_forest_set_root: the tree has been searched already");
  if (sg) {
    CUDA_TRY(cudaMemcpy(f->view.sg_state + tree, state, want, cudaMemcpyHostToDevice));
    if (hist_count) CUDA_TRY(cudaMemcpy(f->view.sg_hist + (size_t)tree * cap, hist, hist_count * key, cudaMemcpyHostToDevice));
  } else {
    memcpy(&h.state, state, want);
    if (hist_count) CUDA_TRY(cudaMemcpy(f->view.hist + (size_t)tree * cap, hist, hist_count * key, cudaMemcpyHostToDevice));
  }
  h.hist_len = hist_count;
  CUDA_TRY(cudaMemcpy(f->view.trees + tree, &h, sizeof(h), cudaMemcpyHostToDevice));
  return 0;
}
// the inverse of b2az_forest_set_root: the root position of `tree` as the host GameState classes hold it (GameData::gs)
int b2az_forest_get_root(b2az_forest* f, uint32_t tree, void* state, uint32_t state_bytes, void* hist, uint32_t hist_cap,
                         uint32_t* hist_count) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));
  using namespace b2az;
  if (!f || !state || !hist_count) return fail(B2AZ_EINVAL, "null argument");
  if (tree >= f->view.n_trees) return fail(B2AZ_EINVAL, "b2az_forest_get_root: tree out of range");
  const bool sg = f->view.sg_state != nullptr;
  const size_t want = sg ? sizeof(SGState) : sizeof(TaflState), key = sg ? sizeof(u64) : sizeof(TaflKey);
  const size_t cap = sg ? f->view.sg_hist_cap : (size_t)f->view.max_turns + 2u;
  if (state_bytes != want) return fail(B2AZ_EINVAL, "b2az_forest_get_root: state record of the wrong size");
  CUDA_TRY(cudaDeviceSynchronize());
  ForestTree h;
  CUDA_TRY(cudaMemcpy(&h, f->view.trees + tree, sizeof(h), cudaMemcpyDeviceToHost));
  if (sg) CUDA_TRY(cudaMemcpy(state, f->view.sg_state + tree, want, cudaMemcpyDeviceToHost));
  else memcpy(state, &h.state, want);
  const uint32_t n = h.hist_len;
  *hist_count = n;
  if (n > cap) return fail(B2AZ_ESTATE, "b2az_forest_get_root: corrupt history length");
  if (n) {
    if (!hist || hist_cap < n) return fail(B2AZ_EINVAL, "b2az_forest_get_root: history buffer too small");
    CUDA_TRY(cudaMemcpy(hist, sg ? (const void*)(f->view.sg_hist + (size_t)tree * cap) : (const void*)(f->view.hist + (size_t)tree * cap),
                        n * key, cudaMemcpyDeviceToHost));
  }
  return 0;
}
int b2az_forest_counts(b2az_forest* f, void* stream, uint32_t* counts_host, float* q_host, uint32_t* info_host) {
  if (f) CUDA_TRY(cudaSetDevice(f->device));  // the CUDA current device is per host thread
  using namespace b2az;
  if (!f) return fail(B2AZ_EINVAL, "null forest");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = f->view.n_trees, A = f->actions;
  u32 *dc = nullptr, *di = nullptr;
  float* dq = nullptr;
  int rc = 0;
  if (counts_host) rc = rc ? rc : dev_alloc(&dc, n * A);
  if (q_host) rc = rc ? rc : dev_alloc(&dq, n * A);
  if (info_host) rc = rc ? rc : dev_alloc(&di, n * 16);
  if (!rc) {
    FOREST_DISPATCH(f, (k_forest_counts<G_><<<forest_ctas(f), 128, 0, s>>>(f->view, dc, dq, di)));
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_forest_counts launch failed");
  }
  if (!rc && counts_host) rc = copy_d2h(counts_host, dc, n * A * 4, s);
  if (!rc && q_host) rc = copy_d2h(q_host, dq, n * A * 4, s);
  if (!rc && info_host) rc = copy_d2h(info_host, di, n * 16 * 4, s);
  if (!rc) rc = stream_sync(s);
  dev_free(dc); dev_free(dq); dev_free(di);
  return rc;
}
#endif

}  // extern "C"
