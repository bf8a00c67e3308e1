// oracle/ref_tafl_driver.cc — extern "C" driver around the UNMODIFIED reference tafl games
// (/root/reference/src/{brandubh,opentafl,tawlbwrdd}_gs.cc, compiled in place against oracle/shim/).
// TEST INFRASTRUCTURE ONLY — nothing in the product path links it.
//
// The three game files are pulled into THIS translation unit with #include: their headers define
// non-inline functions (brandubh_gs.h:85-98 operator==, tafl_helper.h:7-14 policyLocation ...), so two
// translation units that both include them cannot be linked together. Nothing is copied: the
// preprocessor reads the sources where they lie (-I$(REF)).
//
// Built by oracle/Makefile into oracle/_ref/libazref_tafl.so (git-ignored, travels to the GPU box).
//
//   azref_tafl_dims         board side, action count, canonical planes of a game
//   azref_tafl_random_game  a random legal game from the start position (moves chosen with a small
//                           deterministic generator, so the same transcript can be regenerated)
//   azref_tafl_search       a whole single-tree MCTS run (MCTS::find_leaf / process_result / update_root,
//                           mcts.cc) over one game with a caller-supplied or dumb_eval evaluator
//   azref_tafl_selfplay     the UNMODIFIED PlayManager (play_manager.cc) playing `games_to_play` tafl games one after the
//                           other in ONE slot on the calling thread (EvalType::RANDOM), after
//                           MCTS::seed_thread_rng(seed): history samples, scores and metrics
//   azref_tafl_replay       replays a transcript through GameState::play_move and records, after
//                           every move: to_bytes() board, player, turn, repetition count, scores(),
//                           valid_moves(), canonicalized()
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "brandubh_gs.cc"
#include "mcts.h"
#include "play_manager.h"
#include "opentafl_gs.cc"
#include "tawlbwrdd_gs.cc"
#include "star_gambit_gs.h"  // Star Gambit: star_gambit_gs.o is linked in (its header has no non-inline definitions)

using namespace alphazero;

namespace {
thread_local std::string g_err;
bool g_gumbel_full = false;  // azref_tafl_set_gumbel_full: MCTS(gumbel_full) / PlayParams::gumbel_full of the next runs

std::unique_ptr<GameState> make_game(int game, uint16_t max_turns) {
  switch (game) {
    case 0: return std::make_unique<brandubh_gs::BrandubhGS>(max_turns);
    case 1: return std::make_unique<opentafl_gs::OpenTaflGS>(max_turns);
    case 2: return std::make_unique<tawlbwrdd_gs::TawlbwrddGS>(max_turns);
    // Star Gambit: 10 + variant = the variant's own class, 20 + variant = StarGambitUnifiedGS pinned to it,
    // 24 = StarGambitUnifiedGS with the random variant mix (an unseedable mt19937: not a parity target)
    case 10: return std::make_unique<star_gambit_gs::StarGambitSkirmishGS>();
    case 11: return std::make_unique<star_gambit_gs::StarGambitShowdownGS>();
    case 12: return std::make_unique<star_gambit_gs::StarGambitClashGS>();
    case 13: return std::make_unique<star_gambit_gs::StarGambitBattleGS>();
    case 20: case 21: case 22: case 23: return std::make_unique<star_gambit_gs::StarGambitUnifiedGS>(game - 20);
    case 24: return std::make_unique<star_gambit_gs::StarGambitUnifiedGS>(-1);
  }
  return nullptr;
}
std::string state_bytes(int game, const GameState& gs) {
  switch (game) {
    case 0: return static_cast<const brandubh_gs::BrandubhGS&>(gs).to_bytes();
    case 1: return static_cast<const opentafl_gs::OpenTaflGS&>(gs).to_bytes();
    case 2: return static_cast<const tawlbwrdd_gs::TawlbwrddGS&>(gs).to_bytes();
    default: return gs.to_bytes();  // virtual for the Star Gambit classes (game_state.h)
  }
}
uint64_t next_u64(uint64_t& s) {  // splitmix64
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
}  // namespace

extern "C" {

const char* azref_tafl_last_error() { return g_err.c_str(); }
void azref_tafl_set_gumbel_full(int on) { g_gumbel_full = on != 0; }

int azref_tafl_dims(int game, uint32_t* side, uint32_t* actions, uint32_t* planes) {
  auto gs = make_game(game, 10);
  if (!gs) return -1;
  auto c = gs->canonicalized();
  *planes = (uint32_t)c.dimension(0);
  *side = (uint32_t)c.dimension(1);
  *actions = gs->num_moves();
  return 0;
}

// Plays uniformly random legal moves until scores() reports the end or max_len moves; returns the length.
uint32_t azref_tafl_random_game(int game, uint16_t max_turns, uint64_t seed, uint32_t max_len, uint32_t* moves_out) {
  auto gs = make_game(game, max_turns);
  if (!gs) return 0;
  uint32_t len = 0;
  std::vector<uint32_t> legal;
  while (len < max_len && !gs->scores().has_value()) {
    auto vm = gs->valid_moves();
    legal.clear();
    for (uint32_t m = 0; m < (uint32_t)vm.size(); ++m)
      if (vm(m)) legal.push_back(m);
    if (legal.empty()) break;
    const uint32_t mv = legal[next_u64(seed) % legal.size()];
    gs->play_move(mv);
    moves_out[len++] = mv;
  }
  return len;
}

// Records the state after k = 0..len moves. Array shapes: boards [len+1][3*S*S], players/reps/terminal [len+1],
// turns [len+1], scores [len+1][3], n_valid [len+1], valid [len+1][A] (may be NULL), canonical [len+1][C*S*S]
// (may be NULL). terminal = 0 (scores() == nullopt) or 1 + argmax of the one-hot score vector.
int azref_tafl_replay(int game, uint16_t max_turns, const uint32_t* moves, uint32_t len, int8_t* boards,
                      uint8_t* players, uint32_t* turns, uint8_t* reps, uint8_t* terminal, float* scores,
                      uint32_t* n_valid, uint8_t* valid, float* canonical) {
  try {
    auto gs = make_game(game, max_turns);
    if (!gs) { g_err = "unknown game"; return -1; }
    auto c0 = gs->canonicalized();
    const size_t S = (size_t)c0.dimension(1), planes = (size_t)c0.dimension(0), A = gs->num_moves();
    const size_t bb = 3 * S * S;
    for (uint32_t k = 0; k <= len; ++k) {
      if (k > 0) gs->play_move(moves[k - 1]);
      const std::string bytes = state_bytes(game, *gs);
      std::memcpy(boards + k * bb, bytes.data(), bb);
      players[k] = gs->current_player();
      turns[k] = gs->current_turn();
      reps[k] = (uint8_t)bytes[bb + 5];  // board | turn u16 | max_turns u16 | player i8 | rep_count u8
      auto sc = gs->scores();
      terminal[k] = 0;
      for (int j = 0; j < 3; ++j) scores[k * 3 + j] = 0.0f;
      if (sc.has_value()) {
        for (int j = 0; j < 3; ++j) {
          scores[k * 3 + j] = (*sc)(j);
          if ((*sc)(j) == 1.0f && terminal[k] == 0) terminal[k] = (uint8_t)(j + 1);
        }
      }
      auto vm = gs->valid_moves();
      uint32_t nv = 0;
      for (size_t m = 0; m < A; ++m) nv += vm(m) ? 1u : 0u;
      n_valid[k] = nv;
      if (valid) std::memcpy(valid + k * A, vm.data(), A);
      if (canonical) {
        auto c = gs->canonicalized();
        std::memcpy(canonical + k * planes * S * S, c.data(), planes * S * S * sizeof(float));
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// One arbitrary position (the 7-argument constructors, e.g. brandubh_gs.h:124-151, with an empty repetition
// table): scores / valid_moves / canonicalized of the position itself, then optionally play_move(move) and the
// board afterwards. boards int8[3][S][S]. Returns 0, or -1 when the reference threw (message in last_error).
int azref_tafl_position(int game, const int8_t* board, int8_t player, uint16_t turn, uint16_t max_turns, uint8_t rep,
                        uint32_t move, uint8_t* terminal, uint32_t* n_valid, uint8_t* valid, float* canonical,
                        int8_t* board_out) {
  try {
    std::unique_ptr<GameState> gs;
    if (game == 0) {
      brandubh_gs::BoardTensor b{};
      std::memcpy(b.data(), board, b.size());
      gs = std::make_unique<brandubh_gs::BrandubhGS>(
          b, player, turn, max_turns, rep,
          absl::flat_hash_map<const std::shared_ptr<brandubh_gs::RepetitionKey>, uint8_t>{},
          std::make_shared<absl::flat_hash_set<brandubh_gs::RepetitionKeyWrapper>>());
    } else if (game == 1) {
      opentafl_gs::BoardTensor b{};
      std::memcpy(b.data(), board, b.size());
      gs = std::make_unique<opentafl_gs::OpenTaflGS>(
          b, player, turn, max_turns, rep,
          absl::flat_hash_map<const std::shared_ptr<opentafl_gs::RepetitionKey>, uint8_t>{},
          std::make_shared<absl::flat_hash_set<opentafl_gs::RepetitionKeyWrapper>>());
    } else if (game == 2) {
      tawlbwrdd_gs::BoardTensor b{};
      std::memcpy(b.data(), board, b.size());
      gs = std::make_unique<tawlbwrdd_gs::TawlbwrddGS>(
          b, player, turn, max_turns, rep,
          absl::flat_hash_map<const std::shared_ptr<tawlbwrdd_gs::RepetitionKey>, uint8_t>{},
          std::make_shared<absl::flat_hash_set<tawlbwrdd_gs::RepetitionKeyWrapper>>());
    } else {
      g_err = "unknown game";
      return -1;
    }
    auto sc = gs->scores();
    *terminal = 0;
    if (sc.has_value())
      for (int j = 0; j < 3; ++j)
        if ((*sc)(j) == 1.0f && *terminal == 0) *terminal = (uint8_t)(j + 1);
    auto vm = gs->valid_moves();
    uint32_t nv = 0;
    for (long m = 0; m < (long)vm.size(); ++m) nv += vm(m) ? 1u : 0u;
    *n_valid = nv;
    if (valid) std::memcpy(valid, vm.data(), vm.size());
    if (canonical) {
      auto c = gs->canonicalized();
      std::memcpy(canonical, c.data(), c.size() * sizeof(float));
    }
    if (move != 0xFFFFFFFFu && board_out) {
      gs->play_move(move);
      const std::string bytes = state_bytes(game, *gs);
      auto c = gs->canonicalized();
      const size_t S = (size_t)c.dimension(1);
      std::memcpy(board_out, bytes.data(), 3 * S * S);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- Star Gambit (games 10-13 the variants' own classes, 20-23 StarGambitUnifiedGS pinned to a variant)
// azref_sg_replay: the state after k = 0..len moves of a transcript. Per position: the inner state's to_bytes()
// (star_gambit_gs.cc:2246-2288; for the Unified classes the 25-byte prefix is skipped; the key history after hist_len is
// replaced by its 64-bit FNV-1a) padded to `bytes_stride`, its
// length, player, turn, terminal (0, 1 + winner, 3 draw), scores [3], n_valid, valid [A] (may be NULL),
// canonical [C*D*D] (may be NULL).
int azref_sg_replay(int game, const uint32_t* moves, uint32_t len, uint32_t bytes_stride, uint8_t* bytes_out,
                    uint32_t* bytes_len, uint8_t* players, uint32_t* turns, uint8_t* terminal, float* scores,
                    uint32_t* n_valid, uint8_t* valid, float* canonical) {
  try {
    auto gs = make_game(game, 0);
    if (!gs || game < 10) { g_err = "unknown game"; return -1; }
    const size_t A = gs->num_moves();
    auto c0 = gs->canonicalized();
    const size_t C = (size_t)c0.size();
    for (uint32_t k = 0; k <= len; ++k) {
      if (k > 0) gs->play_move(moves[k - 1]);
      std::string b = gs->to_bytes();
      if (game >= 20) b = b.substr(25);
      {  // the key history can be thousands of entries long: keep [units .. hist_len] and replace the keys by their FNV-1a
        uint32_t nu = 0;
        std::memcpy(&nu, b.data(), 4);
        const size_t fixed = 4 + 9 * (size_t)nu + 8 + 1 + 4 + 3 + 4;
        uint64_t h = 0xcbf29ce484222325ULL;
        for (size_t i = fixed; i < b.size(); ++i) h = (h ^ (uint8_t)b[i]) * 0x100000001b3ULL;
        b.resize(fixed);
        b.append(reinterpret_cast<const char*>(&h), 8);
      }
      if (b.size() > bytes_stride) { g_err = "bytes_stride too small"; return -1; }
      std::memset(bytes_out + (size_t)k * bytes_stride, 0, bytes_stride);
      std::memcpy(bytes_out + (size_t)k * bytes_stride, b.data(), b.size());
      bytes_len[k] = (uint32_t)b.size();
      players[k] = gs->current_player();
      turns[k] = gs->current_turn();
      auto sc = gs->scores();
      terminal[k] = 0;
      for (int j = 0; j < 3; ++j) scores[k * 3 + j] = 0.0f;
      if (sc.has_value()) {
        terminal[k] = 4;
        for (int j = 0; j < 3; ++j) {
          scores[k * 3 + j] = (*sc)(j);
          if ((*sc)(j) == 1.0f && terminal[k] == 4) terminal[k] = (uint8_t)(j + 1);
        }
      }
      auto vm = gs->valid_moves();
      uint32_t nv = 0;
      for (size_t m = 0; m < A; ++m) nv += vm(m) ? 1u : 0u;
      n_valid[k] = nv;
      if (valid) std::memcpy(valid + (size_t)k * A, vm.data(), A);
      if (canonical) {
        auto c = gs->canonicalized();
        std::memcpy(canonical + (size_t)k * C, c.data(), C * sizeof(float));
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// get_units() / get_fire_info(move) of the position after the transcript (py_wrapper.cc:589-670): units_out
// [n][8] ints (player, type, slot, hp, anchor_q, anchor_r, facing, moves_left), fire_out [5] ints.
int azref_sg_units(int game, const uint32_t* moves, uint32_t len, uint32_t fire_move, int32_t* units_out, uint32_t cap,
                   int32_t* fire_out) {
  try {
    auto gs = make_game(game, 0);
    if (!gs || game < 10) { g_err = "unknown game"; return -1; }
    for (uint32_t k = 0; k < len; ++k) gs->play_move(moves[k]);
    std::vector<star_gambit_gs::UnitInfo> us;
    star_gambit_gs::FireInfo fi{};
    using namespace star_gambit_gs;
    switch (game) {
      case 10: us = static_cast<StarGambitSkirmishGS&>(*gs).get_units(); fi = static_cast<StarGambitSkirmishGS&>(*gs).get_fire_info(fire_move); break;
      case 11: us = static_cast<StarGambitShowdownGS&>(*gs).get_units(); fi = static_cast<StarGambitShowdownGS&>(*gs).get_fire_info(fire_move); break;
      case 12: us = static_cast<StarGambitClashGS&>(*gs).get_units(); fi = static_cast<StarGambitClashGS&>(*gs).get_fire_info(fire_move); break;
      case 13: us = static_cast<StarGambitBattleGS&>(*gs).get_units(); fi = static_cast<StarGambitBattleGS&>(*gs).get_fire_info(fire_move); break;
      default: us = static_cast<StarGambitUnifiedGS&>(*gs).get_units(); fi = static_cast<StarGambitUnifiedGS&>(*gs).get_fire_info(fire_move); break;
    }
    uint32_t n = 0;
    for (const auto& u : us) {
      if (n >= cap) break;
      int32_t* o = units_out + (size_t)n * 8;
      o[0] = u.player; o[1] = u.type; o[2] = u.slot; o[3] = u.hp; o[4] = u.anchor_q; o[5] = u.anchor_r; o[6] = u.facing; o[7] = u.moves_left;
      ++n;
    }
    fire_out[0] = fi.has_target; fire_out[1] = fi.target_player; fire_out[2] = fi.target_type; fire_out[3] = fi.target_slot; fire_out[4] = fi.damage;
    return (int)us.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// One single-tree search run through the reference's MCTS class (mcts.h:50-150), exactly the way Python drives it
// (py_wrapper.cc:192-220; play.py): seed the thread's generator, then for every move `sims` x (find_leaf ->
// evaluate -> process_result(root_noise_enabled = false)), record counts() / root_q_values(), play the most
// visited move (lowest id on ties) with update_root + play_move. eval_kind 0: `cb(canonical, v3, piA, user)`
// supplies the evaluation; 1: dumb_eval (game_state.h:160-173). Stops when the game is over; returns the number
// of moves searched, or -1 when the reference threw.
typedef void (*azref_eval_fn)(const float* canonical, float* v3, float* pi, void* user);
// gumbel_m > 0: Gumbel root search (MCTS built with gumbel_enabled; set_gumbel_num_sims(sims) before every move's
// search like PlayManager, play_manager.cc:531-539); the move played is gumbel_final_action() and
// gumbel_improved_policy() is recorded in policy_out [n_moves][A] (may be NULL).
int azref_tafl_search(int game, uint16_t max_turns, uint64_t seed, float cpuct, float fpu_reduction, int root_fpu_zero,
                      uint32_t n_moves, uint32_t sims, int eval_kind, azref_eval_fn cb, void* user, uint32_t* counts_out,
                      float* q_out, uint32_t* moves_out, uint32_t* depth_sum_out, uint32_t gumbel_m, float gumbel_c_visit,
                      float gumbel_c_scale, float* policy_out, float epsilon, float root_policy_temp, int shaped_dirichlet,
                      uint32_t batch_width, float act_temp, float* probs_out, float pruned_temp) {
  try {
    auto gs = make_game(game, max_turns);
    if (!gs) { g_err = "unknown game"; return -1; }
    const uint32_t A = gs->num_moves();
    MCTS::seed_thread_rng(seed);
    // epsilon > 0: process_result(root_noise_enabled = true), and after every move the reused root gets the root
    // temperature again and fresh noise, as PlayManager does (play_manager.cc:546-553)
    // relative_values: Star Gambit evaluators answer in the mover's frame (mcts.cc:522-524)
    MCTS mcts{cpuct, 2, A, epsilon, root_policy_temp, fpu_reduction, gs->relative_values(), root_fpu_zero != 0, shaped_dirichlet != 0, gumbel_m > 0,
              gumbel_m > 0 ? gumbel_m : 16u, gumbel_c_visit, gumbel_c_scale, g_gumbel_full};
    uint32_t played = 0;
    for (uint32_t m = 0; m < n_moves; ++m) {
      if (gs->scores().has_value()) break;
      if (gumbel_m > 0) mcts.set_gumbel_num_sims(sims);
      // batch_width > 0: WU-UCT — rounds of `batch_width` find_leaf_batched calls, then their process_result_batched
      // in the same order, then reset_batch (how play.py / mcts_analysis.py drive a tree with a batched net)
      for (uint32_t r = 0; batch_width > 0 && r < sims / batch_width; ++r) {
        std::vector<Vector<float>> vs, pis;
        for (uint32_t i = 0; i < batch_width; ++i) {
          auto leaf = mcts.find_leaf_batched(*gs);
          Vector<float> v{3}, pi{A};
          if (eval_kind == 1) {
            auto [vv, pp] = dumb_eval(*leaf);
            v = vv; pi = pp;
          } else {
            auto c = leaf->canonicalized();
            cb(c.data(), v.data(), pi.data(), user);
          }
          vs.push_back(v);
          pis.push_back(pi);
        }
        for (uint32_t i = 0; i < batch_width; ++i) mcts.process_result_batched(*gs, i, vs[i], pis[i], epsilon > 0.0f);
        mcts.reset_batch();
      }
      for (uint32_t i = 0; batch_width == 0 && i < sims; ++i) {
        auto leaf = mcts.find_leaf(*gs);
        Vector<float> v{3}, pi{A};
        if (eval_kind == 1) {
          auto [vv, pp] = dumb_eval(*leaf);
          v = vv; pi = pp;
        } else {
          auto c = leaf->canonicalized();
          cb(c.data(), v.data(), pi.data(), user);
        }
        mcts.process_result(*gs, v, pi, epsilon > 0.0f);
      }
      auto counts = mcts.counts();
      auto q = mcts.root_q_values();
      std::memcpy(counts_out + (size_t)m * A, counts.data(), A * 4);
      std::memcpy(q_out + (size_t)m * A, q.data(), A * 4);
      if (depth_sum_out) depth_sum_out[m] = (uint32_t)(mcts.avg_leaf_depth() * (float)mcts.depth() + 0.5f);
      uint32_t best = 0;
      if (gumbel_m > 0) {
        if (policy_out) {
          auto ip = mcts.gumbel_improved_policy();
          std::memcpy(policy_out + (size_t)m * A, ip.data(), A * 4);
        }
        best = mcts.gumbel_final_action();
      } else if (act_temp >= 0.0f) {  // PlayManager's PUCT acting rule: pick_move(probs(temp)) (play_manager.cc:372-381)
        auto pr = mcts.probs(act_temp);
        if (probs_out) std::memcpy(probs_out + (size_t)m * A, pr.data(), A * 4);
        if (policy_out) {  // the training target under policy_target_pruning (play_manager.cc:418-421)
          auto pp = mcts.probs_pruned(pruned_temp);
          std::memcpy(policy_out + (size_t)m * A, pp.data(), A * 4);
        }
        best = MCTS::pick_move(pr);
      } else {
        for (uint32_t a = 1; a < A; ++a)
          if (counts(a) > counts(best)) best = a;
      }
      moves_out[m] = best;
      mcts.update_root(*gs, best);
      gs->play_move(best);
      if (mcts.root_n() > 0) {
        mcts.apply_root_policy_temp();
        if (epsilon > 0.0f) mcts.add_root_noise();
      }
      ++played;
    }
    return (int)played;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// GameState::symmetries(PlayHistory) (tafl_helper::eightSym through {Brandubh,OpenTafl,Tawlbwrdd}GS::symmetries):
// canon [P][S][S], v [3], pi [A] -> canon_out [8][P][S][S], v_out [8][3], pi_out [8][A]; returns the count (8).
int azref_tafl_symmetries(int game, const float* canon, const float* v, const float* pi, float* canon_out, float* v_out,
                          float* pi_out) {
  auto gs = make_game(game, 10);
  if (!gs) return -1;
  auto c0 = gs->canonicalized();
  const long P = c0.dimension(0), S = c0.dimension(1);
  const size_t C = (size_t)(P * S * S), A = gs->num_moves();
  PlayHistory base;
  base.canonical = Tensor<float, 3>(P, S, S);
  std::memcpy(base.canonical.data(), canon, C * 4);
  base.v = Vector<float>{3};
  std::memcpy(base.v.data(), v, 12);
  base.pi = Vector<float>{(long)A};
  std::memcpy(base.pi.data(), pi, A * 4);
  auto syms = gs->symmetries(base);
  for (size_t i = 0; i < syms.size(); ++i) {
    std::memcpy(canon_out + i * C, syms[i].canonical.data(), C * 4);
    std::memcpy(v_out + i * 3, syms[i].v.data(), 12);
    std::memcpy(pi_out + i * A, syms[i].pi.data(), A * 4);
  }
  return (int)syms.size();
}

// One slot of PlayManager::play (play_manager.cc:258-600) over a tafl game: concurrent_games = 1, games_to_play = K,
// EvalType::RANDOM for both seats, on the calling thread after MCTS::seed_thread_rng(seed) — every draw of the run
// (child shuffles of BOTH seats' trees, Dirichlet / Gumbel noise, pick_move) then comes from that one stream.
// History rows come out in history_ order (a game's samples last move first, game after game).
struct AzRefTaflSpCfg {
  uint32_t games_to_play, visits;
  float cpuct, fpu_reduction, epsilon, mcts_root_temp;
  float start_temp, final_temp, temp_decay_half_life;
  uint32_t gumbel_m;
  float gumbel_c_visit, gumbel_c_scale;
  uint8_t root_fpu_zero, shaped_dirichlet, policy_target_pruning, gumbel_enabled, tree_reuse, history_enabled, pad_[2];
  // per-seat budgets, playout-cap randomisation and resignation (0 / false = the reference's defaults)
  uint32_t seat_visits[2], seat_cap_visits[2];
  uint32_t playout_cap_depth;
  float playout_cap_percent, resign_percent, resign_playthrough_percent;
  uint8_t playout_cap_randomization, fast_search_uses_gumbel, pad2_[2];
  // two model groups under ONE seat permutation (model_groups = {0, 1}, seat_perms = {seat_perm}, mcts_visits = group_visits:
  // seat_visits_[0][s] = mcts_visits_[seat_perm[s]], play_manager.cc:70-80) when has_perm != 0
  uint8_t has_perm, seat_perm[2], has_seat;
  uint32_t group_visits[2];
  // per-seat search settings and the per-seat resign rule (PlayParams::seat_* with one permutation) when has_seat != 0
  float seat_epsilon[2], seat_root_temp[2];
  uint8_t seat_root_fpu_zero[2], seat_gumbel_enabled[2];
  uint32_t seat_gumbel_m[2];
  float seat_gumbel_c_visit[2], seat_gumbel_c_scale[2];
  float seat_resign_threshold[2];
  uint32_t seat_resign_consecutive[2];
};
int azref_tafl_selfplay(int game, uint16_t max_turns, uint64_t seed, const AzRefTaflSpCfg* c, uint32_t hist_cap,
                        float* canon_out, float* v_out, float* pi_out, uint32_t* n_hist, float* scores3,
                        uint32_t* games_completed, float* metrics4) {
  try {
    auto gs = make_game(game, max_turns);
    if (!gs) { g_err = "unknown game"; return -1; }
    PlayParams p{};
    p.games_to_play = c->games_to_play;
    p.concurrent_games = 1;
    p.max_batch_size = 1;
    p.max_cache_size = 0;
    p.mcts_visits = {c->visits, c->visits};
    p.cpuct = c->cpuct;
    p.fpu_reduction = c->fpu_reduction;
    p.root_fpu_zero = c->root_fpu_zero != 0;
    p.epsilon = c->epsilon;
    p.mcts_root_temp = c->mcts_root_temp;
    p.shaped_dirichlet = c->shaped_dirichlet != 0;
    p.policy_target_pruning = c->policy_target_pruning != 0;
    p.start_temp = c->start_temp;
    p.final_temp = c->final_temp;
    p.temp_decay_half_life = c->temp_decay_half_life;
    p.gumbel_enabled = c->gumbel_enabled != 0;
    p.gumbel_m = c->gumbel_m;
    p.gumbel_c_visit = c->gumbel_c_visit;
    p.gumbel_c_scale = c->gumbel_c_scale;
    p.gumbel_full = g_gumbel_full;
    p.tree_reuse = c->tree_reuse != 0;
    p.history_enabled = c->history_enabled != 0;
    p.self_play = true;
    p.model_groups = {0, 0};
    p.playout_cap_randomization = c->playout_cap_randomization != 0;
    if (c->playout_cap_depth) p.playout_cap_depth = c->playout_cap_depth;
    p.playout_cap_percent = c->playout_cap_percent;
    p.fast_search_uses_gumbel = c->fast_search_uses_gumbel != 0;
    p.resign_percent = c->resign_percent;
    p.resign_playthrough_percent = c->resign_playthrough_percent;
    if (c->seat_visits[0] && c->seat_visits[1]) p.seat_visits = {{c->seat_visits[0], c->seat_visits[1]}};
    if (c->seat_cap_visits[0] && c->seat_cap_visits[1]) p.seat_cap_visits = {{c->seat_cap_visits[0], c->seat_cap_visits[1]}};
    p.eval_type = {EvalType::RANDOM, EvalType::RANDOM};
    if (c->has_perm) {
      p.model_groups = {0, 1};
      p.mcts_visits = {c->group_visits[0], c->group_visits[1]};
      p.seat_perms = {{c->seat_perm[0], c->seat_perm[1]}};
    }
    if (c->has_seat) {
      p.seat_epsilon = {{c->seat_epsilon[0], c->seat_epsilon[1]}};
      p.seat_mcts_root_temp = {{c->seat_root_temp[0], c->seat_root_temp[1]}};
      p.seat_root_fpu_zero = {{c->seat_root_fpu_zero[0], c->seat_root_fpu_zero[1]}};
      p.seat_gumbel_enabled = {{c->seat_gumbel_enabled[0], c->seat_gumbel_enabled[1]}};
      p.seat_gumbel_m = {{c->seat_gumbel_m[0], c->seat_gumbel_m[1]}};
      p.seat_gumbel_c_visit = {{c->seat_gumbel_c_visit[0], c->seat_gumbel_c_visit[1]}};
      p.seat_gumbel_c_scale = {{c->seat_gumbel_c_scale[0], c->seat_gumbel_c_scale[1]}};
      p.seat_resign_threshold = {{c->seat_resign_threshold[0], c->seat_resign_threshold[1]}};
      p.seat_resign_consecutive = {{c->seat_resign_consecutive[0], c->seat_resign_consecutive[1]}};
    }
    PlayManager pm{std::move(gs), p};
    MCTS::seed_thread_rng(seed);
    pm.play();
    const size_t A = pm.params().mcts_visits.size() ? (size_t)make_game(game, max_turns)->num_moves() : 0;
    uint32_t n = 0;
    while (auto h = pm.pop_hist()) {
      if (n < hist_cap) {
        const size_t C = (size_t)h->canonical.size();
        if (canon_out) std::memcpy(canon_out + (size_t)n * C, h->canonical.data(), C * 4);
        if (v_out) std::memcpy(v_out + (size_t)n * 3, h->v.data(), 12);
        if (pi_out) std::memcpy(pi_out + (size_t)n * A, h->pi.data(), A * 4);
      }
      ++n;
    }
    *n_hist = n;
    auto sc = c->has_perm ? pm.perm_scores(0) : pm.scores();
    for (int i = 0; i < 3; ++i) scores3[i] = sc(i);
    *games_completed = c->has_perm ? pm.perm_games_completed(0) : pm.games_completed();
    metrics4[0] = pm.avg_game_length();
    metrics4[1] = pm.avg_leaf_depth();
    metrics4[2] = pm.avg_valid_moves();
    metrics4[3] = pm.avg_search_entropy();
    // (the array is 10 floats long on the Python side) fast-search metrics and resign scores
    metrics4[4] = pm.fast_avg_leaf_depth();
    metrics4[5] = pm.fast_avg_search_entropy();
    auto rs = pm.resign_scores();
    for (int i = 0; i < 3; ++i) metrics4[6 + i] = rs(i);
    metrics4[9] = pm.avg_moves_per_turn();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
