// oracle/ref_driver.cc — extern "C" driver around the UNMODIFIED reference sources
// (/root/reference/src/{mcts,play_manager,connect4_gs,game_state}.cc, compiled in place against
// oracle/shim/) so Python tests and bench.py's cpu_baseline / --impl reference legs can run the
// real reference through ctypes.  TEST INFRASTRUCTURE ONLY — nothing in the product path links it.
//
// Built by oracle/Makefile into oracle/_ref/libazref.so (git-ignored, travels to the GPU box).
//
// What it exposes (each maps 1:1 onto a reference entry point; no logic of its own beyond the
// lock-step harness described in SURVEY.md Appendix A "Scheduling order"):
//   azref_c4_*     Connect4GS   (connect4_gs.h:24-92)
//   azref_mcts_*   MCTS         (mcts.h:50-150)
//   azref_pm_*     PlayManager  (play_manager.h:159-366) + the build_batch / build_history_batch
//                  lambdas of py_wrapper.cc:393-424, 449-504 restated without pybind
//   azref_cache_*  S3FIFOCache  (s3fifo_cache.h:15-227)
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "connect4_gs.h"
#include "pcg/pcg_random.hpp"
#include "mcts.h"
#include "play_manager.h"
#include "s3fifo_cache.h"

using namespace alphazero;
using connect4_gs::Connect4GS;

namespace {
thread_local std::string g_err;
template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
}  // namespace

extern "C" {

const char* azref_last_error() { return g_err.c_str(); }
void azref_seed_thread_rng(uint64_t seed) { MCTS::seed_thread_rng(seed); }

// ----------------------------------------------------------------------------- Connect4GS
void* azref_c4_new() { return new Connect4GS(); }
void* azref_c4_from_board(const int8_t* board84, int8_t player, int32_t turn) {
  connect4_gs::BoardTensor b{};
  std::memcpy(b.data(), board84, 84);
  return new Connect4GS(b, player, turn);
}
void* azref_c4_copy(void* gs) { return static_cast<Connect4GS*>(gs)->copy().release(); }
void azref_c4_free(void* gs) { delete static_cast<GameState*>(gs); }
int azref_c4_play(void* gs, uint32_t move) {
  return guarded([&] { static_cast<GameState*>(gs)->play_move(move); });
}
void azref_c4_valid(void* gs, uint8_t* out7) {
  auto v = static_cast<GameState*>(gs)->valid_moves();
  std::memcpy(out7, v.data(), 7);
}
// returns 1 and fills out3 if terminal, else 0
int azref_c4_scores(void* gs, float* out3) {
  auto s = static_cast<GameState*>(gs)->scores();
  if (!s.has_value()) return 0;
  std::memcpy(out3, s->data(), 3 * sizeof(float));
  return 1;
}
void azref_c4_canonical(void* gs, float* out168) {
  auto c = static_cast<GameState*>(gs)->canonicalized();
  std::memcpy(out168, c.data(), 168 * sizeof(float));
}
int azref_c4_player(void* gs) { return static_cast<GameState*>(gs)->current_player(); }
uint32_t azref_c4_turn(void* gs) { return static_cast<GameState*>(gs)->current_turn(); }
void azref_c4_to_bytes(void* gs, uint8_t* out89) {
  auto s = static_cast<GameState*>(gs)->to_bytes();
  std::memcpy(out89, s.data(), 89);
}
int azref_c4_equal(void* a, void* b) { return *static_cast<GameState*>(a) == *static_cast<GameState*>(b); }
uint64_t azref_c4_hash(void* gs) { return hash_game_state(*static_cast<GameState*>(gs)); }
// symmetries(): writes the mirrored sample (index 1) of (canonical, v, pi)
void azref_c4_mirror(void* gs, const float* canon168, const float* v3, const float* pi7, float* canon_out,
                     float* v_out, float* pi_out) {
  PlayHistory base;
  base.canonical = Tensor<float, 3>(4, 6, 7);
  std::memcpy(base.canonical.data(), canon168, 168 * sizeof(float));
  base.v = Vector<float>{3};
  std::memcpy(base.v.data(), v3, 12);
  base.pi = Vector<float>{7};
  std::memcpy(base.pi.data(), pi7, 28);
  auto syms = static_cast<GameState*>(gs)->symmetries(base);
  std::memcpy(canon_out, syms[1].canonical.data(), 168 * sizeof(float));
  std::memcpy(v_out, syms[1].v.data(), 12);
  std::memcpy(pi_out, syms[1].pi.data(), 28);
}

// ----------------------------------------------------------------------------- MCTS (single tree)
struct AzRefMctsCfg {
  float cpuct;
  uint32_t num_players;
  uint32_t num_moves;
  float epsilon;
  float root_policy_temp;
  float fpu_reduction;
  uint8_t relative_values;
  uint8_t root_fpu_zero;
  uint8_t shaped_dirichlet;
  uint8_t gumbel_enabled;
  uint32_t gumbel_m;
  float gumbel_c_visit;
  float gumbel_c_scale;
  uint8_t gumbel_full;
};
struct RefTree {
  MCTS mcts;
  std::unique_ptr<GameState> leaf;
};
void* azref_mcts_new(const AzRefMctsCfg* c) {
  return new RefTree{MCTS{c->cpuct, c->num_players, c->num_moves, c->epsilon, c->root_policy_temp, c->fpu_reduction,
                          (bool)c->relative_values, (bool)c->root_fpu_zero, (bool)c->shaped_dirichlet,
                          (bool)c->gumbel_enabled, c->gumbel_m, c->gumbel_c_visit, c->gumbel_c_scale,
                          (bool)c->gumbel_full},
                     nullptr};
}
void azref_mcts_free(void* t) { delete static_cast<RefTree*>(t); }
// find_leaf: keeps the leaf state inside the handle; returns it (borrowed) for azref_c4_* queries
void* azref_mcts_find_leaf(void* t, void* gs) {
  auto* rt = static_cast<RefTree*>(t);
  rt->leaf = rt->mcts.find_leaf(*static_cast<GameState*>(gs));
  return rt->leaf.get();
}
void azref_mcts_process_result(void* t, void* gs, float* v, uint32_t nv, float* pi, uint32_t npi, int root_noise) {
  auto* rt = static_cast<RefTree*>(t);
  Vector<float> vv{nv}, pp{npi};
  std::memcpy(vv.data(), v, nv * 4);
  std::memcpy(pp.data(), pi, npi * 4);
  rt->mcts.process_result(*static_cast<GameState*>(gs), vv, pp, root_noise != 0);
  std::memcpy(v, vv.data(), nv * 4);  // process_result mutates value (Appendix C.11)
}
int azref_mcts_update_root(void* t, void* gs, uint32_t move) {
  return guarded([&] { static_cast<RefTree*>(t)->mcts.update_root(*static_cast<GameState*>(gs), move); });
}
void azref_mcts_counts(void* t, uint32_t* out) {
  auto c = static_cast<RefTree*>(t)->mcts.counts();
  std::memcpy(out, c.data(), c.size() * 4);
}
void azref_mcts_root_q(void* t, float* out) {
  auto c = static_cast<RefTree*>(t)->mcts.root_q_values();
  std::memcpy(out, c.data(), c.size() * 4);
}
void azref_mcts_probs(void* t, float temp, float* out) {
  auto c = static_cast<RefTree*>(t)->mcts.probs(temp);
  std::memcpy(out, c.data(), c.size() * 4);
}
void azref_mcts_probs_pruned(void* t, float temp, float* out) {
  auto c = static_cast<RefTree*>(t)->mcts.probs_pruned(temp);
  std::memcpy(out, c.data(), c.size() * 4);
}
void azref_mcts_root_value(void* t, float* out3) {
  auto c = static_cast<RefTree*>(t)->mcts.root_value();
  std::memcpy(out3, c.data(), 12);
}
uint32_t azref_mcts_depth(void* t) { return static_cast<RefTree*>(t)->mcts.depth(); }
uint32_t azref_mcts_root_n(void* t) { return static_cast<RefTree*>(t)->mcts.root_n(); }
float azref_mcts_avg_leaf_depth(void* t) { return static_cast<RefTree*>(t)->mcts.avg_leaf_depth(); }
float azref_mcts_entropy(void* t) { return static_cast<RefTree*>(t)->mcts.normalized_root_entropy(); }
void azref_mcts_apply_root_policy_temp(void* t) { static_cast<RefTree*>(t)->mcts.apply_root_policy_temp(); }
void azref_mcts_add_root_noise(void* t) { static_cast<RefTree*>(t)->mcts.add_root_noise(); }
void azref_mcts_set_gumbel_num_sims(void* t, uint32_t n) { static_cast<RefTree*>(t)->mcts.set_gumbel_num_sims(n); }
void azref_mcts_gumbel_improved_policy(void* t, float* out) {
  auto c = static_cast<RefTree*>(t)->mcts.gumbel_improved_policy();
  std::memcpy(out, c.data(), c.size() * 4);
}
uint32_t azref_mcts_gumbel_final_action(void* t) { return static_cast<RefTree*>(t)->mcts.gumbel_final_action(); }
int azref_pick_move(const float* p, uint32_t n, uint32_t* out) {
  return guarded([&] {
    Vector<float> pp{n};
    std::memcpy(pp.data(), p, n * 4);
    *out = MCTS::pick_move(pp);
  });
}

// ----------------------------------------------------------------------------- PlayManager
struct AzRefPlayCfg {
  uint32_t games_to_play;
  uint32_t concurrent_games;
  uint32_t max_batch_size;
  uint32_t max_cache_size;
  uint32_t cache_shards;
  uint32_t queue_shards;
  uint32_t mcts_visits[2];
  float cpuct;
  float start_temp;
  float final_temp;
  float temp_decay_half_life;
  uint8_t history_enabled;
  uint8_t self_play;
  uint8_t tree_reuse;
  uint8_t playout_cap_randomization;
  float epsilon;
  float mcts_root_temp;
  uint32_t playout_cap_depth;
  float playout_cap_percent;
  float fpu_reduction;
  uint8_t root_fpu_zero;
  uint8_t shaped_dirichlet;
  uint8_t policy_target_pruning;
  uint8_t gumbel_enabled;
  uint32_t gumbel_m;
  float gumbel_c_visit;
  float gumbel_c_scale;
  uint8_t gumbel_full;
  uint8_t fast_search_uses_gumbel;
  uint8_t eval_type;  // EvalType for every seat: 0 NN, 1 RANDOM, 2 PLAYOUT
  uint8_t pad_;
  float resign_percent;
  float resign_playthrough_percent;
  // model groups / seat permutations (play_manager.h:117-120): used when has_groups != 0
  uint8_t has_groups;
  uint8_t model_groups[2];
  uint8_t n_seat_perms;
  uint8_t seat_perms[8][2];
  uint8_t group_eval[2];  // EvalType per model group when has_groups (eval_types_[group], play_manager.cc:578)
  uint8_t pad2_[2];
};

struct RefPM {
  std::unique_ptr<PlayManager> pm;
  AzRefPlayCfg cfg;
  std::vector<std::thread> workers;
};

static PlayParams to_params(const AzRefPlayCfg& c) {
  PlayParams p{};
  p.games_to_play = c.games_to_play;
  p.concurrent_games = c.concurrent_games;
  p.max_batch_size = c.max_batch_size;
  p.max_cache_size = c.max_cache_size;
  p.cache_shards = static_cast<uint8_t>(c.cache_shards);
  p.queue_shards = static_cast<uint8_t>(c.queue_shards);
  p.mcts_visits = {c.mcts_visits[0], c.mcts_visits[1]};
  p.cpuct = c.cpuct;
  p.start_temp = c.start_temp;
  p.final_temp = c.final_temp;
  p.temp_decay_half_life = c.temp_decay_half_life;
  p.history_enabled = c.history_enabled;
  p.self_play = c.self_play;
  // self_play(): every seat is the same network => one model group (game_runner.py:773-787, 2053-2055)
  if (c.self_play) p.model_groups = {0, 0};
  p.tree_reuse = c.tree_reuse;
  p.epsilon = c.epsilon;
  p.mcts_root_temp = c.mcts_root_temp;
  p.playout_cap_randomization = c.playout_cap_randomization;
  p.playout_cap_depth = c.playout_cap_depth;
  p.playout_cap_percent = c.playout_cap_percent;
  p.fpu_reduction = c.fpu_reduction;
  p.root_fpu_zero = c.root_fpu_zero;
  p.shaped_dirichlet = c.shaped_dirichlet;
  p.policy_target_pruning = c.policy_target_pruning;
  p.gumbel_enabled = c.gumbel_enabled;
  p.gumbel_m = c.gumbel_m;
  p.gumbel_c_visit = c.gumbel_c_visit;
  p.gumbel_c_scale = c.gumbel_c_scale;
  p.gumbel_full = c.gumbel_full;
  p.fast_search_uses_gumbel = c.fast_search_uses_gumbel;
  p.resign_percent = c.resign_percent;
  p.resign_playthrough_percent = c.resign_playthrough_percent;
  if (c.eval_type != 0) p.eval_type = {static_cast<EvalType>(c.eval_type), static_cast<EvalType>(c.eval_type)};
  if (c.has_groups) {
    p.model_groups = {c.model_groups[0], c.model_groups[1]};
    p.seat_perms.clear();
    for (uint8_t i = 0; i < c.n_seat_perms; ++i) p.seat_perms.push_back({c.seat_perms[i][0], c.seat_perms[i][1]});
    p.eval_type = {static_cast<EvalType>(c.group_eval[0]), static_cast<EvalType>(c.group_eval[1])};
  }
  return p;
}

void* azref_pm_new_connect4(const AzRefPlayCfg* c) {
  void* out = nullptr;
  int rc = guarded([&] {
    auto* r = new RefPM{};
    r->cfg = *c;
    r->pm = std::make_unique<PlayManager>(std::make_unique<Connect4GS>(), to_params(*c));
    out = r;
  });
  return rc == 0 ? out : nullptr;
}
// perm_scores(p) -> s3, returns perm_games_completed(p) (play_manager.h:211-217)
uint32_t azref_pm_perm_scores(void* h, uint32_t perm, float* s3) {
  auto* r = static_cast<RefPM*>(h);
  if (perm >= r->pm->num_seat_perms()) return 0xFFFFFFFFu;
  const auto& sc = r->pm->perm_scores(perm);
  for (int i = 0; i < 3; ++i) s3[i] = sc(i);
  return r->pm->perm_games_completed(perm);
}
void azref_pm_free(void* h) {
  auto* r = static_cast<RefPM*>(h);
  r->pm->stop();
  for (auto& t : r->workers) t.join();
  delete r;
}
// Run play() on the calling thread (RANDOM / PLAYOUT eval: finishes on its own).
int azref_pm_play_here(void* h, uint64_t seed, int do_seed) {
  auto* r = static_cast<RefPM*>(h);
  return guarded([&] {
    if (do_seed) MCTS::seed_thread_rng(seed);
    r->pm->play();
  });
}
// Spawn n worker threads running play(); worker k seeds its thread-local RNG with seed+k if do_seed.
void azref_pm_start_workers(void* h, uint32_t n, uint64_t seed, int do_seed) {
  auto* r = static_cast<RefPM*>(h);
  for (uint32_t k = 0; k < n; ++k) {
    r->workers.emplace_back([r, seed, do_seed, k] {
      if (do_seed) MCTS::seed_thread_rng(seed + k);
      r->pm->play();
    });
  }
}
void azref_pm_join(void* h) {
  auto* r = static_cast<RefPM*>(h);
  for (auto& t : r->workers) t.join();
  r->workers.clear();
}
void azref_pm_stop(void* h) { static_cast<RefPM*>(h)->pm->stop(); }

// number of game slots still cycling (play_manager.cc:506-513: a slot retires once games_started_
// has reached games_to_play)
static uint32_t active_slots(const RefPM* r) {
  const uint32_t completed = r->pm->games_completed();
  const uint32_t restarts_possible = r->cfg.games_to_play - r->cfg.concurrent_games;
  const uint32_t restarts = completed < restarts_possible ? completed : restarts_possible;
  return r->cfg.concurrent_games - (completed - restarts);
}
// Lock-step harness: wait until every active game sits in awaiting_inference_ (NN eval).
// returns 1 quiescent, 0 finished (remaining_games()==0), -1 timeout
int azref_pm_wait_quiescent(void* h, uint32_t timeout_ms) {
  auto* r = static_cast<RefPM*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    if (r->pm->remaining_games() == 0) return 0;
    const uint32_t act = active_slots(r);
    if (act > 0 && r->pm->awaiting_inference_count() == act && r->pm->awaiting_mcts_count() == 0) {
      // re-check after reading (games_completed may have moved between the two reads)
      if (active_slots(r) == act) return 1;
    }
    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) return -1;
    std::this_thread::yield();
  }
}
// build_batch (py_wrapper.cc:449-504) without the timing heuristics: pop up to `max` ids of `group`
// in FIFO order and copy their canonicals. Returns the count.
uint32_t azref_pm_build_batch(void* h, uint32_t group, uint32_t max, uint32_t* idx_out, float* canon_out) {
  auto* r = static_cast<RefPM*>(h);
  uint32_t n = 0;
  while (n < max) {
    auto ids = r->pm->pop_games_upto_timed(group, 0, max - n, std::chrono::microseconds{500});
    if (ids.empty()) break;
    for (auto i : ids) {
      const auto& c = r->pm->game_data(i).canonical;
      std::memcpy(canon_out + static_cast<size_t>(n) * c.size(), c.data(), c.size() * sizeof(float));
      idx_out[n++] = i;
    }
  }
  return n;
}
void azref_pm_update_inferences(void* h, uint32_t group, const uint32_t* idx, uint32_t n, const float* v, uint32_t nv,
                                const float* pi, uint32_t npi) {
  auto* r = static_cast<RefPM*>(h);
  std::vector<uint32_t> ids(idx, idx + n);
  Eigen::Ref<const Matrix<float>> vr(v, n, nv), pr(pi, n, npi);
  r->pm->update_inferences(static_cast<uint8_t>(group), ids, vr, pr);
}
// build_history_batch (py_wrapper.cc:393-424), non-blocking flavour: drain what is queued.
uint32_t azref_pm_drain_history(void* h, uint32_t max, float* canon, float* v, float* pi) {
  auto* r = static_cast<RefPM*>(h);
  uint32_t n = 0;
  while (n < max && r->pm->hist_count() > 0) {
    auto hist = r->pm->pop_hist();
    if (!hist.has_value()) break;
    std::memcpy(canon + static_cast<size_t>(n) * hist->canonical.size(), hist->canonical.data(),
                hist->canonical.size() * 4);
    std::memcpy(v + static_cast<size_t>(n) * hist->v.size(), hist->v.data(), hist->v.size() * 4);
    std::memcpy(pi + static_cast<size_t>(n) * hist->pi.size(), hist->pi.data(), hist->pi.size() * 4);
    ++n;
  }
  return n;
}
uint32_t azref_pm_hist_count(void* h) { return static_cast<RefPM*>(h)->pm->hist_count(); }
uint32_t azref_pm_games_completed(void* h) { return static_cast<RefPM*>(h)->pm->games_completed(); }
uint32_t azref_pm_remaining_games(void* h) { return static_cast<RefPM*>(h)->pm->remaining_games(); }
uint32_t azref_pm_awaiting_inference(void* h) { return static_cast<RefPM*>(h)->pm->awaiting_inference_count(); }
uint32_t azref_pm_awaiting_mcts(void* h) { return static_cast<RefPM*>(h)->pm->awaiting_mcts_count(); }
void azref_pm_scores(void* h, float* out3) {
  auto s = static_cast<RefPM*>(h)->pm->scores();
  std::memcpy(out3, s.data(), 12);
}
void azref_pm_resign_scores(void* h, float* out3) {
  auto s = static_cast<RefPM*>(h)->pm->resign_scores();
  std::memcpy(out3, s.data(), 12);
}
// out[0..6]: avg_game_length, avg_leaf_depth, avg_search_entropy, fast_avg_leaf_depth,
//            fast_avg_search_entropy, avg_moves_per_turn, avg_valid_moves
void azref_pm_metrics(void* h, float* out7) {
  auto& pm = *static_cast<RefPM*>(h)->pm;
  out7[0] = pm.avg_game_length();
  out7[1] = pm.avg_leaf_depth();
  out7[2] = pm.avg_search_entropy();
  out7[3] = pm.fast_avg_leaf_depth();
  out7[4] = pm.fast_avg_search_entropy();
  out7[5] = pm.avg_moves_per_turn();
  out7[6] = pm.avg_valid_moves();
}
// out[0..5]: hits, misses, evictions, reinserts, size, max_size
void azref_pm_cache_stats(void* h, uint64_t* out6) {
  auto& pm = *static_cast<RefPM*>(h)->pm;
  out6[0] = pm.cache_hits();
  out6[1] = pm.cache_misses();
  out6[2] = pm.cache_evictions();
  out6[3] = pm.cache_reinserts();
  out6[4] = pm.cache_size();
  out6[5] = pm.cache_max_size();
}
// bench.py only: simulations finished so far = visits * (moves of completed games + moves of the games in
// progress) + simulations of the searches in progress. Read while the workers run: only plain integers that
// live in place (GameData::move_count, MCTS::depth_) are touched, no pointers are followed, so the race is
// benign and the error is bounded by a few simulations per slot.
double azref_pm_progress_sims(void* h, uint32_t visits) {
  auto* r = static_cast<RefPM*>(h);
  const double done_moves = static_cast<double>(r->pm->avg_game_length()) * r->pm->games_completed();
  double moves = r->pm->games_completed() ? done_moves : 0.0;
  double partial = 0.0;
  for (uint32_t i = 0; i < r->cfg.concurrent_games; ++i) {
    const auto& g = r->pm->game_data(i);
    moves += *static_cast<const volatile uint32_t*>(&g.move_count);
    for (const auto& m : g.mcts) partial += m.depth();
  }
  return moves * visits + partial;
}
// Peeks into GameData (py_wrapper.cc:265-288 exposes gs/v/pi/canonical; the trees are C++-only).
// Only call while the worker is quiescent.
void azref_pm_game_state_bytes(void* h, uint32_t i, uint8_t* out89) {
  auto s = static_cast<RefPM*>(h)->pm->game_data(i).gs->to_bytes();
  std::memcpy(out89, s.data(), 89);
}
void azref_pm_game_counts(void* h, uint32_t i, uint32_t seat, uint32_t* out) {
  auto c = static_cast<RefPM*>(h)->pm->game_data(i).mcts[seat].counts();
  std::memcpy(out, c.data(), c.size() * 4);
}
void azref_pm_game_root_q(void* h, uint32_t i, uint32_t seat, float* out) {
  auto c = static_cast<RefPM*>(h)->pm->game_data(i).mcts[seat].root_q_values();
  std::memcpy(out, c.data(), c.size() * 4);
}
void azref_pm_game_root_value(void* h, uint32_t i, uint32_t seat, float* out3) {
  auto c = static_cast<RefPM*>(h)->pm->game_data(i).mcts[seat].root_value();
  std::memcpy(out3, c.data(), 12);
}
uint32_t azref_pm_game_depth(void* h, uint32_t i, uint32_t seat) {
  return static_cast<RefPM*>(h)->pm->game_data(i).mcts[seat].depth();
}
uint32_t azref_pm_game_root_n(void* h, uint32_t i, uint32_t seat) {
  return static_cast<RefPM*>(h)->pm->game_data(i).mcts[seat].root_n();
}

// ----------------------------------------------------------------------------- RNG ground truth
// A free-standing pcg32 driving the real libstdc++ algorithms, exactly as mcts.cc uses them
// (mcts.cc:19 `thread_local pcg32 re`, :100 shuffle, :205 extreme_value, :430-436 gamma, :718 uniform_real).
void* azref_rng_new(uint64_t seed, int use_stream, uint64_t stream) {
  return use_stream ? new pcg32(seed, stream) : new pcg32(seed);
}
void azref_rng_free(void* r) { delete static_cast<pcg32*>(r); }
uint32_t azref_rng_u32(void* r) { return (*static_cast<pcg32*>(r))(); }
void azref_rng_shuffle(void* r, uint32_t n, uint32_t* inout) {
  std::vector<uint32_t> v(inout, inout + n);
  std::shuffle(v.begin(), v.end(), *static_cast<pcg32*>(r));
  std::memcpy(inout, v.data(), n * 4);
}
float azref_rng_uniform01(void* r) {
  std::uniform_real_distribution<float> d{0.0F, 1.0F};
  return d(*static_cast<pcg32*>(r));
}
void azref_rng_gamma(void* r, float alpha, uint32_t n, float* out) {
  auto dist = std::gamma_distribution<float>{alpha, 1.0};
  for (uint32_t i = 0; i < n; ++i) out[i] = dist(*static_cast<pcg32*>(r));
}
float azref_rng_gumbel(void* r) {
  std::extreme_value_distribution<float> d{0.0f, 1.0f};
  return d(*static_cast<pcg32*>(r));
}

// ----------------------------------------------------------------------------- S3FIFOCache
void* azref_cache_new(uint32_t max_size, uint32_t ghost_size, uint32_t np, uint32_t nv) {
  return new S3FIFOCache(max_size, ghost_size, np, nv);
}
void azref_cache_free(void* c) { delete static_cast<S3FIFOCache*>(c); }
int azref_cache_find(void* c, uint64_t hash, float* pi, float* v) {
  return static_cast<S3FIFOCache*>(c)->find(hash, pi, v) ? 1 : 0;
}
void azref_cache_insert(void* c, uint64_t hash, const float* pi, const float* v) {
  static_cast<S3FIFOCache*>(c)->insert(hash, pi, v);
}
void azref_cache_stats(void* c, uint64_t* out6) {
  auto* k = static_cast<S3FIFOCache*>(c);
  out6[0] = k->hits();
  out6[1] = k->misses();
  out6[2] = k->evictions();
  out6[3] = k->reinserts();
  out6[4] = k->size();
  out6[5] = k->max_size();
}

}  // extern "C"
