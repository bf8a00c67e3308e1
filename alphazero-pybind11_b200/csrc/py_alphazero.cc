// py_alphazero.cc — the pybind11 module `alphazero`, backed by the B200 engine through the C ABI.
//
// Drop-in for the reference's src/py_wrapper.cc (module name, class names, method names, argument
// meaning and error behaviour) for the self-play hot path: PlayParams (py_wrapper.cc:295-349),
// PlayManager (352-504), GameData (265-288), GameState / Connect4GS (157-189, 560-586), PlayHistory
// (111-155), EvalType (290-293) and the Tracy no-op hooks (772-787). Everything that touches a game
// tree goes through include/b2az.h (libb2az.so, CUDA) — this file holds no search code.
//
// How the reference's thread pipeline maps onto a device engine (game_runner.py:648-745):
//   play()               the first caller becomes the DRIVER: it runs one generation at a time —
//                        b2az_step (process_result -> move -> find_leaf for every slot), then publishes the
//                        generation's leaf batch and waits until every leaf has been answered. Further callers
//                        (the reference starts `mcts_workers` threads) just wait for the run to end.
//   build_batch()        hands out rows of the published leaf batch, FIFO, with the reference's 500 us
//                        sub-timeouts / eager hand-off (py_wrapper.cc:449-504).
//   update_inferences()  stores v/pi by row; the call that answers the last leaf submits the whole
//                        generation to the device (b2az_submit_eval_host) and wakes the driver.
//   build_history_batch  b2az_drain_history into the caller's arrays.
// max_cache_size > 0 turns on the device position cache (the engine's replacement for ShardedS3FIFOCache).
// Gumbel root search, playout-cap randomisation and resign_percent are carried by the engine.
// Two model groups, seat permutations (slot g plays permutation g % n), per-seat visit budgets and a RANDOM group next
// to an NN one are carried. Not carried (rejected with RuntimeError instead of being ignored): per-seat overrides that
// differ between seats (epsilon, root temperature, Gumbel settings, resign thresholds), PLAYOUT eval. External caches
// size the device cache but are not shared between PlayManagers (see the constructor).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2az.h"
#include "az_connect4.h"
#include "az_tafl.h"
#include "az_stargambit.h"
#include "py_s3fifo.h"
#include <random>
#include <algorithm>

namespace py = pybind11;
using b2az::C4State;

namespace {

constexpr int kA = 7, kP = 2, kCanon = 168;

[[noreturn]] void throw_last(const char* what) {
  throw std::runtime_error(std::string(what) + ": " + b2az_last_error());
}

// ------------------------------------------------------------------------------------ PlayHistory
struct PlayHistory {  // game_state.h:14-18
  std::vector<float> canonical;  // [C][H][W]
  std::array<ssize_t, 3> dims{0, 0, 0};
  std::vector<float> v, pi;
};

// ------------------------------------------------------------------------------------ GameState
// Host-side view of ONE position (what Python tools hold); rules come from the same bitboard header the
// kernels use. Batched rule evaluation on the device is b2az_c4_batch.
class GameState {
 public:
  virtual ~GameState() = default;
  virtual std::unique_ptr<GameState> copy() const = 0;
  virtual bool equals(const GameState& o) const = 0;
  virtual std::string dump() const = 0;
  virtual uint32_t current_turn() const = 0;
  virtual uint8_t current_player() const = 0;
  virtual uint8_t num_players() const = 0;
  virtual uint32_t num_moves() const = 0;
  virtual uint8_t num_symmetries() const = 0;
  virtual bool relative_values() const { return false; }
  virtual std::vector<PlayHistory> symmetries(const PlayHistory& base) const = 0;
  virtual py::array_t<uint8_t> valid_moves() const = 0;
  virtual void play_move(uint32_t m) = 0;
  virtual py::object scores() const = 0;
  virtual void randomize_start() {}
  virtual int num_variants() const { return 0; }
  virtual int get_variant_id() const { return -1; }
  virtual py::array_t<float> canonicalized() const = 0;
  virtual std::string to_bytes() const { throw std::runtime_error("to_bytes() not implemented for this game type"); }
  virtual uint64_t hash() const = 0;
};

class Connect4GS : public GameState {  // connect4_gs.h:24-92
 public:
  C4State s;
  Connect4GS() { b2az::c4_init(s); }
  Connect4GS(const signed char* board84, int player, int turn) { b2az::c4_from_board(s, board84, player, turn); }
  std::unique_ptr<GameState> copy() const override { return std::make_unique<Connect4GS>(*this); }
  bool equals(const GameState& o) const override {  // connect4_gs.cc:23-31: board and player, not the turn
    auto* c = dynamic_cast<const Connect4GS*>(&o);
    return c && c->s.p[0] == s.p[0] && c->s.p[1] == s.p[1] && c->s.player == s.player;
  }
  std::string dump() const override {  // connect4_gs.cc:194-212
    std::string out = "Current Player: " + std::to_string((int)s.player) + '\n';
    for (int h = 0; h < 6; ++h) {
      for (int w = 0; w < 7; ++w) {
        if ((s.p[0] >> b2az::c4_bit(h, w)) & 1ULL) out += 'X';
        else if ((s.p[1] >> b2az::c4_bit(h, w)) & 1ULL) out += 'O';
        else out += '.';
      }
      out += '\n';
    }
    return out;
  }
  uint32_t current_turn() const override { return s.turn; }
  uint8_t current_player() const override { return s.player; }
  uint8_t num_players() const override { return kP; }
  uint32_t num_moves() const override { return kA; }
  uint8_t num_symmetries() const override { return 2; }
  std::vector<PlayHistory> symmetries(const PlayHistory& base) const override {  // connect4_gs.cc:151-170
    std::vector<PlayHistory> out{base};
    PlayHistory m;
    m.v = base.v;
    m.dims = base.dims;
    m.canonical.resize(base.canonical.size());
    const ssize_t C = base.dims[0], H = base.dims[1], W = base.dims[2];
    for (ssize_t f = 0; f < C; ++f)
      for (ssize_t h = 0; h < H; ++h)
        for (ssize_t w = 0; w < W; ++w) m.canonical[(f * H + h) * W + w] = base.canonical[(f * H + h) * W + (W - 1 - w)];
    m.pi.resize(base.pi.size());
    for (size_t w = 0; w < base.pi.size(); ++w) m.pi[w] = base.pi[base.pi.size() - 1 - w];
    out.push_back(std::move(m));
    return out;
  }
  py::array_t<uint8_t> valid_moves() const override {
    py::array_t<uint8_t> a(kA);
    const uint32_t vm = b2az::c4_valid_mask(s);
    for (int w = 0; w < kA; ++w) a.mutable_at(w) = (vm >> w) & 1u;
    return a;
  }
  void play_move(uint32_t m) override {  // connect4_gs.cc:48-58
    if (m >= (uint32_t)kA || !b2az::c4_play(s, m)) throw std::runtime_error("Invalid move: You have a bug in your code.");
  }
  py::object scores() const override {
    const uint32_t t = b2az::c4_terminal(s);
    if (!t) return py::none();
    py::array_t<float> a(kP + 1);
    for (int i = 0; i < kP + 1; ++i) a.mutable_at(i) = (t == (uint32_t)i + 1u) ? 1.0f : 0.0f;
    return std::move(a);
  }
  py::array_t<float> canonicalized() const override {
    py::array_t<float> a({4, 6, 7});
    float* d = a.mutable_data();
    for (uint32_t e = 0; e < (uint32_t)kCanon; ++e) d[e] = b2az::c4_canon_elem(s.p[0], s.p[1], s.player, e);
    return a;
  }
  std::string to_bytes() const override {  // connect4_gs.cc:172-178
    std::string out(89, '\0');
    b2az::c4_to_board(s, reinterpret_cast<signed char*>(&out[0]));
    out[84] = (char)s.player;
    const int32_t turn = (int32_t)s.turn;
    std::memcpy(&out[85], &turn, 4);
    return out;
  }
  static Connect4GS from_bytes(const std::string& data) {
    if (data.size() != 89) throw std::runtime_error("Connect4GS::from_bytes: wrong byte length");
    int32_t turn = 0;
    std::memcpy(&turn, &data[85], 4);
    return Connect4GS(reinterpret_cast<const signed char*>(&data[0]), (signed char)data[84], turn);
  }
  uint64_t hash() const override { return b2az::c4_hash(s); }  // equality class of connect4_gs.cc:33-37
};

#include "py_tafl_gs.h"  // BrandubhGS / OpenTaflGS / TawlbwrddGS
#include "py_stargambit_gs.h"  // StarGambit{Skirmish,Showdown,Clash,Battle}GS, StarGambitUnifiedGS (+ pinned subclasses)
#include "py_mcts.h"           // MCTS: one tree of the device search behind the reference's single-tree API

// ------------------------------------------------------------------------------------ PlayParams
enum class EvalType : uint8_t { NN = 0, RANDOM = 1, PLAYOUT = 2 };

struct PlayParams {  // play_manager.h:60-154, same defaults
  uint32_t games_to_play = 0;
  uint32_t concurrent_games = 0;
  uint32_t max_batch_size = 1;
  uint32_t max_cache_size = 0;
  uint8_t cache_shards = 1;
  uint8_t queue_shards = 1;
  uint8_t eval_pipelines = 1;
  std::vector<uint32_t> mcts_visits{};
  float cpuct = 2.0f;
  float start_temp = 1.0f;
  float final_temp = 1.0f;
  float temp_decay_half_life = 0.0f;
  std::vector<float> temp_decay_half_life_by_variant{};
  bool history_enabled = false;
  bool self_play = false;
  bool tree_reuse = true;
  float epsilon = 0.0f;
  float mcts_root_temp = 1.0f;
  bool playout_cap_randomization = false;
  uint32_t playout_cap_depth = 25;
  float playout_cap_percent = 0.75f;
  float fpu_reduction = 0.0f;
  bool root_fpu_zero = false;
  bool shaped_dirichlet = false;
  bool policy_target_pruning = false;
  bool gumbel_enabled = false;
  uint32_t gumbel_m = 16;
  float gumbel_c_visit = 50.0f;
  float gumbel_c_scale = 1.0f;
  bool gumbel_full = false;
  bool fast_search_uses_gumbel = false;
  float resign_percent = 0.0f;
  float resign_playthrough_percent = 0.0f;
  std::vector<EvalType> eval_type{};
  std::vector<uint8_t> model_groups{};
  std::vector<std::vector<uint8_t>> seat_perms{};
  std::vector<std::vector<uint32_t>> seat_visits{};
  std::vector<std::vector<uint32_t>> seat_cap_visits{};
  std::vector<std::vector<float>> seat_epsilon{};
  std::vector<std::vector<float>> seat_mcts_root_temp{};
  std::vector<std::vector<uint8_t>> seat_root_fpu_zero{};
  std::vector<std::vector<uint8_t>> seat_gumbel_enabled{};
  std::vector<std::vector<uint32_t>> seat_gumbel_m{};
  std::vector<std::vector<float>> seat_gumbel_c_visit{};
  std::vector<std::vector<float>> seat_gumbel_c_scale{};
  std::vector<std::vector<uint8_t>> seat_gumbel_full{};
  std::vector<std::vector<uint8_t>> seat_gumbel_use_improved_policy{};
  std::vector<std::vector<float>> seat_resign_threshold{};
  std::vector<std::vector<uint32_t>> seat_resign_consecutive{};
  // additive (no reference counterpart): engine placement / determinism
  int device = 0;
  uint64_t seed = 0;
  bool deterministic = false;  // one pcg32 stream in slot order (the reference's single-thread order)
  uint64_t pool_nodes = 0;
};

// PlayManager's constructor normalisation (play_manager.cc:24-176): model groups, seat permutations and the per-seat
// 2-D overrides, with the reference's dimension errors. What the engines carry today: ONE model group, ONE seat
// permutation, per-seat visit budgets (seat_visits / seat_cap_visits: self_play's asymmetric search budget,
// game_runner.py:2027-2034); the other per-seat tables must be uniform (every seat the same value), which then
// replaces the global. Anything else is rejected with "not implemented", never ignored.
struct SeatTables {
  std::vector<uint8_t> model_groups;
  uint32_t num_model_groups = 1;
  std::vector<std::vector<uint8_t>> seat_perms;
  std::vector<std::vector<uint32_t>> visits, cap_visits;
  std::vector<std::vector<float>> epsilon, root_temp;
  std::vector<std::vector<uint8_t>> root_fpu_zero, gumbel_enabled, gumbel_full, gumbel_use_improved;
  std::vector<std::vector<uint32_t>> gumbel_m, resign_consecutive;
  std::vector<std::vector<float>> gumbel_c_visit, gumbel_c_scale, resign_threshold;
};
template <typename T>
bool uniform2d(const std::vector<std::vector<T>>& t) {
  for (auto& row : t)
    for (auto& x : row)
      if (!(x == t[0][0])) return false;
  return true;
}
inline SeatTables normalize_seats(const PlayParams& P, size_t np) {
  SeatTables N;
  if (P.mcts_visits.size() != np) throw std::runtime_error{"You must specify MCTS visits for each player"};
  if (P.model_groups.empty()) for (size_t i = 0; i < np; ++i) N.model_groups.push_back((uint8_t)i);
  else N.model_groups = P.model_groups;
  if (N.model_groups.size() < np) throw std::runtime_error{"model_groups must name a group for each player"};
  N.num_model_groups = *std::max_element(N.model_groups.begin(), N.model_groups.end()) + 1u;
  std::vector<uint32_t> group_visits(N.num_model_groups, 0);
  for (size_t i = 0; i < np; ++i) group_visits[N.model_groups[i]] = P.mcts_visits[i];
  if (P.seat_perms.empty()) N.seat_perms.push_back(N.model_groups);
  else N.seat_perms = P.seat_perms;
  const size_t num_perms = N.seat_perms.size();
  auto validate = [&](const auto& vec, const char* name) {
    if (vec.size() != num_perms) throw std::runtime_error{std::string(name) + " outer dimension must match number of seat permutations"};
    for (size_t p = 0; p < num_perms; ++p)
      if (vec[p].size() != np) throw std::runtime_error{std::string(name) + " inner dimension must match number of players"};
  };
  for (auto& perm : N.seat_perms) {
    if (perm.size() != np) throw std::runtime_error{"seat_perms inner dimension must match number of players"};
    for (auto g : perm)
      if (g >= N.num_model_groups) throw std::runtime_error{"seat_perms names a model group that does not exist"};
  }
  auto fill = [&](auto& dst, const auto& src, const char* name, auto def) {
    if (src.empty()) dst.assign(num_perms, std::decay_t<decltype(dst[0])>(np, def));
    else { validate(src, name); dst = src; }
  };
  if (P.seat_visits.empty()) {
    N.visits.resize(num_perms);
    for (size_t p = 0; p < num_perms; ++p)
      for (size_t s = 0; s < np; ++s) N.visits[p].push_back(group_visits[N.seat_perms[p][s]]);
  } else {
    validate(P.seat_visits, "seat_visits");
    N.visits = P.seat_visits;
  }
  fill(N.cap_visits, P.seat_cap_visits, "seat_cap_visits", (uint32_t)P.playout_cap_depth);
  fill(N.epsilon, P.seat_epsilon, "seat_epsilon", P.epsilon);
  fill(N.root_temp, P.seat_mcts_root_temp, "seat_mcts_root_temp", P.mcts_root_temp);
  fill(N.root_fpu_zero, P.seat_root_fpu_zero, "seat_root_fpu_zero", (uint8_t)(P.root_fpu_zero ? 1 : 0));
  fill(N.gumbel_enabled, P.seat_gumbel_enabled, "seat_gumbel_enabled", (uint8_t)(P.gumbel_enabled ? 1 : 0));
  fill(N.gumbel_m, P.seat_gumbel_m, "seat_gumbel_m", (uint32_t)P.gumbel_m);
  fill(N.gumbel_c_visit, P.seat_gumbel_c_visit, "seat_gumbel_c_visit", P.gumbel_c_visit);
  fill(N.gumbel_c_scale, P.seat_gumbel_c_scale, "seat_gumbel_c_scale", P.gumbel_c_scale);
  fill(N.gumbel_full, P.seat_gumbel_full, "seat_gumbel_full", (uint8_t)(P.gumbel_full ? 1 : 0));
  fill(N.gumbel_use_improved, P.seat_gumbel_use_improved_policy, "seat_gumbel_use_improved_policy", (uint8_t)0);
  fill(N.resign_threshold, P.seat_resign_threshold, "seat_resign_threshold", -2.0f);
  fill(N.resign_consecutive, P.seat_resign_consecutive, "seat_resign_consecutive", (uint32_t)1);
  return N;
}
// What the engines do not carry yet (see SeatTables): rejected loudly. On success the uniform per-seat tables have been
// folded into `P`'s globals.
inline void fold_supported_seats(PlayParams& P, const SeatTables& N, const char* engine, uint32_t max_groups, bool per_seat_search) {
  auto reject = [&](bool bad, const char* what) {
    if (bad) throw std::runtime_error(std::string(what) + " is not implemented by the " + engine + " yet");
  };
  reject(N.num_model_groups > max_groups, "this many model groups (model_groups / seat_perms with different networks)");
  reject(N.seat_perms.size() > 8, "more than eight seat permutations");
  if (!per_seat_search) {  // (the wide-tree engine carries a search-settings record per (permutation, seat))
    reject(!uniform2d(N.epsilon) || !uniform2d(N.root_temp) || !uniform2d(N.root_fpu_zero), "different seat_epsilon / seat_mcts_root_temp / seat_root_fpu_zero per seat");
    reject(!uniform2d(N.gumbel_enabled) || !uniform2d(N.gumbel_m) || !uniform2d(N.gumbel_c_visit) || !uniform2d(N.gumbel_c_scale) ||
               !uniform2d(N.gumbel_full), "different Gumbel settings per seat");
    reject(!uniform2d(N.resign_threshold) || N.resign_threshold[0][0] > -1.5f, "seat_resign_threshold");
  }
  reject(N.gumbel_use_improved[0][0] != 0 || !uniform2d(N.gumbel_use_improved), "seat_gumbel_use_improved_policy");
  P.epsilon = N.epsilon[0][0];
  P.mcts_root_temp = N.root_temp[0][0];
  P.root_fpu_zero = N.root_fpu_zero[0][0] != 0;
  P.gumbel_enabled = N.gumbel_enabled[0][0] != 0;
  P.gumbel_m = N.gumbel_m[0][0];
  P.gumbel_c_visit = N.gumbel_c_visit[0][0];
  P.gumbel_c_scale = N.gumbel_c_scale[0][0];
  P.gumbel_full = N.gumbel_full[0][0] != 0;
}

// eval_types_[group] (play_manager.cc:577-587): all RANDOM -> the fused on-device evaluator; all NN -> the leaf batch; a
// RANDOM group next to an NN one (game_runner.play_past against iteration 0: RandPlayer) -> group_random of the engine.
struct EvalPlan {
  bool all_random = false;
  uint8_t group_random[2] = {0, 0};
};
inline EvalPlan plan_eval(const PlayParams& P, const SeatTables& N, const char* who) {
  EvalPlan e;
  if (P.eval_type.empty()) return e;
  uint32_t n_random = 0, n_used = 0;
  for (uint32_t g = 0; g < N.num_model_groups; ++g) {
    bool used = false;
    for (auto& perm : N.seat_perms)
      for (auto x : perm) used |= x == g;
    if (!used) continue;
    ++n_used;
    const EvalType et = g < P.eval_type.size() ? P.eval_type[g] : EvalType::NN;
    if (et == EvalType::PLAYOUT) throw std::runtime_error(std::string("EvalType.PLAYOUT is not implemented by the ") + who + " yet");
    if (et == EvalType::RANDOM) { ++n_random; if (g < 2) e.group_random[g] = 1; }
  }
  if (n_random == n_used) { e.all_random = true; e.group_random[0] = e.group_random[1] = 0; }
  return e;
}
// PlayParams::seat_perms and the per-permutation budgets into a b2az_params / b2az_tafl_selfplay_params
template <class CP>
inline void fill_perms(CP& cp, const SeatTables& N, const EvalPlan& ev) {
  cp.n_seat_perms = (uint32_t)N.seat_perms.size();
  for (size_t i = 0; i < N.seat_perms.size(); ++i)
    for (int s = 0; s < 2; ++s) {
      cp.seat_perms[i][s] = N.seat_perms[i][s];
      cp.perm_seat_visits[i][s] = N.visits[i][s];
      cp.perm_seat_cap_visits[i][s] = N.cap_visits[i][s];
    }
  cp.group_random[0] = ev.group_random[0];
  cp.group_random[1] = ev.group_random[1];
}

// ------------------------------------------------------------------------------------ DLPack (dlpack.h v0.8 ABI, restated)
// The zero-copy evaluator feed (SURVEY.md 8b "additive exports"; north_star: "fed zero-copy ... via DLPack"): the leaf
// batch stays in the engine's device buffers and is handed to torch as DLPack capsules; the evaluations come back the
// same way (any object with __dlpack__, e.g. a torch CUDA tensor, or a raw capsule).
namespace dl {
struct DLDevice { int32_t device_type; int32_t device_id; };  // kDLCUDA = 2
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };  // kDLUInt = 1, kDLFloat = 2
struct DLTensor {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
};
struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(DLManagedTensor*);
};
struct Owned {  // the engine owns the memory: the capsule only carries the shape
  DLManagedTensor m;
  int64_t shape[4];
};
inline py::capsule make(void* data, int device, uint8_t code, uint8_t bits, std::initializer_list<int64_t> shape) {
  auto* o = new Owned();
  int nd = 0;
  for (int64_t d : shape) o->shape[nd++] = d;
  o->m.dl_tensor = DLTensor{data, DLDevice{2, device}, nd, DLDataType{code, bits, 1}, o->shape, nullptr, 0};
  o->m.manager_ctx = o;
  o->m.deleter = [](DLManagedTensor* m) { delete static_cast<Owned*>(m->manager_ctx); };
  return py::capsule(&o->m, "dltensor", [](PyObject* cap) {
    if (PyCapsule_IsValid(cap, "dltensor")) {  // never consumed: free the descriptor ("used_dltensor" belongs to the consumer)
      auto* m = static_cast<DLManagedTensor*>(PyCapsule_GetPointer(cap, "dltensor"));
      if (m && m->deleter) m->deleter(m);
    }
  });
}
// a consumed view of an incoming tensor: keeps the producer's object alive until released
struct In {
  py::object keep;  // the capsule (renamed "used_dltensor": this side calls the deleter)
  DLManagedTensor* m = nullptr;
  ~In() { release(); }
  void release() {
    if (m && m->deleter) m->deleter(m);
    m = nullptr;
    keep = py::object();
  }
};
inline void take(py::object obj, In& in, int device, int64_t rows, int64_t cols, const char* what) {
  py::object cap = obj;
  if (!PyCapsule_CheckExact(obj.ptr())) {
    if (!py::hasattr(obj, "__dlpack__")) throw std::runtime_error(std::string(what) + ": expected a DLPack capsule or an object with __dlpack__");
    cap = obj.attr("__dlpack__")();
  }
  if (!PyCapsule_IsValid(cap.ptr(), "dltensor")) throw std::runtime_error(std::string(what) + ": not an unconsumed dltensor capsule");
  auto* m = static_cast<DLManagedTensor*>(PyCapsule_GetPointer(cap.ptr(), "dltensor"));
  PyCapsule_SetName(cap.ptr(), "used_dltensor");
  PyCapsule_SetDestructor(cap.ptr(), nullptr);
  in.release();
  in.keep = cap;
  in.m = m;
  const DLTensor& t = m->dl_tensor;
  bool ok = t.device.device_type == 2 && t.device.device_id == device && t.dtype.code == 2 && t.dtype.bits == 32 && t.dtype.lanes == 1 &&
            t.ndim == 2 && t.shape[0] >= rows && t.shape[1] == cols;
  if (ok && t.strides) ok = t.strides[1] == 1 && (t.strides[0] == cols || t.shape[0] <= 1);
  if (!ok) {
    in.release();
    throw std::runtime_error(std::string(what) + ": expected a contiguous float32 CUDA tensor [rows >= " + std::to_string(rows) + ", " +
                             std::to_string(cols) + "] on device " + std::to_string(device));
  }
}
inline const float* ptr(const In& in) { return reinterpret_cast<const float*>(static_cast<const char*>(in.m->dl_tensor.data) + in.m->dl_tensor.byte_offset); }
}  // namespace dl

// ------------------------------------------------------------------------------------ PlayManager
class PlayManager;
struct GameData {  // play_manager.h:33-58, the part Python sees (py_wrapper.cc:265-288)
  PlayManager* pm;
  uint32_t index;
};

class PlayManager {
 public:
  // PlayManager(BrandubhGS | OpenTaflGS | TawlbwrddGS | StarGambit*GS, params): the wide-tree self-play engine
  // (b2az_tafl_selfplay_*). `rows` = staging rows per game slot (the game's max_turns for the tafl games, a bound on the
  // samples of one Star Gambit game), `k_typ` = a typical branching factor for sizing the node slabs.
  void setup_wide(uint32_t game, uint32_t rows, int planes, int side, int actions, uint32_t k_typ, bool relative_values,
                  const char* who) {
    tables_ = normalize_seats(params_, kP);  // play_manager.cc:19-176, with its errors
    PlayParams eff = params_;
    fold_supported_seats(eff, tables_, who, 2, true);
    const PlayParams& P = eff;
    auto reject = [who](bool bad, const char* what) {
      if (bad) throw std::runtime_error(std::string(what) + " is not implemented by the " + who + " yet");
    };
    reject(!P.temp_decay_half_life_by_variant.empty() && game < 20, "temp_decay_half_life_by_variant");
    reject(P.concurrent_games == 0 || P.games_to_play % P.concurrent_games != 0, "games_to_play not a multiple of concurrent_games");
    const EvalPlan ev = plan_eval(P, tables_, who);
    random_eval_ = ev.all_random;
    b2az_tafl_selfplay_params sp{};
    fill_perms(sp, tables_, ev);
    // make_mcts(perm, seat) (play_manager.cc:602-617) + the per-seat resign rule (:335-366): only when a seat differs from
    // the others (uniform tables have been folded into the globals, which the search reads from the constant bank)
    bool any_resign = false;
    for (auto& row : tables_.resign_threshold) for (float x : row) any_resign |= x > -2.0f;
    // (a uniform seat_gumbel_enabled that differs from the global flag also needs the table: the global picks the policy target)
    sp.has_seat_search = (any_resign || P.gumbel_enabled != params_.gumbel_enabled || !uniform2d(tables_.epsilon) || !uniform2d(tables_.root_temp) || !uniform2d(tables_.root_fpu_zero) ||
                          !uniform2d(tables_.gumbel_enabled) || !uniform2d(tables_.gumbel_m) || !uniform2d(tables_.gumbel_c_visit) ||
                          !uniform2d(tables_.gumbel_c_scale) || !uniform2d(tables_.gumbel_full)) ? 1 : 0;
    for (size_t i = 0; i < tables_.seat_perms.size(); ++i)
      for (int s = 0; s < 2; ++s) {
        sp.seat_epsilon[i][s] = tables_.epsilon[i][s];
        sp.seat_root_temp[i][s] = tables_.root_temp[i][s];
        sp.seat_root_fpu_zero[i][s] = tables_.root_fpu_zero[i][s];
        sp.seat_gumbel_enabled[i][s] = tables_.gumbel_enabled[i][s];
        sp.seat_gumbel_full[i][s] = tables_.gumbel_full[i][s];
        sp.seat_gumbel_m[i][s] = tables_.gumbel_m[i][s];
        sp.seat_gumbel_c_visit[i][s] = tables_.gumbel_c_visit[i][s];
        sp.seat_gumbel_c_scale[i][s] = tables_.gumbel_c_scale[i][s];
        sp.seat_resign_threshold[i][s] = tables_.resign_threshold[i][s];
        sp.seat_resign_consecutive[i][s] = tables_.resign_consecutive[i][s];
      }
    sp.forest.game = game;
    sp.forest.max_turns = rows;
    sp.forest.relative_values = relative_values;
    // slab per tree: each half holds the kept subtree + one move's new nodes (1 + 8k words per expanded node)
    sp.forest.words_per_tree = P.pool_nodes ? (uint32_t)P.pool_nodes
                                            : 2u * (1u + 4u * max_seat_visits() * (1u + 8u * k_typ));
    sp.forest.cpuct = P.cpuct; sp.forest.fpu_reduction = P.fpu_reduction; sp.forest.epsilon = P.epsilon;
    sp.forest.root_policy_temp = P.mcts_root_temp; sp.forest.root_fpu_zero = P.root_fpu_zero;
    sp.forest.gumbel_enabled = params_.gumbel_enabled;  // the GLOBAL flag: it alone picks the policy target (play_manager.cc:412-419)
    sp.forest.gumbel_m = P.gumbel_m; sp.forest.seed = P.seed;
    sp.forest.gumbel_full = P.gumbel_full;
    sp.forest.gumbel_c_visit = P.gumbel_c_visit; sp.forest.gumbel_c_scale = P.gumbel_c_scale;
    sp.forest.shaped_dirichlet = P.shaped_dirichlet;
    sp.n_games = P.concurrent_games;
    sp.games_per_slot = P.games_to_play / P.concurrent_games;
    sp.visits = max_seat_visits();
    sp.seat_visits[0] = tables_.visits[0][0]; sp.seat_visits[1] = tables_.visits[0][1];
    sp.seat_cap_visits[0] = tables_.cap_visits[0][0]; sp.seat_cap_visits[1] = tables_.cap_visits[0][1];
    sp.playout_cap_randomization = P.playout_cap_randomization; sp.playout_cap_depth = P.playout_cap_depth;
    sp.playout_cap_percent = P.playout_cap_percent; sp.fast_search_uses_gumbel = P.fast_search_uses_gumbel;
    sp.resign_percent = P.resign_percent; sp.resign_playthrough_percent = P.resign_playthrough_percent;
    sp.start_temp = P.start_temp; sp.final_temp = P.final_temp; sp.temp_decay_half_life = P.temp_decay_half_life;
    sp.history_enabled = P.history_enabled; sp.policy_target_pruning = P.policy_target_pruning; sp.tree_reuse = P.tree_reuse;
    sp.n_variant_half_life = (uint32_t)std::min<size_t>(4, P.temp_decay_half_life_by_variant.size());
    for (uint32_t i = 0; i < sp.n_variant_half_life; ++i) sp.variant_half_life[i] = P.temp_decay_half_life_by_variant[i];
    for (int i = 0; i < 4; ++i) sp.variant_probs[i] = sg_probs_[i];
    sp.cache_entries = P.max_cache_size;  // one model group: the whole budget (play_manager.cc:195-203)
    // history_ is unbounded in the reference; here the sample ring holds what a run can produce between drains: every
    // sample of the run when that fits an 8 GB budget (play() first, build_history_batch afterwards works), else the
    // budget — a full ring drops samples and play() then fails loudly (B2AZ_DEVERR_HIST)
    const uint64_t row_bytes = 4ull * ((uint64_t)planes * side * side + actions + 4);
    const uint64_t want = (uint64_t)P.games_to_play * rows, floor_rows = (uint64_t)P.concurrent_games * rows;
    sp.hist_capacity = (uint32_t)std::min<uint64_t>(0x7FFFFFFFull, std::max<uint64_t>(floor_rows, std::min<uint64_t>(want, (8ull << 30) / row_bytes)));
    if (b2az_tafl_selfplay_create(&sp, P.device, &tsp_) != 0) throw_last("PlayManager");
    canon_sz_ = (uint32_t)(planes * side * side);
    A_ = (uint32_t)actions;
    cdims_[0] = planes; cdims_[1] = cdims_[2] = side;
  }
  template <int GAME>
  bool try_tafl(const GameState* gs) {
    auto* t = dynamic_cast<const TaflGS<GAME>*>(gs);
    if (!t) return false;
    if (t->s.turn != 0 || t->hist_len != 0) throw std::runtime_error("the B200 engine starts every game from the initial position");
    setup_wide(GAME, t->s.max_turns, TaflGS<GAME>::P, TaflGS<GAME>::S, TaflGS<GAME>::A, GAME == B2AZ_TAFL_BRANDUBH ? 64u : 200u,
               false, "B200 tafl engine");
    return true;
  }
  // PlayManager(StarGambit{Skirmish,Showdown,Clash,Battle}GS | StarGambitUnifiedGS pinned to a variant, params)
  bool try_star_gambit(const GameState* gs) {
    auto* t = dynamic_cast<const StarGambitBase*>(gs);
    if (!t) return false;
    if (t->s.turn != 1 || t->s.n_units != 2) throw std::runtime_error("the B200 engine starts every game from the initial position");
    const bool mix = t->unified && (t->pinned < 0 || t->pinned > 3);  // every new game draws its variant (randomize_start)
    has_variants_ = t->unified;
    if (mix) for (int i = 0; i < 4; ++i) sg_probs_[i] = t->probs[i];
    const auto sp = t->space();
    // a Star Gambit game has no small bound on its actions (200 turns of several actions each): 512 staged samples per
    // game slot; a longer game keeps its most recent 512
    setup_wide(mix ? 24u : (t->unified ? 20u : 10u) + t->s.variant, 512u, sp.planes(t->unified), sp.udim, sp.num_moves(), 64u, true,
               "B200 Star Gambit engine");
    return true;
  }
  uint32_t max_seat_visits() const {
    uint32_t m = 1;
    for (auto& row : tables_.visits) for (auto x : row) m = std::max(m, x);
    return m;
  }
  void size_buffers() {
    G_ = params_.concurrent_games;
    canon_.resize((size_t)G_ * canon_sz_);
    ids_.resize(G_);
    v_.assign((size_t)G_ * (kP + 1), 0.0f);
    pi_.assign((size_t)G_ * A_, 0.0f);
    row_of_game_.assign(G_, 0xFFFFFFFFu);
    seats_.assign(G_, 0);
    group_rows_.assign(tables_.num_model_groups, {});
    group_next_.assign(tables_.num_model_groups, 0);
    refresh_stats_locked();
  }
  PlayManager(const GameState* gs, PlayParams p) : params_(std::move(p)), base_gs_(gs->copy()) {
    if (try_tafl<B2AZ_TAFL_BRANDUBH>(gs) || try_tafl<B2AZ_TAFL_OPENTAFL>(gs) || try_tafl<B2AZ_TAFL_TAWLBWRDD>(gs) ||
        try_star_gambit(gs)) {
      size_buffers();
      return;
    }
    auto* c4 = dynamic_cast<const Connect4GS*>(gs);
    if (!c4) throw std::runtime_error("the B200 engine implements Connect4GS, the tafl games and Star Gambit only");
    if (c4->s.p[0] || c4->s.p[1] || c4->s.player || c4->s.turn)
      throw std::runtime_error("the B200 engine starts every game from the initial Connect4 position");
    tables_ = normalize_seats(params_, kP);  // play_manager.cc:19-176, with its errors
    PlayParams eff = params_;
    fold_supported_seats(eff, tables_, "B200 engine", 2, false);
    const PlayParams& P = eff;
    auto reject = [](bool bad, const char* what) {
      if (bad) throw std::runtime_error(std::string(what) + " is not implemented by the B200 engine yet");
    };
    reject(!P.temp_decay_half_life_by_variant.empty(), "temp_decay_half_life_by_variant");
    const EvalPlan ev = plan_eval(P, tables_, "B200 engine");
    random_eval_ = ev.all_random;
    b2az_params bp;
    b2az_params_default(&bp);
    fill_perms(bp, tables_, ev);
    bp.games_to_play = P.games_to_play;
    bp.concurrent_games = P.concurrent_games;
    bp.max_batch_size = P.max_batch_size;
    bp.max_cache_size = P.max_cache_size;  // one model group: the whole budget (play_manager.cc:195-203)
    bp.mcts_visits[0] = tables_.visits[0][0];  // seat_visits / seat_cap_visits: the per-seat budgets (play_manager.cc:70-90)
    bp.mcts_visits[1] = tables_.visits[0][1];
    bp.seat_cap_visits[0] = tables_.cap_visits[0][0];
    bp.seat_cap_visits[1] = tables_.cap_visits[0][1];
    bp.model_groups[0] = tables_.seat_perms[0][0];  // the network that searches for seat s (play_manager.cc:577)
    bp.model_groups[1] = tables_.seat_perms[0][1];
    bp.cpuct = P.cpuct;
    bp.start_temp = P.start_temp;
    bp.final_temp = P.final_temp;
    bp.temp_decay_half_life = P.temp_decay_half_life;
    bp.history_enabled = P.history_enabled;
    bp.self_play = P.self_play;
    bp.tree_reuse = P.tree_reuse;
    bp.epsilon = P.epsilon;
    bp.mcts_root_temp = P.mcts_root_temp;
    bp.gumbel_enabled = P.gumbel_enabled;
    bp.gumbel_m = P.gumbel_m;
    bp.gumbel_c_visit = P.gumbel_c_visit;
    bp.gumbel_c_scale = P.gumbel_c_scale;
    bp.gumbel_full = P.gumbel_full;
    bp.fast_search_uses_gumbel = P.fast_search_uses_gumbel;
    bp.playout_cap_randomization = P.playout_cap_randomization;
    bp.resign_percent = P.resign_percent;
    bp.resign_playthrough_percent = P.resign_playthrough_percent;
    bp.playout_cap_depth = P.playout_cap_depth;
    bp.playout_cap_percent = P.playout_cap_percent;
    bp.fpu_reduction = P.fpu_reduction;
    bp.root_fpu_zero = P.root_fpu_zero;
    bp.shaped_dirichlet = P.shaped_dirichlet;
    bp.policy_target_pruning = P.policy_target_pruning;
    bp.eval_type = random_eval_ ? B2AZ_EVAL_RANDOM : B2AZ_EVAL_NN;
    bp.rng_mode = P.deterministic ? B2AZ_RNG_GLOBAL : B2AZ_RNG_PER_GAME;
    bp.seed = P.seed;
    bp.pool_nodes = P.pool_nodes;
    if (b2az_create(&bp, P.device, &eng_) != 0) throw_last("PlayManager");
    size_buffers();
  }
  ~PlayManager() {
    stop();
    if (eng_) b2az_destroy(eng_);
    if (tsp_) b2az_tafl_selfplay_destroy(tsp_);
  }
  PlayManager(const PlayManager&) = delete;

  const PlayParams& params() const { return params_; }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);  // under the waiters' mutex: a wait that has just tested its predicate cannot miss this
      stopped_.store(true);
    }
    cv_.notify_all();
  }
  bool stopped() const { return stopped_.load(); }
  uint32_t games_completed() const { return games_completed_.load(); }
  uint32_t remaining_games() const {  // play_manager.h:177-180
    if (stopped_.load()) return 0;
    const uint32_t done = games_completed_.load();
    return done >= params_.games_to_play ? 0 : params_.games_to_play - done;
  }

  // PlayManager::play (play_manager.cc:258-600)
  void play() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (driver_active_ || finished_) {  // the reference's extra worker threads: nothing left for them to do
        cv_.wait(lk, [&] { return finished_ || stopped_.load(); });
        return;
      }
      driver_active_ = true;
    }
    try {
      drive();
    } catch (...) {
      std::lock_guard<std::mutex> lk(mu_);
      driver_active_ = false;
      finished_ = true;
      stopped_.store(true);
      cv_.notify_all();
      throw;
    }
    std::lock_guard<std::mutex> lk(mu_);
    driver_active_ = false;
    finished_ = true;
    cv_.notify_all();
  }

  // build_batch (py_wrapper.cc:449-504)
  std::vector<uint32_t> build_batch(uint32_t group, float* batch, ssize_t ndim, const ssize_t* shape, uint32_t /*shard*/) {
    if (group >= tables_.num_model_groups) throw std::runtime_error("model group out of range");
    std::vector<uint32_t> out;
    const uint32_t mbs = params_.max_batch_size;
    out.reserve(mbs);
    auto max_bs = [&]() -> uint32_t {
      const uint32_t rem = remaining_games();
      return std::min(mbs, rem > 0 ? rem : 1u);
    };
    uint32_t empty = 0;
    bool checked = false;
    std::unique_lock<std::mutex> lk(mu_);
    while (out.size() < max_bs()) {
      if (eager_.load() && !out.empty()) break;
      if (remaining_games() == 0) break;
      const std::vector<uint32_t>& rows = group_rows_[group];  // this group's rows of the published batch, FIFO
      const uint32_t avail = (uint32_t)rows.size() - group_next_[group];
      if (avail == 0) {
        if (++empty >= 20u) break;  // MAX_EMPTY x SUB_TIMEOUT = ~10 ms
        cv_.wait_for(lk, std::chrono::microseconds(500));
        continue;
      }
      empty = 0;
      if (!checked) {
        if (ndim != 4 || shape[1] != cdims_[0] || shape[2] != cdims_[1] || shape[3] != cdims_[2]) throw std::runtime_error("Improper batch size");
        checked = true;
      }
      const uint32_t cap = (uint32_t)std::min<ssize_t>(shape[0], (ssize_t)max_bs());
      if (out.size() >= cap) break;
      const uint32_t n = std::min<uint32_t>(avail, cap - (uint32_t)out.size());
      for (uint32_t i = 0; i < n; ++i) {
        const uint32_t r = rows[group_next_[group] + i];
        std::memcpy(batch + (out.size() + i) * canon_sz_, canon_.data() + (size_t)r * canon_sz_, (size_t)canon_sz_ * sizeof(float));
      }
      for (uint32_t i = 0; i < n; ++i) out.push_back(ids_[rows[group_next_[group] + i]]);
      group_next_[group] += n;
    }
    return out;
  }

  // PlayManager::update_inferences (play_manager.cc:619-642)
  void update_inferences(uint32_t group, const std::vector<uint32_t>& idx, const float* v, ssize_t vrows, ssize_t vcols,
                         const float* pi, ssize_t prows, ssize_t pcols) {
    if (group >= tables_.num_model_groups) throw std::runtime_error("model group out of range");
    if (vcols != kP + 1 || pcols != (ssize_t)A_ || vrows < (ssize_t)idx.size() || prows < (ssize_t)idx.size())
      throw std::runtime_error("Eigen is angry!!!");  // shapes.h:4-6: the reference asserts on bad shapes
    std::unique_lock<std::mutex> lk(mu_);
    for (size_t i = 0; i < idx.size(); ++i) {
      if (idx[i] >= G_ || row_of_game_[idx[i]] == 0xFFFFFFFFu) throw std::runtime_error("update_inferences: game has no pending leaf");
      const uint32_t r = row_of_game_[idx[i]];
      std::memcpy(&v_[(size_t)r * (kP + 1)], v + i * (kP + 1), (kP + 1) * sizeof(float));
      std::memcpy(&pi_[(size_t)r * A_], pi + i * A_, A_ * sizeof(float));
      row_of_game_[idx[i]] = 0xFFFFFFFFu;
    }
    answered_ += (uint32_t)idx.size();
    if (answered_ == leaf_count_ && leaf_count_ > 0) cv_.notify_all();
  }

  // pop_game / pop_games_upto / push_inference: the per-game flavour of the same hand-off
  py::object pop_game(uint32_t group) {
    auto v = pop_games_upto(group, 1);
    if (v.empty()) return py::none();
    return py::int_(v[0]);
  }
  std::vector<uint32_t> pop_games_upto(uint32_t group, size_t n) {
    if (group >= tables_.num_model_groups) throw std::runtime_error("model group out of range");
    std::unique_lock<std::mutex> lk(mu_);
    if (group_rows_[group].size() == group_next_[group]) cv_.wait_for(lk, std::chrono::milliseconds(10));  // MAX_WAIT (play_manager.h:26)
    const uint32_t take = (uint32_t)std::min<size_t>(n, group_rows_[group].size() - group_next_[group]);
    std::vector<uint32_t> out;
    for (uint32_t i = 0; i < take; ++i) out.push_back(ids_[group_rows_[group][group_next_[group] + i]]);
    group_next_[group] += take;
    return out;
  }
  void push_inference(uint32_t i) {  // the caller has written game_data(i).v() / .pi() in place
    std::unique_lock<std::mutex> lk(mu_);
    if (i >= G_ || row_of_game_[i] == 0xFFFFFFFFu) throw std::runtime_error("push_inference: game has no pending leaf");
    row_of_game_[i] = 0xFFFFFFFFu;
    if (++answered_ == leaf_count_) cv_.notify_all();
  }

  // build_history_batch (py_wrapper.cc:393-424)
  uint32_t build_history_batch(float* canon, ssize_t n, float* v, float* pi) {
    uint32_t cur = 0;
    while (cur < (uint32_t)n && (remaining_games() > 0 || hist_count() > 0)) {
      uint32_t got = 0;
      {
        std::lock_guard<std::mutex> lk(api_);
        const int rc = tsp_ ? b2az_tafl_selfplay_drain_history(tsp_, nullptr, (uint32_t)n - cur, canon + (size_t)cur * canon_sz_,
                                                               v + (size_t)cur * (kP + 1), pi + (size_t)cur * A_, nullptr, &got)
                            : b2az_drain_history(eng_, nullptr, (uint32_t)n - cur, canon + (size_t)cur * canon_sz_,
                                                 v + (size_t)cur * (kP + 1), pi + (size_t)cur * A_, 0, &got);
        if (rc != 0) throw_last("build_history_batch");
      }
      cur += got;
      if (got == 0) {
        if (finished_.load()) {
          refresh_stats();
          if (hist_count() == 0) break;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
      }
    }
    refresh_stats();
    return cur;
  }

  // ---- the zero-copy evaluator feed (additive: SURVEY.md 8b). A synchronous driver of its own — the caller alternates
  // leaf_batch_dlpack / update_inferences_dlpack until remaining_games() == 0; play() must not be running.
  // leaf_batch_dlpack(group) -> (canonical capsule float32[B,C,H,W] cuda, valid-move capsule uint8[B,A] cuda | None, B).
  // Connect4: the B leaves of this generation, compacted (cache hits never become rows). Wide-tree games: B =
  // concurrent_games, row g = slot g (retired slots keep stale rows; their answers are ignored), no mask.
  py::tuple leaf_batch_dlpack(uint32_t group) {
    if (group != 0 || tables_.num_model_groups != 1) throw std::runtime_error("leaf_batch_dlpack: one model group only");
    if (random_eval_) throw std::runtime_error("leaf_batch_dlpack: needs EvalType.NN");
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (driver_active_) throw std::runtime_error("leaf_batch_dlpack: play() is driving this PlayManager");
    }
    std::lock_guard<std::mutex> lk(api_);
    dl_v_.release(); dl_pi_.release();  // the previous answers have been consumed by the step that ran since
    if (tsp_) {
      const float* canon = nullptr;
      if (b2az_tafl_selfplay_find_leaf(tsp_, nullptr, &canon) != 0) throw_last("leaf_batch_dlpack");
      dl_rows_ = G_;
      return py::make_tuple(dl::make(const_cast<float*>(canon), params_.device, 2, 32, {(int64_t)G_, cdims_[0], cdims_[1], cdims_[2]}),
                            py::none(), G_);
    }
    uint32_t n = 0;
    const float* canon = nullptr;
    const uint32_t* ids = nullptr;
    for (int tries = 0; tries < 1 << 20; ++tries) {  // a generation whose leaves all hit the cache has no rows: step on
      if (!dl_stepped_ && b2az_step(eng_, 1, nullptr) != 0) throw_last("leaf_batch_dlpack");
      dl_stepped_ = false;
      if (b2az_leaf_batch(eng_, nullptr, &n, &canon, &ids) != 0) throw_last("leaf_batch_dlpack");
      if (n > 0) break;
      b2az_stats st;
      if (b2az_get_stats(eng_, nullptr, &st) != 0) throw_last("leaf_batch_dlpack");
      if (st.active_games == 0) break;
    }
    dl_rows_ = n;
    refresh_stats_unlocked_api();
    if (n == 0) return py::make_tuple(py::none(), py::none(), 0u);
    const uint8_t* valid = nullptr;
    if (b2az_leaf_valid_device(eng_, nullptr, &valid) != 0) throw_last("leaf_batch_dlpack");
    return py::make_tuple(dl::make(const_cast<float*>(canon), params_.device, 2, 32, {(int64_t)n, cdims_[0], cdims_[1], cdims_[2]}),
                          dl::make(const_cast<uint8_t*>(valid), params_.device, 1, 8, {(int64_t)n, (int64_t)A_}), n);
  }
  // update_inferences_dlpack(group, v float32[B,P+1], pi float32[B,A]): CUDA tensors (anything with __dlpack__) in the row
  // order of the batch; the engine reads them in place on its next step, which is enqueued here.
  void update_inferences_dlpack(uint32_t group, py::object v, py::object pi) {
    if (group != 0) throw std::runtime_error("update_inferences_dlpack: one model group only");
    std::lock_guard<std::mutex> lk(api_);
    if (dl_rows_ == 0) throw std::runtime_error("update_inferences_dlpack: no leaf batch is waiting (call leaf_batch_dlpack)");
    dl::take(v, dl_v_, params_.device, dl_rows_, kP + 1, "update_inferences_dlpack(v)");
    dl::take(pi, dl_pi_, params_.device, dl_rows_, A_, "update_inferences_dlpack(pi)");
    if (tsp_) {
      uint32_t active = 0;
      if (b2az_tafl_selfplay_process_result(tsp_, nullptr, dl::ptr(dl_v_), dl::ptr(dl_pi_), 0, &active) != 0) throw_last("update_inferences_dlpack");
    } else {
      if (b2az_submit_eval(eng_, dl::ptr(dl_v_), dl::ptr(dl_pi_), dl_rows_) != 0) throw_last("update_inferences_dlpack");
      if (b2az_step(eng_, 1, nullptr) != 0) throw_last("update_inferences_dlpack");  // consumes the answers (stream ordered)
      dl_stepped_ = true;
    }
    dl_rows_ = 0;
    refresh_stats_unlocked_api();
  }
  b2az_perm_stats perm(size_t i) {
    if (i >= tables_.seat_perms.size()) throw std::out_of_range("seat permutation out of range");
    b2az_perm_stats out[8];
    std::lock_guard<std::mutex> lk(api_);
    if ((tsp_ ? b2az_tafl_selfplay_perm_stats(tsp_, nullptr, out, nullptr) : b2az_perm_scores(eng_, nullptr, out, nullptr)) != 0) throw_last("perm stats");
    return out[i];
  }
  int num_tracked_variants() const { return has_variants_ ? 4 : 0; }
  b2az_variant_stats variant(int v) {
    if (!has_variants_) throw std::out_of_range("this game has no variants");
    if (v < 0 || v >= 4) throw std::out_of_range("variant out of range");
    b2az_variant_stats out[4];
    std::lock_guard<std::mutex> lk(api_);
    if (b2az_tafl_selfplay_variant_stats(tsp_, nullptr, out) != 0) throw_last("variant stats");
    return out[v];
  }
  void refresh_stats_unlocked_api() {  // (api_ is held by the caller; mu_ is never taken under api_: lock order mu_ -> api_)
    b2az_stats st;
    const int rc = tsp_ ? b2az_tafl_selfplay_get_stats(tsp_, nullptr, &st) : b2az_get_stats(eng_, nullptr, &st);
    if (rc != 0) throw_last("stats");
    dl_stats_ = st;
    games_completed_.store((uint32_t)st.games_completed);
    hist_count_.store((uint32_t)st.hist_count);
    if (st.active_games == 0) finished_.store(true);
  }
  b2az_stats stats() {
    refresh_stats();
    std::lock_guard<std::mutex> lk(mu_);
    return stats_;
  }
  uint32_t hist_count() const { return hist_count_.load(); }
  size_t awaiting_inference_count() {
    std::lock_guard<std::mutex> lk(mu_);
    size_t n = 0;
    for (size_t g = 0; g < group_rows_.size(); ++g) n += group_rows_[g].size() - group_next_[g];
    return n;
  }
  size_t awaiting_mcts_count() {
    std::lock_guard<std::mutex> lk(mu_);
    return leaf_count_ ? answered_ : 0;
  }
  void set_eager(bool e) { eager_.store(e); }
  uint32_t num_model_groups() const { return tables_.num_model_groups; }
  size_t num_seat_perms() const { return tables_.seat_perms.size(); }
  uint32_t concurrent() const { return G_; }

  // GameData accessors
  // GameData::gs (play_manager.h:37): the slot's current position as a host GameState of the game's own class
  std::unique_ptr<GameState> game_state(uint32_t i) {
    std::lock_guard<std::mutex> lk(api_);
    if (tsp_) {
      std::unique_ptr<GameState> g = base_gs_->copy();
      uint32_t n = 0;
      if (auto* sg = dynamic_cast<StarGambitBase*>(g.get())) {
        sg->hist.assign(4096, 0);
        if (b2az_tafl_selfplay_root_state(tsp_, i, &sg->s, (uint32_t)sizeof(sg->s), sg->hist.data(), (uint32_t)sg->hist.size(), &n) != 0) throw_last("game_data");
        sg->hist.resize(n);
        return g;
      }
      auto fill = [&](auto* t) {
        if (!t) return false;
        t->hist.assign((size_t)t->s.max_turns + 66, b2az::TaflKey{});
        if (b2az_tafl_selfplay_root_state(tsp_, i, &t->s, (uint32_t)sizeof(t->s), t->hist.data(), (uint32_t)t->hist.size(), &n) != 0) throw_last("game_data");
        t->hist_len = n;
        return true;
      };
      if (fill(dynamic_cast<TaflGS<B2AZ_TAFL_BRANDUBH>*>(g.get())) || fill(dynamic_cast<TaflGS<B2AZ_TAFL_OPENTAFL>*>(g.get())) ||
          fill(dynamic_cast<TaflGS<B2AZ_TAFL_TAWLBWRDD>*>(g.get())))
        return g;
      throw std::runtime_error("game_data(i).gs(): unknown game class");
    }
    uint8_t st[89];
    if (b2az_peek(eng_, nullptr, i, 0, st, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) != 0) throw_last("game_data");
    return std::make_unique<Connect4GS>(Connect4GS::from_bytes(std::string(reinterpret_cast<char*>(st), 89)));
  }
  // the slot's pending leaf row (canonical planes + the v / pi the evaluator writes), or -1
  ssize_t row(uint32_t i) {
    std::lock_guard<std::mutex> lk(mu_);
    if (i >= G_) throw std::runtime_error("game index out of range");
    for (uint32_t r = 0; r < leaf_count_; ++r)
      if (ids_[r] == i) return r;
    return -1;
  }
  float* canon_row(size_t r) { return canon_.data() + r * canon_sz_; }
  float* v_row(size_t r) { return v_.data() + r * (kP + 1); }
  float* pi_row(size_t r) { return pi_.data() + r * A_; }
  ssize_t num_actions() const { return (ssize_t)A_; }
  const ssize_t* canon_dims() const { return cdims_; }

 private:
  void refresh_stats() {
    std::lock_guard<std::mutex> lk(mu_);
    refresh_stats_locked();
  }
  void refresh_stats_locked() {
    std::lock_guard<std::mutex> lk(api_);
    if ((tsp_ ? b2az_tafl_selfplay_get_stats(tsp_, nullptr, &stats_) : b2az_get_stats(eng_, nullptr, &stats_)) != 0) throw_last("stats");
    games_completed_.store(stats_.games_completed);
    hist_count_.store(stats_.hist_count);
  }
  // hand the generation's n rows to the batcher threads: row r goes to the queue of the model group that searches for
  // its seat (play_manager.cc:577, 598)
  void publish_locked(uint32_t n) {
    for (auto& g : group_rows_) g.clear();
    std::fill(group_next_.begin(), group_next_.end(), 0u);
    for (uint32_t r = 0; r < n; ++r) {
      row_of_game_[ids_[r]] = r;
      group_rows_[tables_.num_model_groups > 1 ? std::min<uint32_t>(seats_[r], tables_.num_model_groups - 1u) : 0].push_back(r);
    }
    answered_ = 0;
    leaf_count_ = n;
  }
  void retract_locked() {
    leaf_count_ = 0;
    for (auto& g : group_rows_) g.clear();
    std::fill(group_next_.begin(), group_next_.end(), 0u);
  }
  void drive_tafl() {
    for (;;) {
      if (stopped_.load()) return;
      uint32_t n = 0;
      if (random_eval_) {  // EvalType::RANDOM: searches and moves stay on the device, eight moves per call
        {
          std::lock_guard<std::mutex> lk(api_);
          if (b2az_tafl_selfplay_play(tsp_, nullptr, 8, &n) != 0) throw_last("play");
        }
        refresh_stats();
        std::lock_guard<std::mutex> lk(mu_);
        if (stats_.device_error) throw std::runtime_error("play: device error (training-sample ring exhausted)");
        if (n == 0) return;
        continue;
      }
      {
        std::lock_guard<std::mutex> lk(api_);
        if (b2az_tafl_selfplay_leaf_batch_host(tsp_, nullptr, G_, canon_.data(), ids_.data(), &n) != 0) throw_last("play");
      }
      std::unique_lock<std::mutex> lk(mu_);
      refresh_stats_locked();
      if (stats_.device_error) throw std::runtime_error("play: device error (tree slab / training-sample ring exhausted)");
      if (n == 0 && stats_.active_games == 0) return;  // every slot retired
      if (n > 0) {
        publish_locked(n);
        cv_.notify_all();
        cv_.wait(lk, [&] { return answered_ == leaf_count_ || stopped_.load(); });
        if (stopped_.load()) return;
        retract_locked();
      }  // (n == 0 with slots still cycling: every leaf of this round hit the position cache — only the moves are due)
      {
        std::lock_guard<std::mutex> lk2(api_);
        if (b2az_tafl_selfplay_submit_eval_host(tsp_, nullptr, ids_.data(), v_.data(), pi_.data(), n) != 0) throw_last("update_inferences");
      }
    }
  }
  void drive() {
    if (tsp_) return drive_tafl();
    if (random_eval_) {
      // EvalType::RANDOM: the evaluator runs inside the step kernel; fuse a search's worth of generations
      const uint32_t chunk = std::max<uint32_t>(1, std::min<uint32_t>(params_.mcts_visits[0], 512));
      for (;;) {
        if (stopped_.load()) return;
        {
          std::lock_guard<std::mutex> lk(api_);
          if (b2az_step(eng_, chunk, nullptr) != 0) throw_last("play");
        }
        refresh_stats();
        std::lock_guard<std::mutex> lk(mu_);
        if (stats_.device_error) throw std::runtime_error("play: device error (pool / history ring exhausted)");
        if (stats_.active_games == 0) return;
      }
    }
    for (;;) {
      if (stopped_.load()) return;
      uint32_t n = 0;
      {
        std::lock_guard<std::mutex> lk(api_);
        if (b2az_step(eng_, 1, nullptr) != 0) throw_last("play");
        if (b2az_leaf_batch_host(eng_, nullptr, G_, canon_.data(), ids_.data(), &n) != 0) throw_last("play");
        if (tables_.num_model_groups > 1 && n > 0 && b2az_leaf_groups_host(eng_, nullptr, seats_.data(), n) != 0) throw_last("play");
      }
      std::unique_lock<std::mutex> lk(mu_);
      refresh_stats_locked();
      if (n == 0) {
        if (stats_.active_games == 0) return;  // every slot retired
        continue;  // only searches of an EvalType::RANDOM group are under way: nothing for the evaluators this generation
      }
      publish_locked(n);
      cv_.notify_all();
      cv_.wait(lk, [&] { return answered_ == leaf_count_ || stopped_.load(); });
      if (stopped_.load()) return;
      retract_locked();
      {
        std::lock_guard<std::mutex> lk2(api_);
        if (b2az_submit_eval_host(eng_, nullptr, ids_.data(), v_.data(), pi_.data(), n) != 0) throw_last("update_inferences");
      }
    }
  }

  PlayParams params_;
  std::unique_ptr<GameState> base_gs_;  // base_gs_ of the reference: the class (and variant mix) every game starts from
  SeatTables tables_;
  b2az_engine* eng_ = nullptr;
  b2az_tafl_selfplay* tsp_ = nullptr;  // set instead of eng_ for a tafl GameState
  uint32_t canon_sz_ = kCanon, A_ = kA;
  ssize_t cdims_[3] = {4, 6, 7};
  uint32_t G_ = 0;
  bool random_eval_ = false;
  std::mutex mu_;   // generation state below
  std::mutex api_;  // serialises calls into the C ABI (one engine, one stream)
  std::condition_variable cv_;
  bool driver_active_ = false;
  std::atomic<bool> finished_{false};
  std::atomic<bool> stopped_{false};
  std::atomic<bool> eager_{false};
  std::atomic<uint32_t> games_completed_{0}, hist_count_{0};
  std::vector<float> canon_, v_, pi_;
  std::vector<uint32_t> ids_, row_of_game_;
  std::vector<uint8_t> seats_;                    // model group of every row (several model groups only)
  std::vector<std::vector<uint32_t>> group_rows_; // rows of the published batch per model group, FIFO
  std::vector<uint32_t> group_next_;              // per group: rows already handed out
  uint32_t leaf_count_ = 0, answered_ = 0;
  b2az_stats stats_{};
  float sg_probs_[4] = {0.25f, 0.25f, 0.25f, 0.25f};  // StarGambitUnifiedGS variant weights (variant mix)
  bool has_variants_ = false;                          // num_variants() > 0: per-variant score tables
  dl::In dl_v_, dl_pi_;      // the evaluations of the DLPack feed, kept alive until the engine has read them
  uint32_t dl_rows_ = 0;
  bool dl_stepped_ = false;  // update_inferences_dlpack has already run the next generation's step
  b2az_stats dl_stats_{};
};

py::array_t<float> vec3(const float* p) {
  py::array_t<float> a(3);
  for (int i = 0; i < 3; ++i) a.mutable_at(i) = p[i];
  return a;
}

}  // namespace

// NOLINTNEXTLINE
PYBIND11_MODULE(alphazero, m) {
  m.doc() = "the c++ parts of an alphazero implementation (B200 engine behind the reference's API)";

  py::class_<PlayHistory>(m, "PlayHistory")
      .def(py::init([](py::array_t<float, py::array::c_style | py::array::forcecast> canonical,
                       py::array_t<float, py::array::c_style | py::array::forcecast> v,
                       py::array_t<float, py::array::c_style | py::array::forcecast> pi) {
             if (canonical.ndim() != 3 || v.ndim() != 1 || pi.ndim() != 1) throw std::runtime_error("PlayHistory: bad shapes");
             PlayHistory ph;
             ph.dims = {canonical.shape(0), canonical.shape(1), canonical.shape(2)};
             ph.canonical.assign(canonical.data(), canonical.data() + canonical.size());
             ph.v.assign(v.data(), v.data() + v.size());
             ph.pi.assign(pi.data(), pi.data() + pi.size());
             return ph;
           }),
           py::arg().none(false), py::arg().none(false), py::arg().none(false))
      .def("v", [](PlayHistory& ph) { return py::array_t<float>({(ssize_t)ph.v.size()}, ph.v.data(), py::cast(&ph)); })
      .def("pi", [](PlayHistory& ph) { return py::array_t<float>({(ssize_t)ph.pi.size()}, ph.pi.data(), py::cast(&ph)); })
      .def("canonical", [](PlayHistory& ph) {
        const ssize_t sz = sizeof(float);
        return py::memoryview::from_buffer(ph.canonical.data(), {ph.dims[0], ph.dims[1], ph.dims[2]},
                                           {sz * ph.dims[1] * ph.dims[2], sz * ph.dims[2], sz});
      }, py::keep_alive<0, 1>());

  py::class_<GameState>(m, "GameState")
      .def("copy", &GameState::copy)
      .def("__eq__", [](const GameState& a, const GameState& b) { return a.equals(b); })
      .def("__str__", &GameState::dump)
      .def("current_turn", &GameState::current_turn)
      .def("current_player", &GameState::current_player)
      .def("num_players", &GameState::num_players)
      .def("num_moves", &GameState::num_moves)
      .def("num_symmetries", &GameState::num_symmetries)
      .def("relative_values", &GameState::relative_values)
      .def("symmetries", &GameState::symmetries)
      .def("valid_moves", &GameState::valid_moves)
      .def("play_move", &GameState::play_move)
      .def("scores", &GameState::scores)
      .def("randomize_start", &GameState::randomize_start)
      .def("num_variants", &GameState::num_variants)
      .def("get_variant_id", &GameState::get_variant_id)
      .def("canonicalized", &GameState::canonicalized);

  m.def("hash_game_state", [](const GameState& gs) { return gs.hash(); }, py::arg("gs"));

  py::class_<Connect4GS, GameState>(m, "Connect4GS")
      .def(py::init<>())
      .def(py::init([](const py::array_t<int8_t, py::array::c_style | py::array::forcecast>& board, int8_t player, int32_t turn) {
        if (board.ndim() != 3 || board.shape(0) != 2 || board.shape(1) != 6 || board.shape(2) != 7)
          throw std::runtime_error{"Improper connect 4 board shape"};
        return Connect4GS(reinterpret_cast<const signed char*>(board.data()), player, turn);
      }))
      .def_static("NUM_PLAYERS", [] { return 2; })
      .def_static("NUM_MOVES", [] { return 7; })
      .def_static("NUM_SYMMETRIES", [] { return 2; })
      .def_static("CANONICAL_SHAPE", [] { return std::array<int64_t, 3>{4, 6, 7}; })
      .def(py::pickle([](const Connect4GS& gs) { return py::bytes(gs.to_bytes()); },
                      [](py::bytes b) { return Connect4GS::from_bytes(std::string(b)); }));

  bind_tafl_gs<BrandubhGS>(m, "BrandubhGS");    // py_wrapper.cc:527-536
  bind_tafl_gs<OpenTaflGS>(m, "OpenTaflGS");    // py_wrapper.cc:538-547
  bind_tafl_gs<TawlbwrddGS>(m, "TawlbwrddGS");  // py_wrapper.cc:549-558
  bind_star_gambit(m);                         // py_wrapper.cc:589-695
  bind_mcts(m);                                // py_wrapper.cc:191-220

  py::enum_<EvalType>(m, "EvalType").value("NN", EvalType::NN).value("RANDOM", EvalType::RANDOM).value("PLAYOUT", EvalType::PLAYOUT);

  py::class_<PlayParams>(m, "PlayParams")
      .def(py::init<>())
#define RW(f) .def_readwrite(#f, &PlayParams::f)
      RW(games_to_play) RW(concurrent_games) RW(max_batch_size) RW(max_cache_size) RW(cache_shards) RW(queue_shards)
      RW(eval_pipelines) RW(mcts_visits) RW(cpuct) RW(playout_cap_randomization) RW(playout_cap_depth)
      RW(playout_cap_percent) RW(start_temp) RW(final_temp) RW(temp_decay_half_life) RW(temp_decay_half_life_by_variant)
      RW(history_enabled) RW(tree_reuse) RW(self_play) RW(epsilon) RW(fpu_reduction) RW(root_fpu_zero) RW(shaped_dirichlet)
      RW(policy_target_pruning) RW(gumbel_enabled) RW(gumbel_m) RW(gumbel_c_visit) RW(gumbel_c_scale) RW(gumbel_full)
      RW(fast_search_uses_gumbel) RW(mcts_root_temp) RW(resign_percent) RW(resign_playthrough_percent) RW(eval_type)
      RW(model_groups) RW(seat_perms) RW(seat_visits) RW(seat_cap_visits) RW(seat_epsilon) RW(seat_mcts_root_temp)
      RW(seat_root_fpu_zero) RW(seat_gumbel_enabled) RW(seat_gumbel_m) RW(seat_gumbel_c_visit) RW(seat_gumbel_c_scale)
      RW(seat_gumbel_full) RW(seat_gumbel_use_improved_policy) RW(seat_resign_threshold) RW(seat_resign_consecutive)
      RW(device) RW(seed) RW(deterministic) RW(pool_nodes)
#undef RW
      ;

  py::class_<GameData>(m, "GameData")
      .def("gs", [](const GameData& gd) { return gd.pm->game_state(gd.index); })
      .def("valid_moves", [](const GameData& gd) { return gd.pm->game_state(gd.index)->valid_moves(); })
      .def("v", [](const GameData& gd) {
        const ssize_t r = gd.pm->row(gd.index);
        if (r < 0) throw std::runtime_error("GameData.v(): the game has no pending leaf");
        return py::array_t<float>({(ssize_t)3}, gd.pm->v_row((size_t)r), py::cast(gd.pm));
      })
      .def("pi", [](const GameData& gd) {
        const ssize_t r = gd.pm->row(gd.index);
        if (r < 0) throw std::runtime_error("GameData.pi(): the game has no pending leaf");
        return py::array_t<float>({gd.pm->num_actions()}, gd.pm->pi_row((size_t)r), py::cast(gd.pm));
      })
      .def("canonical", [](const GameData& gd) {
        const ssize_t r = gd.pm->row(gd.index);
        if (r < 0) throw std::runtime_error("GameData.canonical(): the game has no pending leaf");
        const ssize_t sz = sizeof(float);
        const ssize_t* d = gd.pm->canon_dims();
        return py::memoryview::from_buffer(gd.pm->canon_row((size_t)r), {d[0], d[1], d[2]}, {sz * d[1] * d[2], sz * d[2], sz});
      });

  py::class_<PlayManager>(m, "PlayManager")
      .def(py::init([](const GameState* gs, PlayParams params) { return std::make_unique<PlayManager>(gs, std::move(params)); }),
           py::arg().none(false), py::arg())
      // PlayManager(gs, params, caches) (play_manager.cc:644-649; tournament.py:224, visit_sweep_elo.py:324): the reference
      // shares the given host caches between successive PlayManagers. The engine's position cache lives in HBM inside the
      // engine, so the given objects only SIZE it (the sum of their max_size(), as if params.max_cache_size had been set);
      // its hits / misses are reported by this PlayManager's cache_hits() / cache_misses(), the host objects stay empty and
      // nothing carries over to the next PlayManager. A cache may forget, so the games are the same — only colder.
      .def(py::init([](const GameState* gs, PlayParams params, std::vector<py::object> caches) {
             uint64_t total = 0;
             for (auto& c : caches) {
               if (c.is_none()) continue;
               if (!py::hasattr(c, "max_size")) throw py::type_error("caches: expected S3FIFOCache / ShardedS3FIFOCache objects or None");
               total += c.attr("max_size")().cast<uint64_t>();
             }
             params.max_cache_size = (uint32_t)std::min<uint64_t>(total, 0x7FFFFFFFull);
             return std::make_unique<PlayManager>(gs, std::move(params));
           }),
           py::arg().none(false), py::arg(), py::arg("caches"))
      .def("game_data", [](PlayManager& pm, uint32_t i) {
        if (i >= pm.concurrent()) throw std::runtime_error("game index out of range");
        return GameData{&pm, i};
      }, py::keep_alive<0, 1>())
      .def("params", &PlayManager::params, py::return_value_policy::reference_internal)
      .def("scores", [](PlayManager& pm) { auto s = pm.stats(); return vec3(s.scores); })
      .def("resign_scores", [](PlayManager& pm) { auto s = pm.stats(); return vec3(s.resign_scores); })
      .def("games_completed", &PlayManager::games_completed)
      .def("remaining_games", &PlayManager::remaining_games)
      .def("stop", &PlayManager::stop)
      .def("stopped", &PlayManager::stopped)
      .def("awaiting_inference_count", &PlayManager::awaiting_inference_count)
      .def("awaiting_mcts_count", &PlayManager::awaiting_mcts_count)
      .def("hist_count", [](PlayManager& pm) { return pm.stats().hist_count; })
      .def("cache_hits", [](PlayManager& pm) { return pm.stats().cache_hits; })
      .def("cache_misses", [](PlayManager& pm) { return pm.stats().cache_misses; })
      .def("cache_evictions", [](PlayManager& pm) { return pm.stats().cache_evictions; })
      .def("cache_reinserts", [](PlayManager& pm) { return pm.stats().cache_reinserts; })
      .def("cache_max_size", [](PlayManager& pm) { return pm.stats().cache_max_size; })
      .def("cache_size", [](PlayManager& pm) { return pm.stats().cache_size; })
      .def("avg_game_length", [](PlayManager& pm) { return pm.stats().avg_game_length; })
      .def("avg_leaf_depth", [](PlayManager& pm) { return pm.stats().avg_leaf_depth; })
      .def("avg_search_entropy", [](PlayManager& pm) { return pm.stats().avg_search_entropy; })
      .def("fast_avg_leaf_depth", [](PlayManager& pm) { return pm.stats().fast_avg_leaf_depth; })
      .def("fast_avg_search_entropy", [](PlayManager& pm) { return pm.stats().fast_avg_search_entropy; })
      .def("avg_moves_per_turn", [](PlayManager& pm) { return pm.stats().avg_moves_per_turn; })
      .def("avg_valid_moves", [](PlayManager& pm) { return pm.stats().avg_valid_moves; })
      .def("simulations", [](PlayManager& pm) { return pm.stats().simulations; })  // additive
      .def("play", &PlayManager::play, py::call_guard<py::gil_scoped_release>())
      .def("pop_game", [](PlayManager& pm, uint32_t g) {
        std::vector<uint32_t> v;
        { py::gil_scoped_release rel; v = pm.pop_games_upto(g, 1); }
        return v.empty() ? py::object(py::none()) : py::object(py::int_(v[0]));
      })
      .def("pop_games_upto", &PlayManager::pop_games_upto, py::call_guard<py::gil_scoped_release>())
      .def("push_inference", &PlayManager::push_inference, py::call_guard<py::gil_scoped_release>())
      .def("update_inferences",
           [](PlayManager& pm, uint8_t group, const std::vector<uint32_t>& idx,
              py::array_t<float, py::array::c_style | py::array::forcecast> v,
              py::array_t<float, py::array::c_style | py::array::forcecast> pi) {
             if (v.ndim() != 2 || pi.ndim() != 2) throw std::runtime_error("Eigen is angry!!!");
             const float *vp = v.data(), *pp = pi.data();
             const ssize_t vr = v.shape(0), vc = v.shape(1), pr = pi.shape(0), pc = pi.shape(1);
             py::gil_scoped_release rel;
             pm.update_inferences(group, idx, vp, vr, vc, pp, pr, pc);
           })
      .def("build_history_batch",
           [](PlayManager& pm, py::array_t<float, py::array::c_style>& canonical, py::array_t<float, py::array::c_style>& v,
              py::array_t<float, py::array::c_style>& pi) {
             const ssize_t* d = pm.canon_dims();
             if (canonical.ndim() != 4 || v.ndim() != 2 || pi.ndim() != 2 || canonical.shape(1) != d[0] || canonical.shape(2) != d[1] ||
                 canonical.shape(3) != d[2] || v.shape(1) != 3 || pi.shape(1) != pm.num_actions())
               throw std::runtime_error("Improper history batch shape");
             const ssize_t n = std::min(canonical.shape(0), std::min(v.shape(0), pi.shape(0)));
             float *c = canonical.mutable_data(), *vv = v.mutable_data(), *pp = pi.mutable_data();
             py::gil_scoped_release rel;
             return pm.build_history_batch(c, n, vv, pp);
           })
      .def("num_model_groups", &PlayManager::num_model_groups)
      .def("num_seat_perms", &PlayManager::num_seat_perms)
      .def("perm_scores", [](PlayManager& pm, size_t i) { return vec3(pm.perm(i).scores); })
      .def("perm_games_completed", [](PlayManager& pm, size_t i) { return pm.perm(i).games_completed; })
      // per-variant tracking (play_manager.h:218-275): games with variants (StarGambitUnifiedGS) only
      .def("num_tracked_variants", &PlayManager::num_tracked_variants)
#define VAR_F(name, expr) .def(name, [](PlayManager& pm, int v) -> float { const b2az_variant_stats m = pm.variant(v); return expr; })
      VAR_F("variant_avg_game_length", m.games_completed ? (float)m.game_length / (float)m.games_completed : 0.0f)
      VAR_F("variant_avg_leaf_depth", m.full_move_count ? (float)(m.leaf_depth / (double)m.full_move_count) : 0.0f)
      VAR_F("variant_avg_search_entropy", m.full_move_count ? (float)(m.entropy / (double)m.full_move_count) : 0.0f)
      VAR_F("variant_fast_avg_leaf_depth", m.fast_move_count ? (float)(m.fast_leaf_depth / (double)m.fast_move_count) : 0.0f)
      VAR_F("variant_fast_avg_search_entropy", m.fast_move_count ? (float)(m.fast_entropy / (double)m.fast_move_count) : 0.0f)
      VAR_F("variant_avg_moves_per_turn", m.game_length ? (float)m.total_move_count / (float)m.game_length : 0.0f)
      VAR_F("variant_avg_valid_moves", m.total_move_count ? (float)(m.valid_moves / (double)m.total_move_count) : 0.0f)
#undef VAR_F
      .def("variant_games_completed", [](PlayManager& pm, int v) { return pm.variant(v).games_completed; })
      .def("variant_scores", [](PlayManager& pm, int v) { return vec3(pm.variant(v).scores); })
      .def("variant_perm_scores", [](PlayManager& pm, int v, int p) {
        if (v < 0 || v >= pm.num_tracked_variants()) throw std::out_of_range("variant out of range");
        return vec3(pm.perm((size_t)p).variant_scores[v]);
      })
      .def("variant_perm_games_completed", [](PlayManager& pm, int v, int p) {
        if (v < 0 || v >= pm.num_tracked_variants()) throw std::out_of_range("variant out of range");
        return pm.perm((size_t)p).variant_games_completed[v];
      })
      .def("set_eager", &PlayManager::set_eager)
      .def("leaf_batch_dlpack", &PlayManager::leaf_batch_dlpack, py::arg("group") = 0)
      .def("update_inferences_dlpack", &PlayManager::update_inferences_dlpack, py::arg("group"), py::arg("v"), py::arg("pi"))
      .def("build_batch",
           [](PlayManager& pm, uint32_t group, py::array_t<float, py::array::c_style>& batch, uint32_t shard) {
             float* data = batch.mutable_data();
             const ssize_t nd = batch.ndim();
             ssize_t shape[4] = {0, 0, 0, 0};
             for (ssize_t i = 0; i < std::min<ssize_t>(nd, 4); ++i) shape[i] = batch.shape(i);
             py::gil_scoped_release rel;
             return pm.build_batch(group, data, nd, shape, shard);
           },
           py::arg("group"), py::arg("batch"), py::arg("shard") = 0);

  // S3FIFOCache / ShardedS3FIFOCache (py_wrapper.cc:222-259): the host containers Python tools hold (cache_utils.py)
  using b2az_host::S3FIFOCache;
  using b2az_host::ShardedS3FIFOCache;
  py::class_<S3FIFOCache>(m, "S3FIFOCache")
      .def(py::init<uint32_t, uint32_t, uint32_t, uint32_t>(), py::arg("max_size"), py::arg("ghost_size"), py::arg("num_policy"),
           py::arg("num_value"))
      .def("find",
           [](S3FIFOCache& c, uint64_t hash, uint32_t num_policy, uint32_t num_value) -> py::object {
             if (num_policy != c.num_policy() || num_value != c.num_value()) throw std::runtime_error("S3FIFOCache.find: wrong num_policy / num_value");
             py::array_t<float> policy(num_policy), value(num_value);
             if (!c.find(hash, policy.mutable_data(), value.mutable_data())) return py::none();
             return py::make_tuple(policy, value);
           },
           py::arg("hash"), py::arg("num_policy"), py::arg("num_value"))
      .def("insert",
           [](S3FIFOCache& c, uint64_t hash, py::array_t<float, py::array::c_style | py::array::forcecast> policy,
              py::array_t<float, py::array::c_style | py::array::forcecast> value) {
             if ((uint32_t)policy.size() != c.num_policy() || (uint32_t)value.size() != c.num_value())
               throw std::runtime_error("S3FIFOCache.insert: wrong policy / value length");
             c.insert(hash, policy.data(), value.data());
           },
           py::arg("hash"), py::arg("policy"), py::arg("value"))
      .def("hits", &S3FIFOCache::hits).def("misses", &S3FIFOCache::misses).def("evictions", &S3FIFOCache::evictions)
      .def("reinserts", &S3FIFOCache::reinserts).def("size", &S3FIFOCache::size).def("max_size", &S3FIFOCache::max_size);
  py::class_<ShardedS3FIFOCache, std::shared_ptr<ShardedS3FIFOCache>>(m, "ShardedS3FIFOCache")
      .def(py::init<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>(), py::arg("max_size"), py::arg("shards"), py::arg("ghost_size"),
           py::arg("num_policy"), py::arg("num_value"))
      .def("hits", &ShardedS3FIFOCache::hits).def("misses", &ShardedS3FIFOCache::misses)
      .def("evictions", &ShardedS3FIFOCache::evictions).def("reinserts", &ShardedS3FIFOCache::reinserts)
      .def("size", &ShardedS3FIFOCache::size).def("max_size", &ShardedS3FIFOCache::max_size);

  // playout_eval / playout_eval_batch (py_wrapper.cc:726-770, game_state.cc:10-95): uniform prior over the legal moves of
  // the given state, value = the outcome of ONE uniformly random playout (relative values where the game uses them).
  // A host-side tool over the Python-visible GameState objects; the engines' evaluators run on the device.
  auto playout_one = [](const GameState& gs, std::default_random_engine& re) {
    py::array_t<uint8_t> valids = gs.valid_moves();
    const uint32_t A = gs.num_moves();
    py::array_t<float> policy(A);
    uint8_t wrap = 0;  // Vector<uint8_t>::sum() wraps mod 256 in the reference (game_state.cc:16)
    for (uint32_t i = 0; i < A; ++i) wrap = (uint8_t)(wrap + valids.at(i));
    const float sum = (float)wrap;
    for (uint32_t i = 0; i < A; ++i) policy.mutable_at(i) = sum > 0.0f ? (float)valids.at(i) / sum : 0.0f;
    auto sim = gs.copy();
    for (;;) {
      if (!sim->scores().is_none()) break;
      py::array_t<uint8_t> v = sim->valid_moves();
      std::vector<uint32_t> idx;
      for (uint32_t i = 0; i < A; ++i)
        if (v.at(i) != 0) idx.push_back(i);
      if (idx.empty()) break;
      std::uniform_int_distribution<uint32_t> dist{0, (uint32_t)idx.size() - 1};
      sim->play_move(idx[dist(re)]);
    }
    py::object sc = sim->scores();
    const int P = gs.num_players();
    py::array_t<float> value(P + 1);
    if (!sc.is_none()) {
      auto a = sc.cast<py::array_t<float>>();
      for (int i = 0; i <= P; ++i) value.mutable_at(i) = a.at(i);
      if (gs.relative_values()) {  // absolute_to_relative (game_state.h:24-34): rotate so that index 0 is the mover
        std::vector<float> rel(P + 1);
        const int cp = gs.current_player();
        for (int i = 0; i < P; ++i) rel[i] = a.at((i + cp) % P);
        rel[P] = a.at(P);
        for (int i = 0; i <= P; ++i) value.mutable_at(i) = rel[i];
      }
    } else {
      for (int i = 0; i <= P; ++i) value.mutable_at(i) = (float)(1.0 / (P + 1));
    }
    return std::make_pair(value, policy);
  };
  m.def("playout_eval", [playout_one](const GameState& gs) {
    static thread_local std::default_random_engine re{std::random_device{}()};
    auto r = playout_one(gs, re);
    return py::make_tuple(r.first, r.second);
  });
  m.def("playout_eval_batch", [playout_one](py::list states) {
    static thread_local std::default_random_engine re{std::random_device{}()};
    const ssize_t n = (ssize_t)py::len(states);
    if (n == 0) throw std::runtime_error("playout_eval_batch: empty list");
    std::vector<std::pair<py::array_t<float>, py::array_t<float>>> rs;
    for (auto& item : states) rs.push_back(playout_one(*item.cast<const GameState*>(), re));
    py::array_t<float> v({n, (ssize_t)rs[0].first.size()}), pi({n, (ssize_t)rs[0].second.size()});
    for (ssize_t i = 0; i < n; ++i) {
      for (ssize_t j = 0; j < rs[0].first.size(); ++j) v.mutable_at(i, j) = rs[i].first.at(j);
      for (ssize_t j = 0; j < rs[0].second.size(); ++j) pi.mutable_at(i, j) = rs[i].second.at(j);
    }
    return py::make_tuple(v, pi);
  });

  // Tracy hooks (py_wrapper.cc:772-787): profiling is off in this build
  m.def("_tracy_zone_begin", [](const std::string&, const std::string&, uint32_t) {});
  m.def("_tracy_zone_end", [] {});
  m.def("tracy_frame_mark", [] {});
  m.def("_tracy_set_thread_name", [](const std::string&) {});
  m.def("tracy_is_enabled", [] { return false; });
}
