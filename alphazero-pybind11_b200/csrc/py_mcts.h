// py_mcts.h — the `MCTS` class of the `alphazero` module (py_wrapper.cc:191-220; mcts.h:50-150): ONE tree of the device's
// wide-tree search (b2az_forest_*, n_trees = 1) behind the reference's single-tree API, for the Python tools that drive a
// tree themselves (play.py, mcts_analysis.py, frozen_eval.py). Selection, expansion (with the child shuffle), priors,
// backup, re-rooting, Gumbel bookkeeping and the policy read-outs run on the device; the object keeps a host copy of the
// root position so that find_leaf can return the leaf GameState (the root copy replayed along the device's path).
// Games: Connect4, the tafl games and Star Gambit.
#pragma once

class MCTS {
 public:
  MCTS(float cpuct, uint32_t num_players, uint32_t num_moves, float epsilon = 0.0f, float root_policy_temp = 1.0f,
       float fpu_reduction = 0.0f, bool relative_values = false, bool root_fpu_zero = false, bool shaped_dirichlet = false,
       bool gumbel_enabled = false, uint32_t gumbel_m = 16, float gumbel_c_visit = 50.0f, float gumbel_c_scale = 1.0f,
       bool gumbel_full = false)
      : num_players_(num_players), num_moves_(num_moves) {
    if (num_players != 2) throw std::runtime_error("MCTS: the B200 search implements two-player games");
    std::memset(&fp_, 0, sizeof(fp_));
    fp_.n_trees = 1;
    fp_.words_per_tree = 1u << 24;  // 64 MB of node slab for the one tree
    fp_.cpuct = cpuct; fp_.fpu_reduction = fpu_reduction; fp_.epsilon = epsilon; fp_.root_policy_temp = root_policy_temp;
    fp_.root_fpu_zero = root_fpu_zero; fp_.relative_values = relative_values; fp_.gumbel_enabled = gumbel_enabled;
    fp_.gumbel_m = gumbel_m; fp_.gumbel_c_visit = gumbel_c_visit; fp_.gumbel_c_scale = gumbel_c_scale;
    fp_.shaped_dirichlet = shaped_dirichlet;
    fp_.gumbel_full = gumbel_full;
    fp_.max_in_flight = 64;
    fp_.seed = std::random_device{}();  // the reference's thread-local generator starts from random_device as well
  }
  ~MCTS() { if (f_) b2az_forest_destroy(f_); }
  MCTS(const MCTS&) = delete;
  MCTS& operator=(const MCTS&) = delete;

  // (additive) a reproducible stream: tree == the reference's MCTS driven after MCTS::seed_thread_rng(seed)
  void seed(uint64_t s) {
    if (f_) throw std::runtime_error("MCTS.seed: the tree exists already");
    fp_.seed = s;
  }
  void update_root(const GameState& gs, uint32_t move) {
    ensure(gs);
    const uint32_t mv = move;
    check(b2az_forest_update_root(f_, nullptr, &mv), "update_root");
    tree_error("update_root");
    root_->play_move(move);
    last_leaf_.reset();
    inflight_.clear();
  }
  std::unique_ptr<GameState> find_leaf(const GameState& gs) {
    ensure(gs);
    check(b2az_forest_find_leaf(f_, nullptr, nullptr), "find_leaf");
    last_leaf_ = leaf_state(-1);
    tree_error("find_leaf");
    return last_leaf_->copy();
  }
  void process_result(const GameState& gs, py::array_t<float, py::array::c_style> value,
                      py::array_t<float, py::array::c_style | py::array::forcecast> pi, bool root_noise_enabled) {
    ensure(gs);
    if (!last_leaf_) throw std::runtime_error("MCTS.process_result: no pending leaf (call find_leaf first)");
    if (value.size() != 3 || (uint32_t)pi.size() != num_moves_) throw std::runtime_error("Eigen is angry!!!");
    check(b2az_forest_process_result_host(f_, nullptr, value.data(), pi.data(), root_noise_enabled ? 1 : 0), "process_result");
    settle_value(*last_leaf_, value);
    last_leaf_.reset();
  }
  std::unique_ptr<GameState> find_leaf_batched(const GameState& gs) {
    ensure(gs);
    const int slot = (int)inflight_.size();
    if (slot >= (int)fp_.max_in_flight) throw std::runtime_error("MCTS.find_leaf_batched: more than 64 leaves in flight");
    check(b2az_forest_find_leaf_batched(f_, nullptr, nullptr), "find_leaf_batched");
    inflight_.push_back(leaf_state(slot));
    tree_error("find_leaf_batched");
    return inflight_.back()->copy();
  }
  void process_result_batched(const GameState& gs, uint32_t leaf_index, py::array_t<float, py::array::c_style> value,
                              py::array_t<float, py::array::c_style | py::array::forcecast> pi, bool root_noise_enabled) {
    ensure(gs);
    if (leaf_index >= inflight_.size()) throw std::out_of_range("vector::_M_range_check");  // in_flight_.at(leaf_index)
    if (value.size() != 3 || (uint32_t)pi.size() != num_moves_) throw std::runtime_error("Eigen is angry!!!");
    check(b2az_forest_process_result_batched(f_, nullptr, leaf_index, value.data(), pi.data(), root_noise_enabled ? 1 : 0, 1),
          "process_result_batched");
    settle_value(*inflight_[leaf_index], value);
  }
  uint32_t in_flight_count() const { return (uint32_t)inflight_.size(); }
  void reset_batch() {
    inflight_.clear();
    if (f_) check(b2az_forest_reset_batch(f_, nullptr), "reset_batch");
  }
  py::array_t<float> root_value() {
    py::array_t<float> a(3);
    uint32_t info[16] = {0};
    if (f_) check(b2az_forest_counts(f_, nullptr, nullptr, nullptr, info), "root_value");
    else { const float l = 1.0f; std::memcpy(&info[13], &l, 4); }  // no search yet: w = 0, l = 1, d = 0
    std::memcpy(a.mutable_data(), &info[12], 12);
    return a;
  }
  py::array_t<uint32_t> counts() {
    py::array_t<uint32_t> a(num_moves_);
    std::memset(a.mutable_data(), 0, 4ull * num_moves_);
    if (f_) check(b2az_forest_counts(f_, nullptr, a.mutable_data(), nullptr, nullptr), "counts");
    return a;
  }
  py::array_t<float> root_q_values() {
    py::array_t<float> a(num_moves_);
    std::memset(a.mutable_data(), 0, 4ull * num_moves_);
    if (f_) check(b2az_forest_counts(f_, nullptr, nullptr, a.mutable_data(), nullptr), "root_q_values");
    return a;
  }
  py::array_t<float> probs_impl(float temp, int pruned) {
    py::array_t<float> a(num_moves_);
    std::memset(a.mutable_data(), 0, 4ull * num_moves_);
    if (f_) check(b2az_forest_probs(f_, nullptr, temp, pruned, 0, a.mutable_data(), nullptr), "probs");
    return a;
  }
  py::array_t<float> probs(float temp) { return probs_impl(temp, 0); }
  py::array_t<float> probs_pruned(float temp) { return probs_impl(temp, 1); }
  py::array_t<uint32_t> principal_variation(uint32_t depth) {
    std::vector<uint32_t> mv(std::max(1u, depth));
    uint32_t len = 0;
    if (f_ && depth) check(b2az_forest_principal_variation(f_, nullptr, depth, mv.data(), &len), "principal_variation");
    py::array_t<uint32_t> a(len);
    for (uint32_t i = 0; i < len; ++i) a.mutable_at(i) = mv[i];
    return a;
  }
  uint32_t info_word(int i) {
    uint32_t info[16] = {0};
    if (f_) check(b2az_forest_counts(f_, nullptr, nullptr, nullptr, info), "info");
    return info[i];
  }
  uint32_t depth() { return info_word(0); }
  uint32_t root_n() { return info_word(1); }
  void add_root_noise() {  // (with epsilon == 0 the mix leaves the priors as they are: nothing to do)
    if (f_ && fp_.epsilon > 0.0f) check(b2az_forest_root_ops(f_, nullptr, 0, 1), "add_root_noise");
  }
  void apply_root_policy_temp() {
    if (f_) check(b2az_forest_root_ops(f_, nullptr, 1, 0), "apply_root_policy_temp");
  }
  void set_gumbel_num_sims(uint32_t n) {
    pending_gumbel_ = n;
    have_pending_gumbel_ = true;
    if (f_ && fp_.gumbel_enabled) { check(b2az_forest_set_gumbel_num_sims(f_, nullptr, n), "set_gumbel_num_sims"); have_pending_gumbel_ = false; }
  }
  bool gumbel_enabled() const { return fp_.gumbel_enabled != 0; }
  py::array_t<float> gumbel_improved_policy() {
    py::array_t<float> a(num_moves_);
    std::memset(a.mutable_data(), 0, 4ull * num_moves_);
    if (f_ && fp_.gumbel_enabled) check(b2az_forest_gumbel_result(f_, nullptr, nullptr, a.mutable_data()), "gumbel_improved_policy");
    return a;
  }
  uint32_t gumbel_final_action() {
    uint32_t a = 0;
    if (!f_ || !fp_.gumbel_enabled) throw std::runtime_error("MCTS.gumbel_final_action: no Gumbel search has run");
    check(b2az_forest_gumbel_result(f_, nullptr, &a, nullptr), "gumbel_final_action");
    return a;
  }
  // MCTS::pick_move (mcts.cc:717-735): one uniform draw, the first move whose running sum exceeds it, else the last
  // positive entry. A host function in the reference as well (it is static and takes the vector from Python).
  static uint32_t pick_move(py::array_t<float, py::array::c_style | py::array::forcecast> p) {
    thread_local std::mt19937_64 eng{std::random_device{}()};
    const float choice = std::uniform_real_distribution<float>(0.0f, 1.0f)(eng);
    float sum = 0.0f;
    uint32_t last_pos = 0;
    bool any = false;
    for (ssize_t m = 0; m < p.size(); ++m) {
      const float v = p.data()[m];
      if (v > 0.0f) { last_pos = (uint32_t)m; any = true; }
      sum += v;
      if (sum > choice) return (uint32_t)m;
    }
    if (!any) throw std::runtime_error("pick_move: no positive probability");
    return last_pos;
  }

 private:
  void check(int rc, const char* what) {
    if (rc != 0) throw_last(what);
  }
  void tree_error(const char* what) {
    uint32_t info[16] = {0};
    check(b2az_forest_counts(f_, nullptr, nullptr, nullptr, info), what);
    if (info[7] & 1u) throw std::runtime_error(std::string("MCTS.") + what + ": the tree outgrew its 64 MB node slab");
    if (info[7] & 8u) throw std::runtime_error(std::string("MCTS.") + what + ": unknown move (not a child of the root)");
    if (info[7]) throw std::runtime_error(std::string("MCTS.") + what + ": device search error " + std::to_string(info[7]));
  }
  // the leaf position of a pending leaf: the root copy replayed along the device's path
  std::unique_ptr<GameState> leaf_state(int slot) {
    uint32_t path[96], len = 0;
    check(b2az_forest_leaf_path(f_, nullptr, slot, path, &len), "leaf_path");
    auto leaf = root_->copy();
    for (uint32_t i = 0; i < len; ++i) leaf->play_move(path[i]);
    return leaf;
  }
  // process_result writes through its `value` argument (mcts.cc:503-504, 522-524): a terminal leaf's scores replace the
  // evaluation, a relative-values evaluation comes back rotated to absolute seats
  void settle_value(const GameState& leaf, py::array_t<float, py::array::c_style>& value) {
    float* v = value.mutable_data();
    py::object sc = leaf.scores();
    if (!sc.is_none()) {
      auto a = sc.cast<py::array_t<float>>();
      for (int i = 0; i < 3; ++i) v[i] = a.at(i);
    } else if (fp_.relative_values && leaf.current_player() == 1) {
      std::swap(v[0], v[1]);
    }
  }
  template <int GAME>
  bool try_tafl(const GameState& gs) {
    auto* t = dynamic_cast<const TaflGS<GAME>*>(&gs);
    if (!t) return false;
    fp_.game = GAME;
    fp_.max_turns = t->s.max_turns;
    create();
    check(b2az_forest_set_root(f_, 0, &t->s, (uint32_t)sizeof(t->s), t->hist.data(), t->hist_len), "set_root");
    return true;
  }
  void create() {
    if (b2az_forest_create(&fp_, 0, &f_) != 0) throw_last("MCTS");
  }
  void ensure(const GameState& gs) {
    if (gs.num_moves() != num_moves_) throw std::runtime_error("MCTS: num_moves does not match the GameState");
    if (f_) return;
    if (!(try_tafl<B2AZ_TAFL_BRANDUBH>(gs) || try_tafl<B2AZ_TAFL_OPENTAFL>(gs) || try_tafl<B2AZ_TAFL_TAWLBWRDD>(gs))) {
      if (auto* c4 = dynamic_cast<const Connect4GS*>(&gs)) {  // the root record: stones in the first words (FGame<B2AZ_FOREST_C4>)
        fp_.game = 30u;
        fp_.max_turns = 64;
        create();
        b2az::TaflState ts;
        std::memset(&ts, 0, sizeof(ts));
        ts.king.lo = c4->s.p[0]; ts.king.hi = c4->s.p[1]; ts.turn = c4->s.turn; ts.player = (uint8_t)c4->s.player;
        check(b2az_forest_set_root(f_, 0, &ts, (uint32_t)sizeof(ts), nullptr, 0), "set_root");
      } else {
        auto* sg = dynamic_cast<const StarGambitBase*>(&gs);
        if (!sg) throw std::runtime_error("MCTS: the B200 single-tree search implements Connect4, the tafl games and Star Gambit");
        fp_.game = (sg->unified ? 20u : 10u) + sg->s.variant;
        fp_.max_turns = 512;
        create();
        check(b2az_forest_set_root(f_, 0, &sg->s, (uint32_t)sizeof(sg->s), sg->hist.data(), (uint32_t)sg->hist.size()), "set_root");
      }
    }
    root_ = gs.copy();
    if (have_pending_gumbel_ && fp_.gumbel_enabled) {
      check(b2az_forest_set_gumbel_num_sims(f_, nullptr, pending_gumbel_), "set_gumbel_num_sims");
      have_pending_gumbel_ = false;
    }
  }

  uint32_t num_players_, num_moves_;
  b2az_forest_params fp_;
  b2az_forest* f_ = nullptr;
  std::unique_ptr<GameState> root_, last_leaf_;
  std::vector<std::unique_ptr<GameState>> inflight_;
  uint32_t pending_gumbel_ = 0;
  bool have_pending_gumbel_ = false;
};

inline void bind_mcts(py::module_& m) {
  py::class_<MCTS>(m, "MCTS")
      .def(py::init<float, uint32_t, uint32_t>())
      .def(py::init<float, uint32_t, uint32_t, float, float, float>())
      .def(py::init<float, uint32_t, uint32_t, float, float, float, bool>())
      .def(py::init<float, uint32_t, uint32_t, float, float, float, bool, bool, bool>())
      .def(py::init<float, uint32_t, uint32_t, float, float, float, bool, bool, bool, bool, uint32_t, float, float, bool>())
      .def("update_root", &MCTS::update_root)
      .def("find_leaf", &MCTS::find_leaf)
      .def("process_result", &MCTS::process_result, py::arg("gs"), py::arg("value"), py::arg("pi"),
           py::arg("root_noise_enabled") = false)
      .def("root_value", &MCTS::root_value)
      .def("counts", &MCTS::counts)
      .def("root_q_values", &MCTS::root_q_values)
      .def("probs", &MCTS::probs)
      .def("probs_pruned", &MCTS::probs_pruned)
      .def("principal_variation", &MCTS::principal_variation, py::arg("depth") = 5)
      .def("depth", &MCTS::depth)
      .def("add_root_noise", &MCTS::add_root_noise)
      .def("apply_root_policy_temp", &MCTS::apply_root_policy_temp)
      .def("root_n", &MCTS::root_n)
      .def("find_leaf_batched", &MCTS::find_leaf_batched)
      .def("process_result_batched", &MCTS::process_result_batched, py::arg("gs"), py::arg("leaf_index"), py::arg("value"),
           py::arg("pi"), py::arg("root_noise_enabled") = false)
      .def("in_flight_count", &MCTS::in_flight_count)
      .def("reset_batch", &MCTS::reset_batch)
      .def("set_gumbel_num_sims", &MCTS::set_gumbel_num_sims)
      .def("gumbel_enabled", &MCTS::gumbel_enabled)
      .def("gumbel_improved_policy", &MCTS::gumbel_improved_policy)
      .def("gumbel_final_action", &MCTS::gumbel_final_action)
      .def("seed", &MCTS::seed)  // additive: MCTS::seed_thread_rng is C++-only in the reference (mcts.h:149)
      .def_static("pick_move", &MCTS::pick_move);
}
