// oracle/az_oracle.cc — CPU restatement of the reference's self-play hot path (Connect4, PUCT).
//
// TEST INFRASTRUCTURE ONLY (see az_oracle.h for who may load it and for the parity status).
// Written to be READ next to the reference: plain pointer-and-vector data structures in the
// reference's own shapes (array-of-structs nodes, int8 board, FIFO queues), sequential float
// arithmetic in the reference's operation order, and libstdc++'s own <random>/<algorithm> for every
// random draw. It shares no code with the product (alphazero-pybind11_b200/csrc): that is the point.
//
// Build: g++ -std=c++20 -O3 -fPIC -shared (no -march: the reference's release build has none, so
// no a*b+c contraction happens on x86-64).
#include "az_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <random>
#include <vector>

namespace {

// ---------------------------------------------------------------------------- pcg32
// setseq_xsh_rr_64_32 (pcg_random.hpp:1663): 64-bit LCG state, output = XSH-RR of the state BEFORE
// the step; engine(seed): inc = default, state = 0; bump; state += seed; bump  (484-490);
// engine(seed, stream): inc = stream*2+1 (494-501).
struct Pcg {
  using result_type = uint32_t;
  static constexpr uint64_t kMult = 6364136223846793005ULL;
  uint64_t state, inc;
  explicit Pcg(uint64_t seed) : inc(1442695040888963407ULL) { init(seed); }
  Pcg(uint64_t seed, uint64_t stream) : inc((stream << 1) | 1ULL) { init(seed); }
  void init(uint64_t seed) {
    state = 0;
    state = state * kMult + inc;
    state += seed;
    state = state * kMult + inc;
    // == (seed + inc) * mult + inc
  }
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return 0xFFFFFFFFu; }
  result_type operator()() {
    const uint64_t old = state;
    state = old * kMult + inc;
    const uint32_t x = static_cast<uint32_t>(((old >> 18) ^ old) >> 27);
    const uint32_t rot = static_cast<uint32_t>(old >> 59);
    return (x >> rot) | (x << ((-rot) & 31));
  }
};

// ---------------------------------------------------------------------------- Connect4
constexpr int H = 6, W = 7, A = 7, P = 2;
struct Board {
  int8_t b[2][H][W];
  uint8_t player;
  uint32_t turn;
};
void c4_clear(Board& s) { std::memset(&s, 0, sizeof(s)); }
// connect4_gs.cc:39-46
void c4_valid(const Board& s, uint8_t* v) {
  for (int w = 0; w < W; ++w) v[w] = (s.b[0][0][w] == 0 && s.b[1][0][w] == 0) ? 1 : 0;
}
// connect4_gs.cc:48-58
bool c4_play(Board& s, uint32_t m) {
  for (int h = H - 1; h >= 0; --h) {
    if (s.b[0][h][m] == 0 && s.b[1][h][m] == 0) {
      s.b[s.player][h][m] = 1;
      s.player = (s.player + 1) % 2;
      ++s.turn;
      return true;
    }
  }
  return false;
}
// connect4_gs.cc:60-129. Returns 0 (not over) or 1 + index of the one-hot score.
int c4_term(const Board& s) {
  for (int p = 0; p < 2; ++p) {
    for (int h = 0; h < H; ++h) {
      int run = 0;
      for (int w = 0; w < W; ++w) {
        run = (s.b[p][h][w] == 1) ? run + 1 : 0;
        if (run == 4) return 1 + p;
      }
    }
    for (int w = 0; w < W; ++w) {
      int run = 0;
      for (int h = 0; h < H; ++h) {
        run = (s.b[p][h][w] == 1) ? run + 1 : 0;
        if (run == 4) return 1 + p;
      }
    }
    for (int h = 0; h < H - 3; ++h) {
      for (int w = 0; w < W - 3; ++w) {
        bool all = true;
        for (int x = 0; x < 4; ++x) all = all && s.b[p][h + x][w + x] != 0;
        if (all) return 1 + p;
      }
      for (int w = W - 4; w < W; ++w) {
        bool all = true;
        for (int x = 0; x < 4; ++x) all = all && s.b[p][h + x][w - x] != 0;
        if (all) return 1 + p;
      }
    }
  }
  uint8_t v[A];
  c4_valid(s, v);
  for (int w = 0; w < W; ++w)
    if (v[w]) return 0;
  return 3;
}
// connect4_gs.cc:131-149
void c4_canon(const Board& s, float* out) {
  for (int p = 0; p < 2; ++p)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) out[(p * H + h) * W + w] = s.b[p][h][w];
  for (int i = 0; i < H * W; ++i) {
    out[(2 + s.player) * H * W + i] = 1.0f;
    out[(2 + (s.player + 1) % 2) * H * W + i] = 0.0f;
  }
}

// ---------------------------------------------------------------------------- Node / MCTS
struct Node {  // mcts.h:14-48
  float q = 0, d = 0, v = 0, policy = 0;
  uint32_t move = 0, n = 0;
  int8_t player = 0;
  int term = 0;  // 0 = scores == nullptr, else 1 + one-hot index
  std::vector<Node> children;
};

struct Tree {  // MCTS, mcts.h:50-201
  Node root;
  Node* current = nullptr;
  std::vector<Node*> path;
  uint32_t depth = 0, total_leaf_depth = 0;
  // Gumbel state (mcts.h:163-176)
  uint32_t gumbel_num_sims_target = 0;
  bool gumbel_initialized = false;
  uint32_t gumbel_effective_m = 0;
  std::vector<float> gumbel_g;
  std::vector<size_t> gumbel_survivors;
  std::vector<std::pair<uint32_t, uint32_t>> gumbel_phases;
  size_t gumbel_phase_idx = 0;
  uint32_t gumbel_sims_in_phase = 0;
};

struct Cfg : azo_cfg {};

// mcts.cc:93-101
void add_children(Node& nd, const uint8_t* valids, Pcg& re) {
  for (int w = 0; w < A; ++w)
    if (valids[w] == 1) {
      Node c;
      c.move = static_cast<uint32_t>(w);
      nd.children.push_back(std::move(c));
    }
  std::shuffle(nd.children.begin(), nd.children.end(), re);
}
// mcts.cc:109-121
void set_policy_normalized(Node& nd, const float* pi, bool apply_temp, float inv_temp) {
  float sum = 0.0f;
  for (auto& c : nd.children) {
    float p = pi[c.move];
    if (apply_temp) p = std::pow(p, inv_temp);
    c.policy = p;
    sum += p;
  }
  for (auto& c : nd.children) c.policy /= sum;
}
// mcts.cc:123-149 (n_in_flight == 0 on the PlayManager path)
Node* best_child(Node& nd, float cpuct, float fpu_reduction) {
  float seen = 0.0f;
  for (const auto& c : nd.children)
    if (c.n > 0) seen += c.policy;
  const float fpu_value = nd.v - fpu_reduction * std::sqrt(seen);
  const float sqrt_n = std::sqrt(static_cast<float>(nd.n));
  auto uct = [&](const Node& c) {
    return (c.n == 0 ? fpu_value : c.q) + cpuct * c.policy * sqrt_n / static_cast<float>(c.n + 1);
  };
  size_t best_i = 0;
  float best = uct(nd.children[0]);
  for (size_t i = 1; i < nd.children.size(); ++i) {
    const float u = uct(nd.children[i]);
    if (u > best) {
      best = u;
      best_i = i;
    }
  }
  return &nd.children[best_i];
}
// mcts.cc:403-446
void add_root_noise(Tree& t, const Cfg& c, Pcg& re) {
  const size_t k = t.root.children.size();
  float noise[A] = {0};
  double sum = 0.0;
  if (c.shaped_dirichlet && k > 1) {
    const float N = static_cast<float>(k);
    float log_sum = 0.0f;
    for (const auto& ch : t.root.children) log_sum += std::log(std::min(ch.policy, 0.01f) + 1e-20f);
    const float log_mean = log_sum / N;
    float shaped_sum = 0.0f;
    for (const auto& ch : t.root.children) {
      const float lp = std::log(std::min(ch.policy, 0.01f) + 1e-20f);
      shaped_sum += std::max(0.0f, lp - log_mean);
    }
    const float uniform = 1.0f / N;
    for (auto& ch : t.root.children) {
      const float lp = std::log(std::min(ch.policy, 0.01f) + 1e-20f);
      const float shaped = std::max(0.0f, lp - log_mean);
      float alpha_prop = (shaped_sum > 0) ? 0.5f * (shaped / shaped_sum + uniform) : uniform;
      alpha_prop = std::max(alpha_prop, 1e-6f);
      std::gamma_distribution<float> dist{10.83f * alpha_prop, 1.0f};
      noise[ch.move] = dist(re);
      sum += noise[ch.move];
    }
  } else {
    std::gamma_distribution<float> dist{10.83f / static_cast<float>(k), 1.0};
    for (auto& ch : t.root.children) {
      noise[ch.move] = dist(re);
      sum += noise[ch.move];
    }
  }
  for (auto& ch : t.root.children)
    ch.policy = ch.policy * (1 - c.epsilon) + c.epsilon * noise[ch.move] / static_cast<float>(sum);
}
// mcts.cc:448-460
void apply_root_policy_temp(Tree& t, const Cfg& c) {
  if (c.mcts_root_temp == 1.0f) return;
  float sum = 0.0f;
  for (auto& ch : t.root.children) {
    ch.policy = std::pow(ch.policy, 1.0f / c.mcts_root_temp);
    sum += ch.policy;
  }
  if (sum > 0.0f)
    for (auto& ch : t.root.children) ch.policy /= sum;
}
// mcts.cc:462-498 (PUCT branch). Returns the leaf position in `leaf`.
// ---------------------------------------------------------------------------- Gumbel (mcts.cc:24-89, 175-401)
constexpr float GUMBEL_LOG_FLOOR = 1e-20f;
// mcts.cc:28-66
std::vector<std::pair<uint32_t, uint32_t>> seq_halving_phase_plan(uint32_t m, uint32_t n) {
  std::vector<std::pair<uint32_t, uint32_t>> phases;
  if (m <= 1) {
    phases.emplace_back(1, n);
    return phases;
  }
  uint32_t log2m = 0;
  for (uint32_t v = m - 1; v > 0; v >>= 1) ++log2m;
  if (log2m == 0) log2m = 1;
  const uint32_t base_v = std::max<uint32_t>(1u, n / (log2m * m));
  uint32_t sims_used = 0, num_c = m;
  for (uint32_t phase_idx = 0; phase_idx < log2m; ++phase_idx) {
    if (sims_used >= n) break;
    const uint32_t remaining = n - sims_used;
    const bool is_final = (phase_idx == log2m - 1);
    uint32_t v_per = is_final ? std::max<uint32_t>(1u, remaining / num_c) : base_v * (1u << phase_idx);
    if (num_c * v_per > remaining) {
      v_per = remaining / num_c;
      if (v_per == 0) {
        num_c = remaining;
        v_per = 1;
      }
    }
    phases.emplace_back(num_c, v_per);
    sims_used += num_c * v_per;
    num_c = std::max<uint32_t>(1u, num_c / 2);
  }
  return phases;
}
// mcts.cc:71-89
float compute_v_mix_from_children(float raw_v, const std::vector<float>& qs, const std::vector<uint32_t>& ns,
                                  const std::vector<float>& priors) {
  float sum_visits = 0.0f, sum_priors_visited = 0.0f, weighted_num = 0.0f;
  for (size_t i = 0; i < qs.size(); ++i) {
    sum_visits += static_cast<float>(ns[i]);
    if (ns[i] > 0) {
      sum_priors_visited += priors[i];
      weighted_num += priors[i] * qs[i];
    }
  }
  if (sum_priors_visited <= 0.0f) return raw_v;
  const float weighted_q = weighted_num / sum_priors_visited;
  return (raw_v + sum_visits * weighted_q) / (sum_visits + 1.0f);
}
// mcts.cc:180-188, 175-178
void reset_gumbel_state(Tree& t) {
  t.gumbel_initialized = false;
  t.gumbel_effective_m = 0;
  t.gumbel_g.clear();
  t.gumbel_survivors.clear();
  t.gumbel_phases.clear();
  t.gumbel_phase_idx = 0;
  t.gumbel_sims_in_phase = 0;
}
void set_gumbel_num_sims(Tree& t, uint32_t n) {
  t.gumbel_num_sims_target = n;
  reset_gumbel_state(t);
}
// mcts.cc:190-227
void init_gumbel_state(Tree& t, const Cfg& c, Pcg& re) {
  const auto num_legal = static_cast<uint32_t>(t.root.children.size());
  if (num_legal == 0) return;
  const uint32_t remaining = t.depth < t.gumbel_num_sims_target ? t.gumbel_num_sims_target - t.depth : 0;
  if (remaining == 0) return;
  t.gumbel_effective_m = std::max<uint32_t>(1u, std::min({c.gumbel_m, num_legal, remaining}));
  std::extreme_value_distribution<float> gumbel_dist{0.0f, 1.0f};
  t.gumbel_g.resize(num_legal);
  for (uint32_t i = 0; i < num_legal; ++i) t.gumbel_g[i] = gumbel_dist(re);
  std::vector<size_t> idx(num_legal);
  for (uint32_t i = 0; i < num_legal; ++i) idx[i] = i;
  std::partial_sort(idx.begin(), idx.begin() + t.gumbel_effective_m, idx.end(), [&t](size_t a, size_t b) {
    const float la = std::log(t.root.children[a].policy + GUMBEL_LOG_FLOOR);
    const float lb = std::log(t.root.children[b].policy + GUMBEL_LOG_FLOOR);
    return t.gumbel_g[a] + la > t.gumbel_g[b] + lb;
  });
  t.gumbel_survivors.assign(idx.begin(), idx.begin() + t.gumbel_effective_m);
  t.gumbel_phases = seq_halving_phase_plan(t.gumbel_effective_m, remaining);
  t.gumbel_phase_idx = 0;
  t.gumbel_sims_in_phase = 0;
  t.gumbel_initialized = true;
}
// mcts.cc:229-264
void gumbel_advance_phase(Tree& t, const Cfg& c) {
  if (t.gumbel_phase_idx + 1 >= t.gumbel_phases.size()) return;
  const auto next_num_c = t.gumbel_phases[t.gumbel_phase_idx + 1].first;
  if (next_num_c >= t.gumbel_survivors.size()) {
    ++t.gumbel_phase_idx;
    t.gumbel_sims_in_phase = 0;
    return;
  }
  uint32_t max_visit = 0;
  for (const auto child_idx : t.gumbel_survivors) max_visit = std::max(max_visit, t.root.children[child_idx].n);
  const float sigma_scale = (c.gumbel_c_visit + static_cast<float>(max_visit)) * c.gumbel_c_scale;
  std::vector<std::pair<float, size_t>> scored;
  for (const auto child_idx : t.gumbel_survivors) {
    const auto& ch = t.root.children[child_idx];
    const float logit = std::log(ch.policy + GUMBEL_LOG_FLOOR);
    const float q_hat = ch.n > 0 ? ch.q : 0.0f;
    const float score = t.gumbel_g[child_idx] + logit + sigma_scale * q_hat;
    scored.emplace_back(score, child_idx);
  }
  std::partial_sort(scored.begin(), scored.begin() + next_num_c, scored.end(),
                    [](const auto& a, const auto& b) { return a.first > b.first; });
  t.gumbel_survivors.resize(next_num_c);
  for (size_t i = 0; i < next_num_c; ++i) t.gumbel_survivors[i] = scored[i].second;
  ++t.gumbel_phase_idx;
  t.gumbel_sims_in_phase = 0;
}
// mcts.cc:266-283
size_t gumbel_next_root_child(Tree& t, const Cfg& c) {
  if (t.gumbel_phase_idx < t.gumbel_phases.size()) {
    const auto& [num_c, v_per] = t.gumbel_phases[t.gumbel_phase_idx];
    if (t.gumbel_sims_in_phase >= num_c * v_per) gumbel_advance_phase(t, c);
  }
  if (t.gumbel_survivors.empty()) return 0;
  const size_t pick = t.gumbel_sims_in_phase % t.gumbel_survivors.size();
  ++t.gumbel_sims_in_phase;
  return t.gumbel_survivors[pick];
}
// softmax(log prior + sigma * completedQ) shared by mcts.cc:285-334 and :336-373
float gumbel_z(const Node& node, const Cfg& c, std::vector<float>& z, std::vector<uint32_t>& ns, uint32_t& sum_visits) {
  const auto k = node.children.size();
  uint32_t max_visit = 0;
  sum_visits = 0;
  std::vector<float> qs(k), priors(k);
  ns.assign(k, 0);
  for (size_t i = 0; i < k; ++i) {
    max_visit = std::max(max_visit, node.children[i].n);
    sum_visits += node.children[i].n;
    qs[i] = node.children[i].q;
    ns[i] = node.children[i].n;
    priors[i] = node.children[i].policy;
  }
  const float v_mix = compute_v_mix_from_children(node.v, qs, ns, priors);
  const float sigma_scale = (c.gumbel_c_visit + static_cast<float>(max_visit)) * c.gumbel_c_scale;
  z.assign(k, 0.0f);
  float z_max = -std::numeric_limits<float>::infinity();
  for (size_t i = 0; i < k; ++i) {
    const float completed_q = ns[i] > 0 ? qs[i] : v_mix;
    z[i] = std::log(priors[i] + GUMBEL_LOG_FLOOR) + sigma_scale * completed_q;
    if (z[i] > z_max) z_max = z[i];
  }
  float z_sum = 0.0f;
  for (size_t i = 0; i < k; ++i) {
    z[i] = std::exp(z[i] - z_max);
    z_sum += z[i];
  }
  return z_sum;
}
// mcts.cc:285-334
size_t gumbel_interior_select(const Node& node, const Cfg& c) {
  std::vector<float> z;
  std::vector<uint32_t> ns;
  uint32_t sum_visits = 0;
  const float z_sum = gumbel_z(node, c, z, ns, sum_visits);
  const float inv = z_sum > 0 ? (1.0f / z_sum) : 0.0f;
  const float denom = 1.0f + static_cast<float>(sum_visits);
  size_t best = 0;
  float best_score = -std::numeric_limits<float>::infinity();
  for (size_t i = 0; i < z.size(); ++i) {
    const float pi_prime = z[i] * inv;
    const float score = pi_prime - static_cast<float>(ns[i]) / denom;
    if (score > best_score) {
      best_score = score;
      best = i;
    }
  }
  return best;
}
// mcts.cc:336-373
void gumbel_improved_policy(const Tree& t, const Cfg& c, float* out) {
  for (int m = 0; m < A; ++m) out[m] = 0.0f;
  if (t.root.children.empty()) return;
  std::vector<float> z;
  std::vector<uint32_t> ns;
  uint32_t sum_visits = 0;
  const float z_sum = gumbel_z(t.root, c, z, ns, sum_visits);
  if (z_sum <= 0) return;
  for (size_t i = 0; i < z.size(); ++i) out[t.root.children[i].move] = z[i] / z_sum;
}

void find_leaf(Tree& t, const Cfg& c, const Board& gs, Board& leaf, Pcg& re) {
  t.current = &t.root;
  leaf = gs;
  if (c.gumbel_enabled && !t.gumbel_initialized && t.gumbel_num_sims_target > 0 && t.root.n > 0 &&
      !t.root.children.empty()) {
    init_gumbel_state(t, c, re);  // :468-472
  }
  while (t.current->n > 0 && t.current->term == 0) {
    t.path.push_back(t.current);
    if (c.gumbel_enabled && t.gumbel_initialized && t.current == &t.root) {
      t.current = &t.root.children[gumbel_next_root_child(t, c)];
    } else if (c.gumbel_enabled && t.gumbel_initialized && c.gumbel_full) {
      t.current = &t.current->children[gumbel_interior_select(*t.current, c)];
    } else {
      const float fpu = (t.current == &t.root && c.root_fpu_zero) ? 0.0f : c.fpu_reduction;
      t.current = best_child(*t.current, c.cpuct, fpu);
    }
    c4_play(leaf, t.current->move);
  }
  t.total_leaf_depth += static_cast<uint32_t>(t.path.size());
  if (t.current->n == 0) {
    t.current->player = static_cast<int8_t>(leaf.player);
    t.current->term = c4_term(leaf);
    uint8_t valids[A];
    c4_valid(leaf, valids);
    add_children(*t.current, valids, re);
  }
}
// mcts.cc:500-555 (relative_values == false for Connect4)
void process_result(Tree& t, const Cfg& c, float* value, const float* pi, bool root_noise, Pcg& re) {
  if (t.current->term != 0) {
    for (int i = 0; i < P + 1; ++i) value[i] = (t.current->term == i + 1) ? 1.0f : 0.0f;
  } else if (t.current == &t.root) {
    set_policy_normalized(*t.current, pi, c.mcts_root_temp != 1.0f, 1.0f / c.mcts_root_temp);
    if (root_noise && !c.gumbel_enabled) add_root_noise(t, c, re);  // :514-518
  } else {
    set_policy_normalized(*t.current, pi, false, 1.0f);
  }
  const int32_t num_players = P;
  while (!t.path.empty()) {
    Node* parent = t.path.back();
    t.path.pop_back();
    float v = value[parent->player];
    v += value[num_players] / num_players;
    Node* cur = t.current;
    cur->q = (cur->q * static_cast<float>(cur->n) + v) / static_cast<float>(cur->n + 1);
    cur->d = (cur->d * static_cast<float>(cur->n) + value[num_players]) / static_cast<float>(cur->n + 1);
    if (cur->n == 0) cur->v = value[cur->player] + value[num_players] / num_players;
    ++cur->n;
    t.current = parent;
  }
  if (t.root.n == 0) {
    t.root.v = value[t.root.player] + value[num_players] / num_players;
    t.root.d = value[num_players];
  }
  ++t.depth;
  ++t.root.n;
}
// mcts.cc:151-173. Returns false for an unknown move.
bool update_root(Tree& t, const Board& gs, uint32_t move, Pcg& re) {
  t.depth = 0;
  t.total_leaf_depth = 0;
  if (t.root.children.empty()) {
    uint8_t valids[A];
    c4_valid(gs, valids);
    add_children(t.root, valids, re);
  }
  auto x = std::find_if(t.root.children.begin(), t.root.children.end(), [move](const Node& n) { return n.move == move; });
  if (x == t.root.children.end()) return false;
  Node tmp = std::move(*x);
  t.root = std::move(tmp);
  reset_gumbel_state(t);  // mcts.cc:172
  return true;
}
void counts_of(const Tree& t, uint32_t* out) {  // mcts.cc:557-564
  for (int m = 0; m < A; ++m) out[m] = 0;
  for (const auto& c : t.root.children) out[c.move] = c.n;
}
float seq_sum(const float* a) {  // Vector::sum() over num_moves entries, sequential (oracle/shim/Eigen/Dense)
  float s = 0.0f;
  for (int m = 0; m < A; ++m) s += a[m];
  return s;
}
// mcts.cc:575-618
void probs_of(const Tree& t, float temp, float* probs) {
  uint32_t counts[A];
  counts_of(t, counts);
  float fc[A];
  for (int m = 0; m < A; ++m) fc[m] = static_cast<float>(counts[m]);
  const float count_sum = seq_sum(fc);
  if (count_sum == 0) {
    for (int m = 0; m < A; ++m) probs[m] = 0.0f;
    for (const auto& c : t.root.children) probs[c.move] = c.policy;
    if (temp != 0.0f)
      for (int m = 0; m < A; ++m) probs[m] = std::pow(probs[m], 1.0f / temp);
    const float s = seq_sum(probs);
    for (int m = 0; m < A; ++m) probs[m] /= s;
    return;
  }
  if (temp == 0) {
    std::vector<int> best{0};
    uint32_t best_count = counts[0];
    for (int m = 1; m < A; ++m) {
      if (counts[m] > best_count) {
        best_count = counts[m];
        best.assign(1, m);
      } else if (counts[m] == best_count) {
        best.push_back(m);
      }
    }
    for (int m = 0; m < A; ++m) probs[m] = 0.0f;
    for (int m : best) probs[m] = 1.0 / best.size();
    return;
  }
  float s = seq_sum(fc);
  for (int m = 0; m < A; ++m) probs[m] = fc[m] / s;
  for (int m = 0; m < A; ++m) probs[m] = std::pow(probs[m], 1 / temp);
  s = seq_sum(probs);
  for (int m = 0; m < A; ++m) probs[m] /= s;
}
// mcts.cc:620-674
void probs_pruned_of(const Tree& t, const Cfg& c, float temp, float* out) {
  if (t.root.n <= 1) return probs_of(t, temp, out);
  const float explore_scaling = c.cpuct * std::sqrt(static_cast<float>(t.root.n));
  float best_sel = -1e30f;
  for (const auto& ch : t.root.children) {
    if (ch.n == 0) continue;
    const float sel = ch.q + explore_scaling * ch.policy / static_cast<float>(ch.n + 1);
    if (sel > best_sel) best_sel = sel;
  }
  float pruned[A] = {0};
  for (const auto& ch : t.root.children) {
    if (ch.n == 0) continue;
    const float gap = best_sel - ch.q;
    float desired;
    if (gap <= 0) desired = static_cast<float>(ch.n);
    else desired = explore_scaling * ch.policy / gap - 1.0f;
    pruned[ch.move] = std::min(static_cast<float>(ch.n), std::max(0.0f, desired));
  }
  const float total = seq_sum(pruned);
  if (total == 0) return probs_of(t, temp, out);
  if (temp == 0) {
    float best = pruned[0];
    for (int m = 1; m < A; ++m) best = std::max(best, pruned[m]);
    int cnt = 0;
    for (int m = 0; m < A; ++m) cnt += pruned[m] == best;
    for (int m = 0; m < A; ++m) out[m] = pruned[m] == best ? 1.0f / cnt : 0.0f;
    return;
  }
  for (int m = 0; m < A; ++m) out[m] = pruned[m] / total;
  if (temp != 1.0f) {
    for (int m = 0; m < A; ++m) out[m] = std::pow(out[m], 1.0f / temp);
    const float s = seq_sum(out);
    for (int m = 0; m < A; ++m) out[m] /= s;
  }
}
// mcts.cc:717-735; returns A when nothing is positive (the reference throws)
uint32_t pick_move(const float* p, Pcg& re) {
  std::uniform_real_distribution<float> dist{0.0f, 1.0f};
  const float choice = dist(re);
  float sum = 0.0f;
  for (uint32_t m = 0; m < A; ++m) {
    sum += p[m];
    if (sum > choice) return m;
  }
  for (int m = A - 1; m >= 0; --m)
    if (p[m] > 0) return static_cast<uint32_t>(m);
  return A;
}
// mcts.cc:737-750
float root_entropy(const Tree& t) {
  const float k = static_cast<float>(t.root.children.size());
  if (k <= 1 || t.root.n <= 1) return 0.0f;
  const float log_k = std::log(k);
  float entropy = 0.0f;
  const float total_n = static_cast<float>(t.root.n);
  for (const auto& c : t.root.children)
    if (c.n > 0) {
      const float p = static_cast<float>(c.n) / total_n;
      entropy -= p * std::log(p);
    }
  return entropy / log_k;
}
// mcts.h:78-100
void root_value(const Tree& t, float* wld) {
  float q = 0, d = 0;
  bool found = false;
  for (const auto& c : t.root.children)
    if (c.n > 0 && c.q > q) {
      q = c.q;
      d = c.d;
      found = true;
    }
  if (!found && t.root.n > 0) {
    q = t.root.v;
    d = t.root.d;
  }
  const float w = q - d / P;
  const double l = 1.0 - w - d;
  wld[0] = w;
  wld[1] = static_cast<float>(l);
  wld[2] = d;
}

// ---------------------------------------------------------------------------- PlayManager
struct Sample {  // PlayHistory, game_state.h:24-46
  float canon[168];
  float v[P + 1];
  float pi[A];
};
struct Game {  // GameData, play_manager.h:33-58
  Board gs;
  Tree mcts[P];
  float v[P + 1] = {0, 0, 0};
  float pi[A] = {0};
  float canonical[168];
  std::vector<Sample> partial;
  bool initialized = false, capped = false, playthrough = false;
  double total_avg_leaf_depth = 0, total_search_entropy = 0, total_valid_moves = 0;
  double fast_total_avg_leaf_depth = 0, fast_total_search_entropy = 0;
  uint32_t move_count = 0, full_move_count = 0, fast_move_count = 0;
  std::unique_ptr<Pcg> rng;
};
struct PM {
  Cfg cfg;
  std::vector<Game> games;
  std::deque<uint32_t> awaiting_mcts, awaiting_inference;
  std::deque<Sample> history;
  std::unique_ptr<Pcg> global_rng;
  uint32_t games_started = 0, games_completed = 0;
  uint64_t game_length = 0, total_move_count = 0, full_move_count = 0, fast_move_count = 0, simulations = 0, moves = 0;
  double total_avg_leaf_depth = 0, total_search_entropy = 0, total_valid_moves = 0;
  double fast_total_avg_leaf_depth = 0, fast_total_search_entropy = 0;
  float scores[3] = {0, 0, 0};
  float resign_scores[3] = {0, 0, 0};
  Pcg& re(uint32_t g) { return cfg.rng_mode == 1 ? *global_rng : *games[g].rng; }
};

// game_state.h:160-173
void dumb_eval(const Board& leaf, float* v, float* pi) {
  uint8_t valids[A];
  c4_valid(leaf, valids);
  for (int i = 0; i < P + 1; ++i) v[i] = static_cast<float>(1.0 / (P + 1));
  for (int m = 0; m < A; ++m) pi[m] = 0.0f;
  uint8_t s8 = 0;
  for (int m = 0; m < A; ++m) s8 = static_cast<uint8_t>(s8 + valids[m]);  // Vector<uint8_t>::sum() is uint8-typed
  const float sum = s8;
  if (sum == 0.0) return;
  for (int m = 0; m < A; ++m) pi[m] = static_cast<float>(valids[m]) / sum;
}

// One iteration of the loop body of PlayManager::play() for game i (play_manager.cc:277-599).
void play_iteration(PM& pm, uint32_t i) {
  const Cfg& c = pm.cfg;
  Game& game = pm.games[i];
  Pcg& re = pm.re(i);
  if (game.initialized) {
    const int cp = game.gs.player;
    Tree& mcts = game.mcts[cp];
    process_result(mcts, c, game.v, game.pi, c.epsilon > 0 && !game.capped, re);
    ++pm.simulations;
    const uint32_t goal_depth = game.capped ? c.playout_cap_depth : c.mcts_visits[cp];
    if (mcts.depth >= goal_depth) {
      float temp = c.start_temp;
      const float half_life = c.temp_decay_half_life;
      if (half_life != 0) {
        const uint32_t t = game.gs.turn;
        constexpr float ln2 = 0.693;
        const float lambda = ln2 / half_life;
        temp -= c.final_temp;
        temp *= std::exp(-lambda * t);
        temp += c.final_temp;
      }
      // resign_percent (play_manager.cc:305-337)
      int resign_term = 0;
      if (c.resign_percent > 0 && !game.playthrough) {
        float wld[3];
        root_value(mcts, wld);
        const double resign_val = 1.0 - c.resign_percent;
        int t = 0;
        if (wld[0] > resign_val) t = cp + 1;
        else if (wld[1] > resign_val) t = (cp + 1) % 2 + 1;
        else if (wld[2] > resign_val) t = P + 1;
        if (t != 0) {
          std::uniform_real_distribution<float> dist{0.0F, 1.0F};
          if (dist(re) < c.resign_playthrough_percent) game.playthrough = true;
          else resign_term = t;
        }
      }
      float pi[A];
      uint32_t chosen;
      if (c.gumbel_enabled && !game.capped) {
        // gumbel_final_action (mcts.cc:375-401); the improved-policy-sampling opt-in (G3) is a per-seat override
        if (!mcts.gumbel_initialized || mcts.gumbel_survivors.empty()) {
          probs_of(mcts, 0.0f, pi);
          chosen = pick_move(pi, re);
        } else {
          uint32_t max_visit = 0;
          for (const auto& ch : mcts.root.children) max_visit = std::max(max_visit, ch.n);
          const float sigma_scale = (c.gumbel_c_visit + static_cast<float>(max_visit)) * c.gumbel_c_scale;
          size_t best = mcts.gumbel_survivors[0];
          float best_score = -std::numeric_limits<float>::infinity();
          for (const auto child_idx : mcts.gumbel_survivors) {
            const auto& ch = mcts.root.children[child_idx];
            const float logit = std::log(ch.policy + GUMBEL_LOG_FLOOR);
            const float q_hat = ch.n > 0 ? ch.q : 0.0f;
            const float score = mcts.gumbel_g[child_idx] + logit + sigma_scale * q_hat;
            if (score > best_score) {
              best_score = score;
              best = child_idx;
            }
          }
          chosen = mcts.root.children[best].move;
        }
      } else {
        probs_of(mcts, temp, pi);
        chosen = pick_move(pi, re);
      }
      if (c.history_enabled && !game.capped) {
        Sample s;
        c4_canon(game.gs, s.canon);
        for (int k = 0; k < P + 1; ++k) s.v[k] = 0.0f;
        if (c.gumbel_enabled) gumbel_improved_policy(mcts, c, s.pi);  // play_manager.cc:411-417
        else if (c.policy_target_pruning && c.epsilon > 0) probs_pruned_of(mcts, c, 1.0f, s.pi);
        else probs_of(mcts, 1.0f, s.pi);
        game.partial.push_back(s);
      }
      // avg_leaf_depth(): mcts.h:116-119
      const float ald = mcts.depth == 0 ? 0.0f : static_cast<float>(mcts.total_leaf_depth) / static_cast<float>(mcts.depth);
      if (!game.capped) {
        game.total_avg_leaf_depth += ald;
        game.total_search_entropy += root_entropy(mcts);
        ++game.full_move_count;
      } else {
        game.fast_total_avg_leaf_depth += ald;
        game.fast_total_search_entropy += root_entropy(mcts);
        ++game.fast_move_count;
      }
      game.total_valid_moves += mcts.root.children.size();
      ++game.move_count;
      for (auto& m : game.mcts) update_root(m, game.gs, chosen, re);
      c4_play(game.gs, chosen);
      ++pm.moves;
      int term = c4_term(game.gs);
      if (term == 0 && resign_term != 0) term = resign_term;  // :440-444
      else resign_term = 0;
      if (term != 0) {
        if (c.history_enabled) {
          while (!game.partial.empty()) {
            Sample& s = game.partial.back();
            for (int k = 0; k < P + 1; ++k) s.v[k] = (term == k + 1) ? 1.0f : 0.0f;
            pm.history.push_back(s);
            game.partial.pop_back();
          }
        }
        pm.scores[term - 1] += 1.0f;
        if (resign_term != 0) pm.resign_scores[resign_term - 1] += 1.0f;
        ++pm.games_completed;
        pm.game_length += game.gs.turn;
        pm.total_avg_leaf_depth += game.total_avg_leaf_depth;
        pm.total_search_entropy += game.total_search_entropy;
        pm.total_valid_moves += game.total_valid_moves;
        pm.total_move_count += game.move_count;
        pm.full_move_count += game.full_move_count;
        pm.fast_total_avg_leaf_depth += game.fast_total_avg_leaf_depth;
        pm.fast_total_search_entropy += game.fast_total_search_entropy;
        pm.fast_move_count += game.fast_move_count;
        game.total_avg_leaf_depth = game.total_search_entropy = game.total_valid_moves = 0;
        game.fast_total_avg_leaf_depth = game.fast_total_search_entropy = 0;
        game.move_count = game.full_move_count = game.fast_move_count = 0;
        if (pm.games_started >= c.games_to_play) return;  // `continue`: the slot retires (:506-509)
        ++pm.games_started;
        c4_clear(game.gs);
        for (auto& m : game.mcts) m = Tree{};
      }
      {  // :523-524 (`&&` short-circuits: no draw when the flag is off)
        std::uniform_real_distribution<float> dist{0.0F, 1.0F};
        game.capped = c.playout_cap_randomization && (dist(re) < c.playout_cap_percent);
      }
      {  // :531-539
        const int next_cp_g = game.gs.player;
        const uint32_t sims_target = game.capped ? (c.fast_search_uses_gumbel ? c.playout_cap_depth : 0u) : c.mcts_visits[next_cp_g];
        set_gumbel_num_sims(game.mcts[next_cp_g], sims_target);
      }
      if (!c.tree_reuse) {
        for (auto& m : game.mcts) m = Tree{};
      } else {
        Tree& next = game.mcts[game.gs.player];
        if (next.root.n > 0) {
          apply_root_policy_temp(next, c);
          if (c.epsilon > 0 && !game.capped) add_root_noise(next, c, re);
        }
      }
    }
  } else {
    game.initialized = true;
    std::uniform_real_distribution<float> dist{0.0F, 1.0F};
    game.capped = c.playout_cap_randomization && (dist(re) < c.playout_cap_percent);  // :559-560
    {  // :562-570
      const int first_cp = game.gs.player;
      const uint32_t sims_target = game.capped ? (c.fast_search_uses_gumbel ? c.playout_cap_depth : 0u) : c.mcts_visits[first_cp];
      set_gumbel_num_sims(game.mcts[first_cp], sims_target);
    }
  }
  const int cp = game.gs.player;
  Board leaf;
  find_leaf(game.mcts[cp], c, game.gs, leaf, re);
  if (c.eval_type == 1) {
    dumb_eval(leaf, game.v, game.pi);
    pm.awaiting_mcts.push_back(i);
    return;
  }
  c4_canon(leaf, game.canonical);
  pm.awaiting_inference.push_back(i);
}

}  // namespace

extern "C" {

void* azo_pm_new(const azo_cfg* c) {
  if (!c || c->concurrent_games == 0 || c->mcts_visits[0] == 0 || c->mcts_visits[1] == 0) return nullptr;
  auto* pm = new PM;
  static_cast<azo_cfg&>(pm->cfg) = *c;
  pm->games.resize(c->concurrent_games);
  pm->global_rng = std::make_unique<Pcg>(c->seed);
  for (uint32_t g = 0; g < c->concurrent_games; ++g) {  // play_manager.cc:205-255
    c4_clear(pm->games[g].gs);
    pm->games[g].rng = std::make_unique<Pcg>(c->seed + static_cast<uint64_t>(g));  // slot g == a reference thread after seed_thread_rng(seed + g)
    pm->awaiting_mcts.push_back(g);
  }
  pm->games_started = c->concurrent_games;
  return pm;
}
void azo_pm_free(void* h) { delete static_cast<PM*>(h); }
uint32_t azo_pm_run(void* h) {
  auto* pm = static_cast<PM*>(h);
  while (pm->games_completed < pm->cfg.games_to_play && !pm->awaiting_mcts.empty()) {
    const uint32_t i = pm->awaiting_mcts.front();
    pm->awaiting_mcts.pop_front();
    play_iteration(*pm, i);
  }
  return static_cast<uint32_t>(pm->awaiting_inference.size());
}
uint32_t azo_pm_run_iterations(void* h, uint64_t n) {
  auto* pm = static_cast<PM*>(h);
  while (n-- > 0 && pm->games_completed < pm->cfg.games_to_play && !pm->awaiting_mcts.empty()) {
    const uint32_t i = pm->awaiting_mcts.front();
    pm->awaiting_mcts.pop_front();
    play_iteration(*pm, i);
  }
  return static_cast<uint32_t>(pm->awaiting_inference.size());
}
uint32_t azo_pm_build_batch(void* h, uint32_t max, uint32_t* ids, float* canon) {
  auto* pm = static_cast<PM*>(h);
  uint32_t n = 0;
  while (n < max && !pm->awaiting_inference.empty()) {
    const uint32_t i = pm->awaiting_inference.front();
    pm->awaiting_inference.pop_front();
    ids[n] = i;
    std::memcpy(canon + static_cast<size_t>(n) * 168, pm->games[i].canonical, sizeof(float) * 168);
    ++n;
  }
  return n;
}
void azo_pm_update_inferences(void* h, const uint32_t* ids, uint32_t n, const float* v, const float* pi) {
  auto* pm = static_cast<PM*>(h);
  for (uint32_t r = 0; r < n; ++r) {
    Game& g = pm->games[ids[r]];
    std::memcpy(g.v, v + static_cast<size_t>(r) * (P + 1), sizeof(float) * (P + 1));
    std::memcpy(g.pi, pi + static_cast<size_t>(r) * A, sizeof(float) * A);
    pm->awaiting_mcts.push_back(ids[r]);
  }
}
uint32_t azo_pm_drain_history(void* h, uint32_t max, float* canon, float* v, float* pi) {
  auto* pm = static_cast<PM*>(h);
  uint32_t n = 0;
  while (n < max && !pm->history.empty()) {
    const Sample& s = pm->history.front();
    std::memcpy(canon + static_cast<size_t>(n) * 168, s.canon, sizeof(s.canon));
    std::memcpy(v + static_cast<size_t>(n) * (P + 1), s.v, sizeof(s.v));
    std::memcpy(pi + static_cast<size_t>(n) * A, s.pi, sizeof(s.pi));
    pm->history.pop_front();
    ++n;
  }
  return n;
}
uint32_t azo_pm_hist_count(void* h) { return static_cast<uint32_t>(static_cast<PM*>(h)->history.size()); }
uint32_t azo_pm_games_completed(void* h) { return static_cast<PM*>(h)->games_completed; }
uint32_t azo_pm_remaining_games(void* h) {
  auto* pm = static_cast<PM*>(h);
  return pm->cfg.games_to_play - std::min(pm->games_completed, pm->cfg.games_to_play);
}
uint64_t azo_pm_simulations(void* h) { return static_cast<PM*>(h)->simulations; }
uint64_t azo_pm_moves(void* h) { return static_cast<PM*>(h)->moves; }
void azo_pm_scores(void* h, float* out3) { std::memcpy(out3, static_cast<PM*>(h)->scores, sizeof(float) * 3); }
void azo_pm_resign_scores(void* h, float* out3) { std::memcpy(out3, static_cast<PM*>(h)->resign_scores, sizeof(float) * 3); }
void azo_pm_metrics(void* h, float* out7) {  // play_manager.h:288-315
  auto* pm = static_cast<PM*>(h);
  out7[0] = static_cast<float>(pm->game_length) / static_cast<float>(pm->games_completed);
  out7[1] = pm->full_move_count ? static_cast<float>(pm->total_avg_leaf_depth / pm->full_move_count) : 0.0f;
  out7[2] = pm->full_move_count ? static_cast<float>(pm->total_search_entropy / pm->full_move_count) : 0.0f;
  out7[3] = pm->fast_move_count ? static_cast<float>(pm->fast_total_avg_leaf_depth / pm->fast_move_count) : 0.0f;
  out7[4] = pm->fast_move_count ? static_cast<float>(pm->fast_total_search_entropy / pm->fast_move_count) : 0.0f;
  out7[5] = pm->game_length ? static_cast<float>(pm->total_move_count) / static_cast<float>(pm->game_length) : 0.0f;
  out7[6] = pm->total_move_count ? static_cast<float>(pm->total_valid_moves / pm->total_move_count) : 0.0f;
}
void azo_pm_peek(void* h, uint32_t game, uint32_t seat, uint8_t* state89, uint32_t* counts7, float* q7,
                 float* root_value3, uint32_t* depth, uint32_t* root_n, float* policy7) {
  auto* pm = static_cast<PM*>(h);
  const Game& g = pm->games[game];
  const Tree& t = g.mcts[seat];
  if (state89) {  // Connect4GS::to_bytes, connect4_gs.cc:172-178
    std::memcpy(state89, g.gs.b, 84);
    state89[84] = g.gs.player;
    std::memcpy(state89 + 85, &g.gs.turn, 4);
  }
  if (counts7) counts_of(t, counts7);
  if (q7) {
    for (int m = 0; m < A; ++m) q7[m] = 0.0f;
    for (const auto& c : t.root.children) q7[c.move] = c.q;
  }
  if (policy7) {
    for (int m = 0; m < A; ++m) policy7[m] = 0.0f;
    for (const auto& c : t.root.children) policy7[c.move] = c.policy;
  }
  if (root_value3) root_value(t, root_value3);
  if (depth) *depth = t.depth;
  if (root_n) *root_n = t.root.n;
}

int azo_c4_play(int8_t* board84, uint8_t* player, uint32_t* turn, uint32_t move) {
  Board s;
  std::memcpy(s.b, board84, 84);
  s.player = *player;
  s.turn = *turn;
  if (move >= A || !c4_play(s, move)) return -1;
  std::memcpy(board84, s.b, 84);
  *player = s.player;
  *turn = s.turn;
  return 0;
}
void azo_c4_valid(const int8_t* board84, uint8_t* out7) {
  Board s;
  std::memcpy(s.b, board84, 84);
  c4_valid(s, out7);
}
int azo_c4_scores(const int8_t* board84, float* out3) {
  Board s;
  std::memcpy(s.b, board84, 84);
  const int t = c4_term(s);
  for (int i = 0; i < 3; ++i) out3[i] = (t == i + 1) ? 1.0f : 0.0f;
  return t != 0;
}
void azo_c4_canonical(const int8_t* board84, uint8_t player, float* out168) {
  Board s;
  std::memcpy(s.b, board84, 84);
  s.player = player;
  c4_canon(s, out168);
}

void* azo_rng_new(uint64_t seed, int use_stream, uint64_t stream) {
  return use_stream ? new Pcg(seed, stream) : new Pcg(seed);
}
void azo_rng_free(void* r) { delete static_cast<Pcg*>(r); }
uint32_t azo_rng_u32(void* r) { return (*static_cast<Pcg*>(r))(); }
void azo_rng_shuffle(void* r, uint32_t n, uint32_t* inout) { std::shuffle(inout, inout + n, *static_cast<Pcg*>(r)); }
float azo_rng_uniform01(void* r) {
  std::uniform_real_distribution<float> d{0.0f, 1.0f};
  return d(*static_cast<Pcg*>(r));
}
void azo_rng_gamma(void* r, float alpha, uint32_t n, float* out) {
  std::gamma_distribution<float> d{alpha, 1.0f};
  for (uint32_t i = 0; i < n; ++i) out[i] = d(*static_cast<Pcg*>(r));
}

}  // extern "C"
