// py_stargambit_gs.h — host-side view of ONE Star Gambit position for the `alphazero` module:
// StarGambit{Skirmish,Showdown,Clash,Battle}GS, StarGambitUnifiedGS and its four pinned subclasses
// (py_wrapper.cc:589-695; star_gambit_gs.h:596-911) on the rule header the kernels use (az_stargambit.h: every function
// here is the host instantiation of device code). Batched rule evaluation on the device is b2az_sg_replay.
#pragma once

#include <random>

// (az_stargambit.h is included by py_alphazero.cc at file scope)

struct SGUnitInfo {  // star_gambit_gs.h:424-433
  int player, type, slot, hp, anchor_q, anchor_r, facing, moves_left;
};
struct SGFireInfo {  // star_gambit_gs.h:435-441
  bool has_target;
  int target_player, target_type, target_slot, damage;
};

class StarGambitBase : public GameState {
 public:
  b2az::SGState s;
  std::vector<uint64_t> hist;  // position_history_
  bool unified = false;
  int pinned = -1;                                  // Unified only
  std::array<float, 4> probs{0.25f, 0.25f, 0.25f, 0.25f};

  struct Hist {  // the key history as the rule header wants it, over the std::vector
    std::vector<uint64_t>* v;
    void clear() { v->clear(); }
    int count(uint64_t k) const { int c = 0; for (uint64_t x : *v) c += x == k ? 1 : 0; return c; }
    int push_count(uint64_t k) { v->push_back(k); return count(k); }
  };
  void reset(int variant) {
    b2az::sg_init(s, variant);
    hist.clear();
    hist.push_back(b2az::sg_position_key(s));
  }
  b2az::SGSpace space() const { return b2az::sg_space(s.variant, unified); }

  bool equals(const GameState& o) const override {  // star_gambit_gs.cc:297-319 (+ the variant, 2392-2397)
    auto* c = dynamic_cast<const StarGambitBase*>(&o);
    if (!c || c->unified != unified || c->s.variant != s.variant) return false;
    if (c->s.player != s.player || c->s.n_units != s.n_units || c->s.acted != s.acted) return false;
    if (std::memcmp(c->s.reserves, s.reserves, sizeof(s.reserves)) != 0) return false;
    return std::memcmp(c->s.units, s.units, sizeof(b2az::SGUnit) * s.n_units) == 0;
  }
  uint64_t hash() const override {  // equality class of star_gambit_gs.cc:321-338, 2399-2402 (value is free)
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mix = [&h](uint64_t x) { h = (h ^ x) * 0x100000001b3ULL; h ^= h >> 29; };
    mix(((uint64_t)s.player << 8) | s.acted);
    const uint8_t* u = reinterpret_cast<const uint8_t*>(s.units);
    for (size_t i = 0; i < sizeof(b2az::SGUnit) * s.n_units; ++i) mix(u[i]);
    for (int p = 0; p < 2; ++p)
      for (int t = 0; t < 4; ++t) mix(s.reserves[p][t]);
    mix((uint64_t)s.variant * 2 + (unified ? 1 : 0) + 0x9E3779B97F4A7C15ULL);
    return h;
  }
  std::string dump() const override {  // a plain listing (the reference renders an ASCII hex map, star_gambit_gs.cc:1807-2095)
    std::string out = "Turn: " + std::to_string(s.turn) + ", Player: " + std::to_string((int)s.player) + (s.acted ? " (acted)" : "") + "\n";
    for (int p = 0; p < 2; ++p)
      out += "P" + std::to_string(p) + " reserves: F=" + std::to_string((int)s.reserves[p][0]) + " C=" + std::to_string((int)s.reserves[p][1]) +
             " D=" + std::to_string((int)s.reserves[p][2]) + "\n";
    static const char* names = "FCDP";
    for (int i = 0; i < s.n_units; ++i) {
      const auto& u = s.units[i];
      if (!u.hp) continue;
      out += std::string("P") + std::to_string((int)u.player) + " " + names[u.type & 3] + std::to_string(u.slot + 1) + " hp=" + std::to_string((int)u.hp) +
             " at (" + std::to_string((int)u.q) + "," + std::to_string((int)u.r) + ") facing " + std::to_string((int)u.facing) + "\n";
    }
    return out;
  }
  uint32_t current_turn() const override { return s.turn; }
  uint8_t current_player() const override { return s.player; }
  uint8_t num_players() const override { return 2; }
  uint32_t num_moves() const override { return (uint32_t)space().num_moves(); }
  uint8_t num_symmetries() const override { return 2; }
  bool relative_values() const override { return true; }

  py::array_t<uint8_t> valid_moves() const override {
    const auto sp = space();
    py::array_t<uint8_t> a(sp.num_moves());
    uint8_t* d = a.mutable_data();
    std::memset(d, 0, (size_t)sp.num_moves());
    b2az::sg_valid_moves(s, sp, [d](int id) { d[id] = 1; });
    return a;
  }
  void play_move(uint32_t m) override {
    Hist h{&hist};
    b2az::sg_play(s, h, space(), m);
  }
  py::object scores() const override {
    const uint32_t t = b2az::sg_terminal(s);
    if (!t) return py::none();
    py::array_t<float> a(3);
    for (int i = 0; i < 3; ++i) a.mutable_at(i) = (t == (uint32_t)i + 1u) ? 1.0f : 0.0f;
    return std::move(a);
  }
  py::array_t<float> canonicalized() const override {
    const auto sp = space();
    const int P = sp.planes(unified), D = sp.udim;
    py::array_t<float> a({P, D, D});
    float* d = a.mutable_data();
    b2az::SGCanonCtx ctx;
    Hist h{const_cast<std::vector<uint64_t>*>(&hist)};
    b2az::sg_canon_ctx(s, h, sp, ctx);
    for (int ch = 0; ch < P; ++ch)
      for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) d[(ch * D + r) * D + c] = b2az::sg_canon_elem(s, ctx, sp, ch, r, c);
    return a;
  }
  std::vector<PlayHistory> symmetries(const PlayHistory& base) const override {  // star_gambit_gs.cc:1671-1805, 2623-2727
    const auto sp = space();
    const int P = sp.planes(unified), D = sp.udim, side = (D - 1) / 2, A = sp.num_moves();
    if (base.dims[0] != P || base.dims[1] != D || base.dims[2] != D || (int)base.pi.size() != A) throw std::runtime_error("symmetries: bad shapes");
    PlayHistory m = base;
    for (int i = 0; i < P * D * D; ++i) {
      const int src = b2az::sg_mirror_canon_src(D, side, P, i);
      m.canonical[i] = src < 0 ? 0.0f : base.canonical[src];
    }
    for (int i = 0; i < A; ++i) {
      const int src = b2az::sg_mirror_pi_src(D, side, i);
      m.pi[i] = src < 0 ? 0.0f : base.pi[src];
    }
    return {base, m};
  }
  std::string inner_bytes() const {  // star_gambit_gs.cc:2253-2288
    std::string out;
    auto app = [&out](const void* p, size_t n) { out.append(static_cast<const char*>(p), n); };
    const uint32_t n = s.n_units;
    app(&n, 4);
    app(s.units, sizeof(b2az::SGUnit) * n);
    app(s.reserves, 8);
    out.push_back((char)s.player);
    app(&s.turn, 4);
    out.push_back((char)(s.acted ? 1 : 0));
    out.push_back((char)(s.over ? 1 : 0));
    out.push_back((char)s.winner);
    const uint32_t hl = (uint32_t)hist.size();
    app(&hl, 4);
    if (hl) app(hist.data(), (size_t)hl * 8);
    return out;
  }
  void load_inner(const std::string& data, int variant) {  // star_gambit_gs.cc:2290-2338
    size_t off = 0;
    auto rd = [&](void* p, size_t n) {
      if (off + n > data.size()) throw std::runtime_error("StarGambitGS::from_bytes: short data");
      std::memcpy(p, &data[off], n);
      off += n;
    };
    b2az::sg_init(s, variant);
    uint32_t n = 0;
    rd(&n, 4);
    if (n > (uint32_t)b2az::kSGMaxUnits) throw std::runtime_error("StarGambitGS::from_bytes: too many units");
    s.n_units = (uint8_t)n;
    rd(s.units, sizeof(b2az::SGUnit) * n);
    rd(s.reserves, 8);
    uint8_t b = 0;
    rd(&b, 1); s.player = b;
    rd(&s.turn, 4);
    rd(&b, 1); s.acted = b != 0;
    rd(&b, 1); s.over = b != 0;
    rd(&b, 1); s.winner = (int8_t)b;
    uint32_t hl = 0;
    rd(&hl, 4);
    if (off + (size_t)hl * 8 > data.size()) throw std::runtime_error("StarGambitGS::from_bytes: short data");
    hist.resize(hl);
    if (hl) rd(hist.data(), (size_t)hl * 8);
    if (off != data.size()) throw std::runtime_error("StarGambitGS::from_bytes: trailing bytes");
  }
  std::vector<SGUnitInfo> get_units() const {  // star_gambit_gs.cc:2101-2119
    std::vector<SGUnitInfo> out;
    for (int i = 0; i < s.n_units; ++i) {
      const auto& u = s.units[i];
      if (!u.hp) continue;
      out.push_back({u.player, u.type, u.slot, u.hp, u.q, u.r, u.facing, u.moves_left});
    }
    return out;
  }
  SGFireInfo get_fire_info(uint32_t move) const {  // star_gambit_gs.cc:2121-2240: the ENEMY a fire action would hit
    SGFireInfo r{false, -1, -1, -1, 0};
    const auto sp = space();
    if (move >= (uint32_t)sp.deploy_offset()) return r;
    const int slot = (int)(move % 10u), pos = (int)(move / 10u);
    int row = pos / sp.udim - sp.off, col = pos % sp.udim - sp.off;
    if (slot < 5 || row < 0 || row >= sp.dim || col < 0 || col >= sp.dim) return r;
    if (s.player == 1) { row = sp.dim - 1 - row; col = sp.dim - 1 - col; }
    const int q = row - sp.side, rr = col - sp.side;
    int ui = -1;
    for (int i = 0; i < s.n_units && ui < 0; ++i) {
      const auto& u = s.units[i];
      if (u.player == s.player && u.hp > 0 && u.q == q && u.r == rr && u.type != b2az::SG_PORTAL) ui = i;
    }
    if (ui < 0) return r;
    const auto& u = s.units[ui];
    const int ci = b2az::sg_slot_index(u.type, slot);
    if (ci < 0) return r;
    int doff, src, hq[3], hr[3];
    b2az::sg_cannon(u.type, ci, doff, src);
    b2az::sg_unit_hexes(u, sp.side, hq, hr);
    const int d = b2az::sg_rot(u.facing, doff);
    int tq = hq[src], tr = hr[src];
    for (int range = 1; range <= 2; ++range) {
      tq += b2az::sg_dq(d); tr += b2az::sg_dr(d);
      if (!b2az::sg_inb(tq, tr, sp.side)) continue;
      if (range == 2 && b2az::sg_unit_at(s, tq - b2az::sg_dq(d), tr - b2az::sg_dr(d), sp.side) >= 0) break;
      const int t = b2az::sg_unit_at(s, tq, tr, sp.side);
      if (t >= 0 && s.units[t].player != u.player) {
        const auto& tu = s.units[t];
        return {true, tu.player, tu.type, tu.slot, range == 1 ? 2 : 1};
      }
    }
    return r;
  }
};

template <int VARIANT>
class StarGambitGS : public StarGambitBase {
 public:
  static constexpr int SIDE = VARIANT == B2AZ_SG_BATTLE ? 6 : 5, D = 2 * SIDE + 1, A = D * D * 10 + 19;
  StarGambitGS() { reset(VARIANT); }
  std::unique_ptr<GameState> copy() const override { return std::make_unique<StarGambitGS<VARIANT>>(*this); }
  std::string to_bytes() const override { return inner_bytes(); }
  static StarGambitGS<VARIANT> from_bytes(const std::string& data) {
    StarGambitGS<VARIANT> g;
    g.load_inner(data, VARIANT);
    return g;
  }
};

class StarGambitUnifiedGS : public StarGambitBase {  // star_gambit_gs.h:788-887
 public:
  explicit StarGambitUnifiedGS(int pinned_variant = -1, std::array<float, 4> p = {0.25f, 0.25f, 0.25f, 0.25f}) {
    unified = true;
    pinned = pinned_variant;
    probs = p;
    reset(pick());
  }
  int pick() const {  // pick_variant / pick_from_probs (star_gambit_gs.cc:2357-2362, 2404-2409): an unseedable mt19937
    if (pinned >= 0 && pinned <= 3) return pinned;
    thread_local std::mt19937 rng{std::random_device{}()};
    std::discrete_distribution<int> dist(probs.begin(), probs.end());
    return dist(rng);
  }
  std::unique_ptr<GameState> copy() const override { return std::make_unique<StarGambitUnifiedGS>(*this); }
  void randomize_start() override { reset(pick()); }
  int num_variants() const override { return 4; }
  int get_variant_id() const override { return s.variant; }
  std::string to_bytes() const override {  // star_gambit_gs.cc:2451-2465
    std::string out(reinterpret_cast<const char*>(probs.data()), 16);
    const int32_t pv = pinned;
    out.append(reinterpret_cast<const char*>(&pv), 4);
    out.push_back((char)s.variant);
    const std::string inner = inner_bytes();
    const uint32_t n = (uint32_t)inner.size();
    out.append(reinterpret_cast<const char*>(&n), 4);
    return out + inner;
  }
  static StarGambitUnifiedGS from_bytes(const std::string& data) {  // star_gambit_gs.cc:2467-2516
    if (data.size() < 25) throw std::runtime_error("StarGambitUnifiedGS::from_bytes: short data");
    std::array<float, 4> p{};
    std::memcpy(p.data(), data.data(), 16);
    int32_t pv = 0;
    std::memcpy(&pv, &data[16], 4);
    const uint8_t variant = (uint8_t)data[20];
    if (variant > 3) throw std::runtime_error("StarGambitUnifiedGS::from_bytes: bad variant_id");
    uint32_t n = 0;
    std::memcpy(&n, &data[21], 4);
    if (25 + (size_t)n > data.size()) throw std::runtime_error("StarGambitUnifiedGS::from_bytes: short inner");
    if (25 + (size_t)n != data.size()) throw std::runtime_error("StarGambitUnifiedGS::from_bytes: trailing bytes");
    StarGambitUnifiedGS g(pv, p);
    g.load_inner(data.substr(25, n), variant);
    return g;
  }
};
template <int VARIANT>
class StarGambitUnifiedPinnedGS : public StarGambitUnifiedGS {
 public:
  StarGambitUnifiedPinnedGS() : StarGambitUnifiedGS(VARIANT) {}
};

template <int VARIANT>
void bind_sg_gs(py::module_& m, const char* name) {
  typedef StarGambitGS<VARIANT> GS;
  py::class_<GS, GameState>(m, name)
      .def(py::init<>())
      .def("get_units", &GS::get_units)
      .def("get_fire_info", &GS::get_fire_info)
      .def_static("NUM_PLAYERS", [] { return 2; })
      .def_static("NUM_MOVES", [] { return GS::A; })
      .def_static("NUM_SYMMETRIES", [] { return 2; })
      .def_static("CANONICAL_SHAPE", [] { return std::array<int, 3>{b2az::kSGPlanes, GS::D, GS::D}; })
      .def_static("POLICY_SHAPE", [] { return std::array<int, 3>{10, GS::D, GS::D}; })
      .def(py::pickle([](const GS& gs) { return py::bytes(gs.to_bytes()); },
                      [](py::bytes b) { return GS::from_bytes(std::string(b)); }));
}
template <int VARIANT>
void bind_sg_pinned(py::module_& m, const char* name) {  // (pickles as its own type: the state is the Unified layout)
  typedef StarGambitUnifiedPinnedGS<VARIANT> GS;
  py::class_<GS, StarGambitUnifiedGS>(m, name)
      .def(py::init<>())
      .def(py::pickle([](const GS& gs) { return py::bytes(gs.to_bytes()); },
                      [](py::bytes b) {
                        GS g;
                        static_cast<StarGambitBase&>(g) = StarGambitUnifiedGS::from_bytes(std::string(b));
                        return g;
                      }));
}
inline void bind_star_gambit(py::module_& m) {
  py::class_<SGUnitInfo>(m, "UnitInfo")
      .def_readonly("player", &SGUnitInfo::player).def_readonly("type", &SGUnitInfo::type)
      .def_readonly("slot", &SGUnitInfo::slot).def_readonly("hp", &SGUnitInfo::hp)
      .def_readonly("anchor_q", &SGUnitInfo::anchor_q).def_readonly("anchor_r", &SGUnitInfo::anchor_r)
      .def_readonly("facing", &SGUnitInfo::facing).def_readonly("moves_left", &SGUnitInfo::moves_left);
  py::class_<SGFireInfo>(m, "FireInfo")
      .def_readonly("has_target", &SGFireInfo::has_target).def_readonly("target_player", &SGFireInfo::target_player)
      .def_readonly("target_type", &SGFireInfo::target_type).def_readonly("target_slot", &SGFireInfo::target_slot)
      .def_readonly("damage", &SGFireInfo::damage);
  bind_sg_gs<B2AZ_SG_SKIRMISH>(m, "StarGambitSkirmishGS");
  bind_sg_gs<B2AZ_SG_SHOWDOWN>(m, "StarGambitShowdownGS");
  bind_sg_gs<B2AZ_SG_CLASH>(m, "StarGambitClashGS");
  bind_sg_gs<B2AZ_SG_BATTLE>(m, "StarGambitBattleGS");
  py::class_<StarGambitUnifiedGS, GameState>(m, "StarGambitUnifiedGS")
      .def(py::init<int, std::array<float, 4>>(), py::arg("pinned_variant") = -1,
           py::arg("probs") = std::array<float, 4>{0.25f, 0.25f, 0.25f, 0.25f})
      .def("get_units", &StarGambitUnifiedGS::get_units)
      .def("get_fire_info", &StarGambitUnifiedGS::get_fire_info)
      .def_static("NUM_PLAYERS", [] { return 2; })
      .def_static("NUM_MOVES", [] { return b2az::kSGUnifiedMoves; })
      .def_static("NUM_SYMMETRIES", [] { return 2; })
      .def_static("CANONICAL_SHAPE", [] { return std::array<int, 3>{b2az::kSGUnifiedPlanes, 13, 13}; })
      .def_static("POLICY_SHAPE", [] { return std::array<int, 3>{10, 13, 13}; })
      .def(py::pickle([](const StarGambitUnifiedGS& gs) { return py::bytes(gs.to_bytes()); },
                      [](py::bytes b) { return StarGambitUnifiedGS::from_bytes(std::string(b)); }));
  bind_sg_pinned<0>(m, "StarGambitUnifiedSkirmishGS");
  bind_sg_pinned<1>(m, "StarGambitUnifiedShowdownGS");
  bind_sg_pinned<2>(m, "StarGambitUnifiedClashGS");
  bind_sg_pinned<3>(m, "StarGambitUnifiedBattleGS");
}
