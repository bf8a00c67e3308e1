// py_tafl_gs.h — host-side view of ONE tafl position for the `alphazero` module: BrandubhGS / OpenTaflGS /
// TawlbwrddGS (py_wrapper.cc:527-558; brandubh_gs.h:98-225, opentafl_gs.h:88-238, tawlbwrdd_gs.h:89-238) on the same
// bitboard rule template the kernels use (az_tafl.h: every function here is the host instantiation of device code).
// Batched rule evaluation on the device is b2az_tafl_replay / b2az_tafl_positions; self-play is b2az_tafl_selfplay_*.
#pragma once

// (az_tafl.h is included by py_alphazero.cc at file scope)

template <int GAME>
class TaflGS : public GameState {
 public:
  typedef b2az::Tafl<GAME> T;
  static constexpr int S = T::S, A = T::A, P = T::PLANES, CELLS = T::CELLS;
  static constexpr uint16_t kDefaultMaxTurns = GAME == B2AZ_TAFL_BRANDUBH ? 150 : 400;  // brandubh_gs.h:29, opentafl_gs.h:18
  b2az::TaflState s;
  std::vector<b2az::TaflKey> hist;  // the repetition table: keys of the positions since the last capture
  uint32_t hist_len = 0;

  explicit TaflGS(uint16_t max_turns = kDefaultMaxTurns) { T::init(s, max_turns); }
  std::unique_ptr<GameState> copy() const override { return std::make_unique<TaflGS<GAME>>(*this); }
  bool equals(const GameState& o) const override {  // brandubh_gs.cc:86-104; OpenTafl also compares the turn (opentafl_gs.cc:82-100)
    auto* c = dynamic_cast<const TaflGS<GAME>*>(&o);
    if (!c) return false;
    if (!T::same(c->s.king, s.king) || !T::same(c->s.def, s.def) || !T::same(c->s.atk, s.atk)) return false;
    if (c->s.player != s.player || c->s.rep != s.rep) return false;
    return GAME != B2AZ_TAFL_OPENTAFL || c->s.turn == s.turn;
  }
  std::string dump() const override {
    std::string out = "Current Player: " + std::to_string((int)s.player) + " Turn: " + std::to_string(s.turn) + '\n';
    for (int h = 0; h < S; ++h) {
      for (int w = 0; w < S; ++w) {
        const int c = T::sq(h, w);
        out += T::test(s.king, c) ? 'K' : T::test(s.def, c) ? 'D' : T::test(s.atk, c) ? 'A' : '.';
      }
      out += '\n';
    }
    return out;
  }
  uint32_t current_turn() const override { return s.turn; }
  uint8_t current_player() const override { return s.player; }
  uint8_t num_players() const override { return 2; }
  uint32_t num_moves() const override { return A; }
  uint8_t num_symmetries() const override { return 8; }

  // tafl_helper::policyLocation (tafl_helper.h:7-14)
  static int loc(int h, int w, bool column_move, int target) { return (h * S + w) * 2 * S + (column_move ? S : 0) + target; }
  // one quarter turn clockwise (tafl_helper.h:55-137): square (h, w) of the image shows square (S-1-w, h) of the base;
  // a row slide of the image is a column slide of the base (target column x <- target row S-1-x) and vice versa
  static PlayHistory rot90(const PlayHistory& b) {
    PlayHistory o = b;
    const ssize_t C = b.dims[0];
    for (int h = 0; h < S; ++h)
      for (int w = 0; w < S; ++w) {
        const int sh = S - 1 - w, sw = h;
        for (ssize_t c = 0; c < C; ++c) o.canonical[(c * S + h) * S + w] = b.canonical[(c * S + sh) * S + sw];
        for (int x = 0; x < S; ++x) {
          o.pi[loc(h, w, false, x)] = b.pi[loc(sh, sw, true, S - 1 - x)];
          o.pi[loc(h, w, true, x)] = b.pi[loc(sh, sw, false, x)];
        }
      }
    return o;
  }
  // left-right mirror (tafl_helper.h:16-53)
  static PlayHistory mirror(const PlayHistory& b) {
    PlayHistory o = b;
    const ssize_t C = b.dims[0];
    for (int h = 0; h < S; ++h)
      for (int w = 0; w < S; ++w) {
        for (ssize_t c = 0; c < C; ++c) o.canonical[(c * S + h) * S + (S - 1 - w)] = b.canonical[(c * S + h) * S + w];
        for (int x = 0; x < S; ++x) {
          o.pi[loc(h, S - 1 - w, false, S - 1 - x)] = b.pi[loc(h, w, false, x)];
          o.pi[loc(h, S - 1 - w, true, x)] = b.pi[loc(h, w, true, x)];
        }
      }
    return o;
  }
  std::vector<PlayHistory> symmetries(const PlayHistory& base) const override {  // eightSym (tafl_helper.h:139-149)
    if (base.dims[1] != S || base.dims[2] != S || (int)base.pi.size() != A) throw std::runtime_error("symmetries: bad shapes");
    std::vector<PlayHistory> out{base};
    for (int i = 0; i < 3; ++i) out.push_back(rot90(out[i]));
    for (int i = 0; i < 4; ++i) out.push_back(mirror(out[i]));
    return out;
  }
  py::array_t<uint8_t> valid_moves() const override {
    py::array_t<uint8_t> a(A);
    uint8_t* d = a.mutable_data();
    for (int c = 0; c < CELLS; ++c) T::valid_bytes(s, c, d + c * 2 * S);
    return a;
  }
  void play_move(uint32_t m) override {
    if (hist.size() < (size_t)hist_len + 2) hist.resize((size_t)hist_len + 64);
    b2az::TaflState t = s;
    uint32_t len = hist_len;
    if (!T::play_hist(t, m, hist.data(), len)) throw std::runtime_error("Invalid move: You have a bug in your code.");
    s = t;
    hist_len = len;
  }
  py::object scores() const override {
    const uint32_t t = T::terminal(s);
    if (!t) return py::none();
    py::array_t<float> a(3);
    for (int i = 0; i < 3; ++i) a.mutable_at(i) = (t == (uint32_t)i + 1u) ? 1.0f : 0.0f;
    return std::move(a);
  }
  py::array_t<float> canonicalized() const override {
    py::array_t<float> a({P, S, S});
    float* d = a.mutable_data();
    for (uint32_t e = 0; e < (uint32_t)T::CANON; ++e) d[e] = T::canon_elem(s, e);
    return a;
  }
  // board int8[3][S][S] + turn u16 + max_turns u16 + player + repetition count + the repetition table
  // (brandubh_gs.cc:18-40: one entry per DISTINCT position with its count; the reference's from_bytes assigns
  // counts[key] = entry_count, so equal keys are merged here — first occurrence order)
  std::string to_bytes() const override {
    std::string out(3 * CELLS, '\0');
    for (uint32_t e = 0; e < (uint32_t)(3 * CELLS); ++e) out[e] = (char)T::board_byte(s, e);
    const uint16_t turn = (uint16_t)s.turn, mt = s.max_turns;
    out.append(reinterpret_cast<const char*>(&turn), 2);
    out.append(reinterpret_cast<const char*>(&mt), 2);
    out.push_back((char)s.player);
    out.push_back((char)s.rep);
    std::vector<uint32_t> first;   // indices of the distinct keys
    std::vector<uint8_t> counts;
    for (uint32_t i = 0; i < hist_len; ++i) {
      size_t j = 0;
      for (; j < first.size(); ++j)
        if (T::key_eq(hist[first[j]], hist[i])) break;
      if (j == first.size()) { first.push_back(i); counts.push_back(1); }
      else if (counts[j] < 255) ++counts[j];
    }
    const uint32_t n = (uint32_t)first.size();
    out.append(reinterpret_cast<const char*>(&n), 4);
    for (uint32_t d = 0; d < n; ++d) {
      const uint32_t i = first[d];
      b2az::TaflState k = s;
      k.king = hist[i].king; k.def = hist[i].def; k.atk = hist[i].atkp;
      const uint8_t kp = (uint8_t)((T::NARROW ? k.atk.lo >> 63 : k.atk.hi >> 63) & 1ULL);
      if (T::NARROW) k.atk.lo &= ~(1ULL << 63); else k.atk.hi &= ~(1ULL << 63);
      for (uint32_t e = 0; e < (uint32_t)(3 * CELLS); ++e) out.push_back((char)T::board_byte(k, e));
      out.push_back((char)kp);
      out.push_back((char)counts[d]);
    }
    return out;
  }
  static void planes_from_bytes(const char* p, b2az::B128& king, b2az::B128& def, b2az::B128& atk) {
    king = def = atk = b2az::b128(0, 0);
    for (int c = 0; c < CELLS; ++c) {
      if (p[c]) king = king | T::bit(c);
      if (p[CELLS + c]) def = def | T::bit(c);
      if (p[2 * CELLS + c]) atk = atk | T::bit(c);
    }
  }
  static TaflGS<GAME> from_bytes(const std::string& data) {
    const size_t B = 3 * CELLS;
    if (data.size() < B + 10) throw std::runtime_error("from_bytes: data too short");
    TaflGS<GAME> g;
    planes_from_bytes(data.data(), g.s.king, g.s.def, g.s.atk);
    uint16_t turn = 0, mt = 0;
    std::memcpy(&turn, &data[B], 2);
    std::memcpy(&mt, &data[B + 2], 2);
    g.s.turn = turn; g.s.max_turns = mt;
    g.s.player = (uint8_t)data[B + 4];
    g.s.rep = (uint8_t)data[B + 5];
    uint32_t n = 0;
    std::memcpy(&n, &data[B + 6], 4);
    if (B + 10 + (size_t)n * (B + 2) != data.size()) throw std::runtime_error("from_bytes: repetition entry count mismatch");
    g.hist.clear();
    for (uint32_t i = 0; i < n; ++i) {
      const char* e = data.data() + B + 10 + (size_t)i * (B + 2);
      b2az::TaflState k = g.s;
      planes_from_bytes(e, k.king, k.def, k.atk);
      k.player = (uint8_t)e[B];
      for (int r = 0; r < (uint8_t)e[B + 1]; ++r) g.hist.push_back(T::key(k));
    }
    g.hist_len = (uint32_t)g.hist.size();
    return g;
  }
  uint64_t hash() const override {  // equality class of brandubh_gs.cc:105-109 / opentafl_gs.cc:102-107 (value is free)
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mix = [&h](uint64_t x) { h = (h ^ x) * 0x100000001b3ULL; h ^= h >> 29; };
    mix(s.king.lo); mix(s.king.hi); mix(s.def.lo); mix(s.def.hi); mix(s.atk.lo); mix(s.atk.hi);
    mix(((uint64_t)s.player << 8) | s.rep);
    if (GAME == B2AZ_TAFL_OPENTAFL) mix(s.turn);
    mix((uint64_t)GAME + 0x9E3779B97F4A7C15ULL);
    return h;
  }
};
typedef TaflGS<B2AZ_TAFL_BRANDUBH> BrandubhGS;
typedef TaflGS<B2AZ_TAFL_OPENTAFL> OpenTaflGS;
typedef TaflGS<B2AZ_TAFL_TAWLBWRDD> TawlbwrddGS;

template <class GS>
void bind_tafl_gs(py::module_& m, const char* name) {
  py::class_<GS, GameState>(m, name)
      .def(py::init<>())
      .def(py::init<uint16_t>())
      .def_static("NUM_PLAYERS", [] { return 2; })
      .def_static("NUM_MOVES", [] { return GS::A; })
      .def_static("NUM_SYMMETRIES", [] { return 8; })
      .def_static("CANONICAL_SHAPE", [] { return std::array<int64_t, 3>{GS::P, GS::S, GS::S}; })
      .def_static("POLICY_SHAPE", [] { return std::array<int, 3>{2 * GS::S, GS::S, GS::S}; })
      .def(py::pickle([](const GS& gs) { return py::bytes(gs.to_bytes()); },
                      [](py::bytes b) { return GS::from_bytes(std::string(b)); }));
}
