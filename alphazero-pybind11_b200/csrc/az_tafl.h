// az_tafl.h — the three tafl games (Brandubh 7x7, OpenTafl 11x11, Tawlbwrdd 11x11) on three bitboards
// (host+device), one rule template instantiated per game.
//
// Replaces the reference's int8[3][S][S] board walks (brandubh_gs.cc:112-482, opentafl_gs.cc:109-506,
// tawlbwrdd_gs.cc:108-398) with one 128-bit set per piece plane: bit (h, w) = S*h + w, planes king / defenders /
// attackers (brandubh_gs.h:31-33). Attackers are player 0 and move first. Reference semantics kept (SURVEY.md
// Appendix B "Tafl common" and the per-game digests):
//   valid_moves  pieces slide like rooks over empty squares. BR/OT: the four corners admit only the king, and a
//                non-king piece may pass the EMPTY throne but not stop on it (brandubh_gs.cc:137-289,
//                opentafl_gs.cc:135-276); TW: no restricted squares (tawlbwrdd_gs.cc:132-213).
//                move id = (h*S + w)*2S + (row slide ? new_w : S + new_h) (tafl_helper.h:7-14)
//   play_move    moves whatever stands on the source square (no legality check), then tests custodial capture in
//                the order up, down, left, right from the destination. Hostile squares — BR/OT: corners to
//                everybody, the throne to attackers always and to defenders only while the king is not on it;
//                TW: enemy pieces only. BR/TW capture the king like any piece; OT: the king is immune on an
//                edge and otherwise needs hostile squares on all four sides — and that test fires for ANY piece
//                moved next to the king, its own side included (opentafl_gs.cc:295-333).
//   repetition   (board, side to move) counts since the last capture; the start position enters the table on
//                the first move of a game (brandubh_gs.cc:349-355, 419-427). Here: a flat history of keys.
//   scores       third repetition: the side to move wins; king on a corner (TW: on any edge): defenders; no king:
//                attackers; OT only: no king/defender reachable from the edge through non-attacker squares:
//                attackers (opentafl_gs.cc:466-506, a flood fill here); side to move without a legal move: the
//                opponent; turn >= max_turns: draw
//   canonical    planes 0-2 board, plane 3 + player all ones, planes 5/6 repetition count 1 -> (1,0), 2 -> (0,1),
//                >= 3 -> (1,1); OT plane 7 = turn / max_turns (opentafl_gs.cc:508-582)
#pragma once

#include "az_common.h"

namespace b2az {

struct B128 {  // a set of board squares
  u64 lo, hi;
};
AZ_HD B128 b128(u64 lo, u64 hi) { B128 r; r.lo = lo; r.hi = hi; return r; }
AZ_HD B128 b128_bit(int i) { return i < 64 ? b128(1ULL << i, 0) : b128(0, 1ULL << (i - 64)); }
AZ_HD bool b128_test(const B128& a, int i) { return ((i < 64 ? a.lo >> i : a.hi >> (i - 64)) & 1ULL) != 0; }
AZ_HD B128 operator|(const B128& a, const B128& b) { return b128(a.lo | b.lo, a.hi | b.hi); }
AZ_HD B128 operator&(const B128& a, const B128& b) { return b128(a.lo & b.lo, a.hi & b.hi); }
AZ_HD B128 operator~(const B128& a) { return b128(~a.lo, ~a.hi); }
AZ_HD bool b128_any(const B128& a) { return (a.lo | a.hi) != 0; }
AZ_HD bool b128_eq(const B128& a, const B128& b) { return a.lo == b.lo && a.hi == b.hi; }
AZ_HD B128 b128_shl(const B128& a, int n) {  // 0 < n < 64
  return b128(a.lo << n, (a.hi << n) | (a.lo >> (64 - n)));
}
AZ_HD B128 b128_shr(const B128& a, int n) {
  return b128((a.lo >> n) | (a.hi << (64 - n)), a.hi >> n);
}

// `nbits` (<= 32) bits starting at bit `off` (0 <= off < 128)
AZ_HD u32 b128_bits(const B128& a, int off, int nbits) {
  const u64 v = off < 64 ? ((a.lo >> off) | (off ? (a.hi << (64 - off)) : 0ULL)) : (a.hi >> (off - 64));
  return (u32)v & (nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u));
}
AZ_HD u32 b128_word(const B128& a, int j) {  // 32-bit chunk j = squares 32j .. 32j+31
  return j == 0 ? (u32)a.lo : j == 1 ? (u32)(a.lo >> 32) : j == 2 ? (u32)a.hi : (u32)(a.hi >> 32);
}

#define B2AZ_TAFL_BRANDUBH 0
#define B2AZ_TAFL_OPENTAFL 1
#define B2AZ_TAFL_TAWLBWRDD 2

template <int GAME>
struct TaflRules;
template <>
struct TaflRules<B2AZ_TAFL_BRANDUBH> {
  static constexpr int S = 7, PLANES = 7;
  static constexpr bool RESTRICTED = true, KING_FOUR_SIDES = false, EDGE_WIN = false, ENCIRCLE = false;
};
template <>
struct TaflRules<B2AZ_TAFL_OPENTAFL> {
  static constexpr int S = 11, PLANES = 8;
  static constexpr bool RESTRICTED = true, KING_FOUR_SIDES = true, EDGE_WIN = false, ENCIRCLE = true;
};
template <>
struct TaflRules<B2AZ_TAFL_TAWLBWRDD> {
  static constexpr int S = 11, PLANES = 7;
  static constexpr bool RESTRICTED = false, KING_FOUR_SIDES = false, EDGE_WIN = true, ENCIRCLE = false;
};

struct TaflState {
  B128 king, def, atk;
  u32 turn;
  u16 max_turns;
  u8 player;  // side to move: 0 attackers, 1 defenders
  u8 rep;     // current_repetition_count_
};
struct TaflKey {  // repetition key: the three planes + the side to move in the top bit (brandubh_gs.h:54-98)
  B128 king, def, atkp;
};

template <int GAME>
struct Tafl {
  typedef TaflRules<GAME> R;
  static constexpr int S = R::S, CELLS = S * S, A = CELLS * 2 * S, PLANES = R::PLANES, CANON = PLANES * CELLS;
  static constexpr int BOARD_BYTES = 3 * CELLS, MID = S / 2, THRONE = MID * S + MID;

  // Boards of up to 64 squares (Brandubh) never touch the high word: single-word forms of the variable-index helpers
  static constexpr bool NARROW = CELLS <= 64;
  static AZ_HD bool test(const B128& a, int i) { return NARROW ? (((a.lo >> i) & 1ULL) != 0) : b128_test(a, i); }
  static AZ_HD B128 bit(int i) { return NARROW ? b128(1ULL << i, 0) : b128_bit(i); }
  static AZ_HD bool any(const B128& a) { return NARROW ? (a.lo != 0) : b128_any(a); }
  static AZ_HD bool same(const B128& a, const B128& b) { return NARROW ? (a.lo == b.lo) : b128_eq(a, b); }
  static AZ_HD int sq(int h, int w) { return S * h + w; }
  static AZ_HD bool is_corner(int h, int w) { return (h == 0 || h == S - 1) && (w == 0 || w == S - 1); }
  static AZ_HD void put(B128& b, int h, int w) { b = b | bit(sq(h, w)); }

  // the constructors of BrandubhGS / OpenTaflGS / TawlbwrddGS (brandubh_gs.h:102-123, opentafl_gs.h:90-134,
  // tawlbwrdd_gs.h:91-134)
  static AZ_HD void init(TaflState& s, u32 max_turns) {
    s.king = s.def = s.atk = b128(0, 0);
    put(s.king, MID, MID);
    if (GAME == B2AZ_TAFL_BRANDUBH) {
      put(s.def, 2, 3); put(s.def, 3, 2); put(s.def, 4, 3); put(s.def, 3, 4);
      put(s.atk, 1, 3); put(s.atk, 0, 3); put(s.atk, 3, 1); put(s.atk, 3, 0);
      put(s.atk, 5, 3); put(s.atk, 6, 3); put(s.atk, 3, 5); put(s.atk, 3, 6);
    } else if (GAME == B2AZ_TAFL_OPENTAFL) {
      put(s.def, 3, 5); put(s.def, 4, 5); put(s.def, 5, 4); put(s.def, 5, 3);
      put(s.def, 6, 5); put(s.def, 7, 5); put(s.def, 5, 6); put(s.def, 5, 7);
      put(s.def, 4, 4); put(s.def, 4, 6); put(s.def, 6, 4); put(s.def, 6, 6);
      for (int t = 3; t <= 7; ++t) { put(s.atk, 0, t); put(s.atk, 10, t); put(s.atk, t, 0); put(s.atk, t, 10); }
      put(s.atk, 1, 5); put(s.atk, 9, 5); put(s.atk, 5, 1); put(s.atk, 5, 9);
    } else {
      for (int t = 2; t <= 4; ++t) { put(s.def, t, 5); put(s.def, 5, t); put(s.def, t + 4, 5); put(s.def, 5, t + 4); }
      for (int t = 4; t <= 6; ++t) {
        put(s.atk, 0, t); put(s.atk, 1, t); put(s.atk, 9, t); put(s.atk, 10, t);
        put(s.atk, t, 0); put(s.atk, t, 1); put(s.atk, t, 9); put(s.atk, t, 10);
      }
    }
    s.turn = 0;
    s.max_turns = (u16)max_turns;
    s.player = 0;
    s.rep = 1;
  }
  static AZ_HD TaflKey key(const TaflState& s) {
    TaflKey k;
    k.king = s.king; k.def = s.def; k.atkp = s.atk;
    if (NARROW) k.atkp.lo |= (u64)s.player << 63;  // 49 squares: bit 63 of the low word is free
    else k.atkp.hi |= (u64)s.player << 63;
    return k;
  }
  static AZ_HD bool key_eq(const TaflKey& a, const TaflKey& b) {
    return same(a.king, b.king) && same(a.def, b.def) && same(a.atkp, b.atkp);
  }
  static AZ_HD B128 own(const TaflState& s) { return s.player == 0 ? s.atk : (s.king | s.def); }

  // Landing squares of the piece on (h, w): bit new_w of `row`, bit new_h of `col` (is_valid_square + the throne
  // exception of the four slide loops).
  static AZ_HD void slides(const TaflState& s, int h, int w, u32& row, u32& col) {
    const B128 occ = s.king | s.def | s.atk;
    const bool is_king = test(s.king, sq(h, w));
    row = col = 0;
    for (int dir = 0; dir < 4; ++dir) {
      const int dh = dir == 2 ? 1 : dir == 3 ? -1 : 0, dw = dir == 0 ? 1 : dir == 1 ? -1 : 0;
      int th = h + dh, tw = w + dw;
      while (th >= 0 && th < S && tw >= 0 && tw < S) {
        // is_valid_square: a corner is decided by the piece alone (king only), any other square must be empty
        if (R::RESTRICTED && is_corner(th, tw)) { if (!is_king) break; }
        else if (test(occ, sq(th, tw))) break;
        if (!(R::RESTRICTED && !is_king && th == MID && tw == MID)) {  // may pass the empty throne, not land on it
          if (dh == 0) row |= 1u << tw; else col |= 1u << th;
        }
        th += dh; tw += dw;
      }
    }
  }
  // The same landing sets from the occupancy of the piece's row and column as S-bit lines (bit i = square i of the
  // line): the free run next to the piece on either side, by count-trailing / count-leading zeros instead of a
  // walk. `row_occ` bit w' = square (h, w'), `col_occ` bit h' = square (h', w).
  static AZ_HD u32 line_reach(u32 occ, int pos) {
    const u32 above = occ >> (pos + 1);
#if defined(__CUDA_ARCH__)
    const int run_up = above ? (__ffs((int)above) - 1) : (S - 1 - pos);
#else
    const int run_up = above ? __builtin_ctz(above) : (S - 1 - pos);
#endif
    const u32 up = ((1u << run_up) - 1u) << (pos + 1);
    const u32 below_mask = (1u << pos) - 1u;
    const u32 low = occ & below_mask;
#if defined(__CUDA_ARCH__)
    const u32 blocked = low ? ((2u << (31 - __clz((int)low))) - 1u) : 0u;
#else
    const u32 blocked = low ? ((2u << (31 - __builtin_clz(low))) - 1u) : 0u;
#endif
    return up | (below_mask & ~blocked);
  }
  static AZ_HD void slides_lines(bool is_king, int h, int w, u32 row_occ, u32 col_occ, u32& row, u32& col) {
    row = line_reach(row_occ, w);
    col = line_reach(col_occ, h);
    if (R::RESTRICTED && !is_king) {
      const u32 ends = 1u | (1u << (S - 1));
      if (h == 0 || h == S - 1) row &= ~ends;   // corners admit only the king
      if (w == 0 || w == S - 1) col &= ~ends;
      if (h == MID) row &= ~(1u << MID);        // the empty throne may be passed, not landed on
      if (w == MID) col &= ~(1u << MID);
    }
  }

  // valid_moves() as the ascending list of legal move ids (the order Node::add_children walks the mask in,
  // mcts.cc:93-101). Returns the count; `out` may be null (count only).
  static AZ_HD u32 moves(const TaflState& s, u16* out) {
    const B128 mine = own(s);
    u32 n = 0;
    for (int c = 0; c < CELLS; ++c) {
      if (!test(mine, c)) continue;
      u32 row, col;
      slides(s, c / S, c % S, row, col);
      for (int t = 0; t < S; ++t)
        if ((row >> t) & 1u) { if (out) out[n] = (u16)(c * 2 * S + t); ++n; }
      for (int t = 0; t < S; ++t)
        if ((col >> t) & 1u) { if (out) out[n] = (u16)(c * 2 * S + S + t); ++n; }
    }
    return n;
  }
  static AZ_HD bool has_moves(const TaflState& s) {
    const B128 mine = own(s);
    for (int c = 0; c < CELLS; ++c) {
      if (!test(mine, c)) continue;
      u32 row, col;
      slides(s, c / S, c % S, row, col);
      if (row | col) return true;
    }
    return false;
  }
  // the 2S mask bytes of one source square (valid_moves()[c*2S .. c*2S + 2S - 1])
  static AZ_HD void valid_bytes(const TaflState& s, int c, u8* out) {
    u32 row = 0, col = 0;
    if (test(own(s), c)) slides(s, c / S, c % S, row, col);
    for (int t = 0; t < S; ++t) {
      out[t] = (u8)((row >> t) & 1u);
      out[S + t] = (u8)((col >> t) & 1u);
    }
  }

  // piece_to_player: 0 attackers, 1 defenders, 2 = empty square (the reference throws)
  static AZ_HD u32 piece_player(const TaflState& s, int c) {
    if (test(s.atk, c)) return 0;
    if (test(s.king | s.def, c)) return 1;
    return 2;
  }
  static AZ_HD bool opponent_piece(const TaflState& s, u32 player, int c) {
    return test(player == 0 ? (s.king | s.def) : s.atk, c);
  }
  // is_hostile_to (brandubh_gs.cc:291-318, opentafl_gs.cc:278-293, tawlbwrdd_gs.cc:215-219)
  static AZ_HD bool hostile_to(const TaflState& s, u32 player, int h, int w) {
    if (R::RESTRICTED) {
      if (is_corner(h, w)) return true;
      if (h == MID && w == MID) return player == 1 ? !test(s.king, THRONE) : true;
    }
    return opponent_piece(s, player, sq(h, w));
  }
  // captured(): 1 if the piece next to `from` in direction (dh, dw) is captured, 0 if not, 2 where the reference
  // throws (empty `from` square). Mask form: the target and the square beyond it as single-bit sets, the squares
  // hostile to the target's side as one set (the mover's pieces, plus corners / throne where the game has them),
  // evaluated on the CURRENT board (a capture made by an earlier direction of the same move is visible, as in the
  // reference's sequential board updates).
  static AZ_HD u32 captured(const TaflState& s, int fh, int fw, int dh, int dw) {
    const int th = fh + dh, tw = fw + dw;
    if (tw < 0 || tw >= S || th < 0 || th >= S) return 0;
    const B128 tgt = bit(sq(th, tw));
    if (R::KING_FOUR_SIDES && any(s.king & tgt)) {
      if (th == 0 || th == S - 1 || tw == 0 || tw == S - 1) return 0;
      return (hostile_to(s, 1, th - 1, tw) && hostile_to(s, 1, th + 1, tw) && hostile_to(s, 1, th, tw - 1) &&
              hostile_to(s, 1, th, tw + 1)) ? 1u : 0u;
    }
    const B128 from = bit(sq(fh, fw));
    const B128 defs = s.king | s.def;
    const u32 from_player = any(s.atk & from) ? 0u : any(defs & from) ? 1u : 2u;
    if (from_player == 2) return 2;
    if (!any((from_player == 0 ? defs : s.atk) & tgt)) return 0;  // only opponent pieces can be captured
    const int lh = th + dh, lw = tw + dw;
    if (lw < 0 || lw >= S || lh < 0 || lh >= S) return 0;
    // squares hostile to the target (the opponent of the mover): the mover's own pieces ...
    B128 hostile = from_player == 0 ? s.atk : defs;
    if (R::RESTRICTED) {  // ... the corners, and the throne — to defenders only while the king is not on it
      hostile = hostile | mask<4>();
      const B128 throne = bit(THRONE);
      if (from_player == 1 /* target: attackers */ || !any(s.king & throne)) hostile = hostile | throne;
    }
    return any(hostile & bit(sq(lh, lw))) ? 1u : 0u;
  }
  // play_move() without the repetition bookkeeping. Returns false where the reference throws.
  static AZ_HD bool play(TaflState& s, u32 move, bool* captured_any) {
    *captured_any = false;
    if (move >= (u32)A) return false;
    u32 new_loc = move % (u32)(2 * S);
    const bool height_move = new_loc >= (u32)S;
    if (height_move) new_loc -= (u32)S;
    const u32 piece_loc = move / (u32)(2 * S);
    const int pw = (int)(piece_loc % (u32)S), ph = (int)(piece_loc / (u32)S);
    const int nh = height_move ? (int)new_loc : ph, nw = height_move ? pw : (int)new_loc;
    const int from = sq(ph, pw), to = sq(nh, nw);
    const B128 fb = bit(from), tb = bit(to);
    // the three layers of the source square are copied onto the destination, then the source is cleared
    const bool k = test(s.king, from), d = test(s.def, from), a = test(s.atk, from);
    s.king = (s.king & ~tb) | (k ? tb : b128(0, 0));
    s.def = (s.def & ~tb) | (d ? tb : b128(0, 0));
    s.atk = (s.atk & ~tb) | (a ? tb : b128(0, 0));
    s.king = s.king & ~fb; s.def = s.def & ~fb; s.atk = s.atk & ~fb;
    for (int i = 0; i < 4; ++i) {
      const int dh = i == 0 ? -1 : i == 1 ? 1 : 0, dw = i == 2 ? -1 : i == 3 ? 1 : 0;
      const u32 c = captured(s, nh, nw, dh, dw);
      if (c == 2) return false;
      if (c == 1) {
        const B128 rm = ~bit(sq(nh + dh, nw + dw));
        s.king = s.king & rm; s.def = s.def & rm; s.atk = s.atk & rm;
        *captured_any = true;
      }
    }
    s.player ^= 1;
    s.turn = (s.turn + 1u) & 0xFFFFu;  // uint16_t turn_
    return true;
  }
  // The whole play_move() including the repetition table, kept as a flat history of keys since the last capture.
  // `hist` must have room for one more key than moves played since the last clear.
  static AZ_HD bool play_hist(TaflState& s, u32 move, TaflKey* hist, u32& hist_len) {
    if (move >= (u32)A) return false;
    if (s.turn == 0) {  // the start position enters the table with the first move
      hist[0] = key(s);
      hist_len = 1;
    }
    bool cap;
    if (!play(s, move, &cap)) return false;
    if (cap) hist_len = 0;
    const TaflKey k = key(s);
    u32 count = 1;
    for (u32 i = 0; i < hist_len; ++i) count += key_eq(hist[i], k) ? 1u : 0u;
    hist[hist_len++] = k;
    s.rep = (u8)(count > 255u ? 255u : count);
    return true;
  }
  // board masks, evaluated at compile time: 0 all squares, 1 edge squares, 2 not in column 0, 3 not in the last
  // column, 4 the four corners
  static AZ_HD constexpr u64 mask_word(int kind, int word) {
    u64 m = 0;
    for (int h = 0; h < S; ++h)
      for (int w = 0; w < S; ++w) {
        const int c = S * h + w;
        const bool edge = h == 0 || h == S - 1 || w == 0 || w == S - 1;
        const bool in = kind == 0 ? true : kind == 1 ? edge : kind == 2 ? (w != 0) : kind == 3 ? (w != S - 1)
                                                                       : ((h == 0 || h == S - 1) && (w == 0 || w == S - 1));
        if (in && (c >> 6) == word) m |= 1ULL << (c & 63);
      }
    return m;
  }
  template <int KIND>
  static AZ_HD B128 mask() {
    constexpr u64 lo = mask_word(KIND, 0), hi = mask_word(KIND, 1);
    return b128(lo, hi);
  }
  // OpenTafl encirclement (opentafl_gs.cc:466-506): flood from every edge square through squares without an
  // attacker; the defenders can escape iff the flood touches a king/defender square.
  static AZ_HD bool can_escape(const TaflState& s) {
    const B128 all = mask<0>(), not_left = mask<2>(), not_right = mask<3>();
    const B128 open = ~s.atk & all;
    const B128 goal = s.king | s.def;
    B128 seen = mask<1>();
    for (;;) {
      if (any(seen & goal)) return true;
      const B128 src = seen & open;  // squares that spread to their neighbours
      const B128 grown = seen | (b128_shl(src, S) & all) | b128_shr(src, S) | b128_shl(src & not_right, 1) |
                         b128_shr(src & not_left, 1);
      if (same(grown, seen)) return false;
      seen = grown;
    }
  }
  // scores() in two halves around the legal-move test (the callers that already know the move count pass it in):
  // 0 = not over, else 1 + index of the winner (2 = defenders, 3 = draw)
  static AZ_HD u32 terminal_pre(const TaflState& s) {
    if (s.rep >= 3) return 1u + s.player;
    if (any(s.king & (R::EDGE_WIN ? mask<1>() : mask<4>()))) return 2;
    if (!any(s.king)) return 1;
    if (R::ENCIRCLE && !can_escape(s)) return 1;
    return 0;
  }
  static AZ_HD u32 terminal_post(const TaflState& s, bool any_move) {
    if (!any_move) return 1u + (s.player ^ 1u);
    if (s.turn >= s.max_turns) return 3;
    return 0;
  }
  static AZ_HD u32 terminal(const TaflState& s) {
    const u32 pre = terminal_pre(s);
    return pre ? pre : terminal_post(s, has_moves(s));
  }
  // canonicalized() element e in [0, CANON): plane = e / CELLS, cell = e % CELLS
  static AZ_HD float canon_elem(const TaflState& s, u32 e) {
    const u32 c = e / (u32)CELLS;
    const int cell = (int)(e % (u32)CELLS);
    if (c == 0) return test(s.king, cell) ? 1.0f : 0.0f;
    if (c == 1) return test(s.def, cell) ? 1.0f : 0.0f;
    if (c == 2) return test(s.atk, cell) ? 1.0f : 0.0f;
    if (c < 5) return (c - 3u == s.player) ? 1.0f : 0.0f;
    if (c == 5) return (s.rep == 1 || s.rep > 2) ? 1.0f : 0.0f;
    if (c == 6) return (s.rep >= 2) ? 1.0f : 0.0f;
    return fdiv((float)s.turn, (float)s.max_turns);  // OpenTafl plane 7
  }
  static AZ_HD signed char board_byte(const TaflState& s, u32 e) {  // int8[3][S][S] element e (to_bytes layout)
    const B128& plane = e < (u32)CELLS ? s.king : e < (u32)(2 * CELLS) ? s.def : s.atk;
    return (signed char)(test(plane, (int)(e % (u32)CELLS)) ? 1 : 0);
  }
};

}  // namespace b2az
