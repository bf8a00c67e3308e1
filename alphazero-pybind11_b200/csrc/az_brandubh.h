// az_brandubh.h — Brandubh (7x7 tafl) rules on three 49-bit bitboards (host+device).
//
// Replaces the reference's int8[3][7][7] board walks (brandubh_gs.cc:112-482) with one u64 per piece
// plane: bit (h, w) = 7*h + w, planes king / defenders / attackers (brandubh_gs.h:31-33). Attackers are
// player 0 and move first. Reference semantics kept (SURVEY.md Appendix B "Tafl common" + "Brandubh"):
//   valid_moves  pieces slide like rooks over empty squares; the four corners admit only the king; a
//                non-king piece may pass the EMPTY throne (3,3) but not stop on it (brandubh_gs.cc:156-289);
//                move id = (h*7 + w)*14 + (row slide ? new_w : 7 + new_h) (tafl_helper.h:7-14)
//   play_move    moves whatever stands on the source square (no legality check), then tests custodial capture
//                in the order up, down, left, right from the destination (brandubh_gs.cc:338-427); the king is
//                captured like any other piece; corners are hostile to everybody, the throne to attackers
//                always and to defenders (king included) only while the king is not on it (:291-318)
//   repetition   (board, side to move) counts since the last capture; the start position enters the table on
//                the first move of a game (:349-355, 419-427). Here the table is a flat history of 24 B keys.
//   scores       third repetition: the side to move wins; king on a corner: defenders; no king: attackers;
//                side to move without a legal move: the opponent; turn >= max_turns: draw (:441-482)
//   canonical    planes 0-2 board, plane 3 + player all ones, planes 5/6 repetition count 1 -> (1,0),
//                2 -> (0,1), >= 3 -> (1,1) (:484-537)
#pragma once

#include "az_common.h"

namespace b2az {

constexpr int kBrS = 7;                 // board side
constexpr int kBrCells = 49;
constexpr int kBrA = 49 * 14;           // 686 actions
constexpr int kBrPlanes = 7;
constexpr int kBrCanon = kBrPlanes * kBrCells;  // 343
constexpr u64 kBrCorners = (1ULL << 0) | (1ULL << 6) | (1ULL << 42) | (1ULL << 48);
constexpr u64 kBrThrone = 1ULL << 24;   // (3, 3)

struct BrState {
  u64 king, def, atk;
  u32 turn;
  u16 max_turns;
  u8 player;      // side to move: 0 attackers, 1 defenders
  u8 rep;         // current_repetition_count_
};
struct BrKey {    // repetition key: the three planes + the side to move (brandubh_gs.h:66-98)
  u64 king, def, atkp;  // atkp = attackers | player << 63
};

AZ_HD u64 br_bit(int h, int w) { return 1ULL << (7 * h + w); }
AZ_HD void br_init(BrState& s, u32 max_turns) {  // BrandubhGS::BrandubhGS (brandubh_gs.h:102-123)
  s.king = br_bit(3, 3);
  s.def = br_bit(2, 3) | br_bit(3, 2) | br_bit(4, 3) | br_bit(3, 4);
  s.atk = br_bit(1, 3) | br_bit(0, 3) | br_bit(3, 1) | br_bit(3, 0) | br_bit(5, 3) | br_bit(6, 3) | br_bit(3, 5) | br_bit(3, 6);
  s.turn = 0;
  s.max_turns = (u16)max_turns;
  s.player = 0;
  s.rep = 1;
}
AZ_HD BrKey br_key(const BrState& s) {
  BrKey k;
  k.king = s.king; k.def = s.def; k.atkp = s.atk | ((u64)s.player << 63);
  return k;
}
AZ_HD bool br_key_eq(const BrKey& a, const BrKey& b) { return a.king == b.king && a.def == b.def && a.atkp == b.atkp; }
AZ_HD u64 br_own(const BrState& s) { return s.player == 0 ? s.atk : (s.king | s.def); }

// Landing squares of the piece on (h, w): bit new_w of `row`, bit new_h of `col` (is_valid_square + the throne
// exception of the four slide loops, brandubh_gs.cc:137-153, 225-283).
AZ_HD void br_slides(const BrState& s, int h, int w, u32& row, u32& col) {
  const u64 occ = s.king | s.def | s.atk;
  const bool is_king = (s.king >> (7 * h + w)) & 1ULL;
  const u64 blocked = occ | (is_king ? 0ULL : kBrCorners);
  const u64 no_land = is_king ? 0ULL : kBrThrone;
  row = col = 0;
  for (int t = w + 1; t < 7; ++t) {
    const u64 b = br_bit(h, t);
    if (blocked & b) break;
    if (!(no_land & b)) row |= 1u << t;
  }
  for (int t = w - 1; t >= 0; --t) {
    const u64 b = br_bit(h, t);
    if (blocked & b) break;
    if (!(no_land & b)) row |= 1u << t;
  }
  for (int t = h + 1; t < 7; ++t) {
    const u64 b = br_bit(t, w);
    if (blocked & b) break;
    if (!(no_land & b)) col |= 1u << t;
  }
  for (int t = h - 1; t >= 0; --t) {
    const u64 b = br_bit(t, w);
    if (blocked & b) break;
    if (!(no_land & b)) col |= 1u << t;
  }
}
// valid_moves() as the ascending list of legal move ids (the order Node::add_children walks the mask in,
// mcts.cc:93-101). Returns the count; `out` may be null (count only).
AZ_HD u32 br_moves(const BrState& s, u16* out) {
  const u64 own = br_own(s);
  u32 n = 0;
  for (int sq = 0; sq < kBrCells; ++sq) {
    if (!((own >> sq) & 1ULL)) continue;
    u32 row, col;
    br_slides(s, sq / 7, sq % 7, row, col);
    for (int t = 0; t < 7; ++t)
      if ((row >> t) & 1u) { if (out) out[n] = (u16)(sq * 14 + t); ++n; }
    for (int t = 0; t < 7; ++t)
      if ((col >> t) & 1u) { if (out) out[n] = (u16)(sq * 14 + 7 + t); ++n; }
  }
  return n;
}
AZ_HD bool br_has_moves(const BrState& s) {  // has_valid_moves (brandubh_gs.cc:156-223)
  const u64 own = br_own(s);
  for (int sq = 0; sq < kBrCells; ++sq) {
    if (!((own >> sq) & 1ULL)) continue;
    u32 row, col;
    br_slides(s, sq / 7, sq % 7, row, col);
    if (row | col) return true;
  }
  return false;
}
// 14 mask bytes of one source square (valid_moves()[sq*14 .. sq*14+13])
AZ_HD void br_valid_bytes(const BrState& s, int sq, u8* out14) {
  u32 row = 0, col = 0;
  if ((br_own(s) >> sq) & 1ULL) br_slides(s, sq / 7, sq % 7, row, col);
  for (int t = 0; t < 7; ++t) {
    out14[t] = (u8)((row >> t) & 1u);
    out14[7 + t] = (u8)((col >> t) & 1u);
  }
}

// piece_to_player (brandubh_gs.cc:112-123): 0 attackers, 1 defenders, 2 = empty square (the reference throws)
AZ_HD u32 br_piece_player(const BrState& s, int sq) {
  if ((s.atk >> sq) & 1ULL) return 0;
  if (((s.king | s.def) >> sq) & 1ULL) return 1;
  return 2;
}
// captured() (brandubh_gs.cc:304-336). Returns 1 if the piece next to `from` in direction (dh, dw) is captured,
// 0 if not, 2 if the reference would have thrown (empty `from` square).
AZ_HD u32 br_captured(const BrState& s, int fh, int fw, int dh, int dw) {
  const int th = fh + dh, tw = fw + dw;
  if (tw < 0 || tw >= 7 || th < 0 || th >= 7) return 0;
  const u32 from_player = br_piece_player(s, 7 * fh + fw);
  if (from_player == 2) return 2;
  const int tsq = 7 * th + tw;
  const u64 opp_of_from = from_player == 0 ? (s.king | s.def) : s.atk;
  if (!((opp_of_from >> tsq) & 1ULL)) return 0;
  const u32 target_player = from_player ^ 1u;  // an opponent piece stands there
  const int lh = th + dh, lw = tw + dw;
  if (lw < 0 || lw >= 7 || lh < 0 || lh >= 7) return 0;
  const u64 fb = br_bit(lh, lw);
  // is_hostile_to (brandubh_gs.cc:291-318)
  if (fb & kBrCorners) return 1;
  if (fb & kBrThrone) return target_player == 1 ? ((s.king & kBrThrone) ? 0u : 1u) : 1u;
  const u64 opp_of_target = target_player == 0 ? (s.king | s.def) : s.atk;
  return (opp_of_target & fb) ? 1u : 0u;
}
// play_move() without the repetition bookkeeping (brandubh_gs.cc:338-417). Returns false where the reference
// throws (move out of range, empty source square); *captured_any tells the caller to clear the history.
AZ_HD bool br_play(BrState& s, u32 move, bool* captured_any) {
  *captured_any = false;
  if (move >= (u32)kBrA) return false;
  u32 new_loc = move % 14u;
  const bool height_move = new_loc >= 7u;
  if (height_move) new_loc -= 7u;
  const u32 piece_loc = move / 14u;
  const int pw = (int)(piece_loc % 7u), ph = (int)(piece_loc / 7u);
  const int nh = height_move ? (int)new_loc : ph, nw = height_move ? pw : (int)new_loc;
  const u64 fb = br_bit(ph, pw), tb = br_bit(nh, nw);
  // the three layers of the source square are copied onto the destination, then the source is cleared
  const u64 k = s.king & fb, d = s.def & fb, a = s.atk & fb;
  s.king = (s.king & ~tb) | (k ? tb : 0ULL);
  s.def = (s.def & ~tb) | (d ? tb : 0ULL);
  s.atk = (s.atk & ~tb) | (a ? tb : 0ULL);
  s.king &= ~fb; s.def &= ~fb; s.atk &= ~fb;
  const int dh[4] = {-1, 1, 0, 0}, dw[4] = {0, 0, -1, 1};
  for (int i = 0; i < 4; ++i) {
    const u32 c = br_captured(s, nh, nw, dh[i], dw[i]);
    if (c == 2) return false;
    if (c == 1) {
      const u64 rm = ~br_bit(nh + dh[i], nw + dw[i]);
      s.king &= rm; s.def &= rm; s.atk &= rm;
      *captured_any = true;
    }
  }
  s.player ^= 1;
  s.turn = (s.turn + 1u) & 0xFFFFu;  // uint16_t turn_
  return true;
}
// The whole play_move() including the repetition table, kept as a flat history of keys since the last capture.
// `hist` must have room for one more key than moves played since the last clear.
AZ_HD bool br_play_hist(BrState& s, u32 move, BrKey* hist, u32& hist_len) {
  if (move >= (u32)kBrA) return false;
  if (s.turn == 0) {  // the start position enters the table with the first move (brandubh_gs.cc:349-355)
    hist[0] = br_key(s);
    hist_len = 1;
  }
  bool cap;
  if (!br_play(s, move, &cap)) return false;
  if (cap) hist_len = 0;
  const BrKey k = br_key(s);
  u32 count = 1;
  for (u32 i = 0; i < hist_len; ++i) count += br_key_eq(hist[i], k) ? 1u : 0u;
  hist[hist_len++] = k;
  s.rep = (u8)(count > 255u ? 255u : count);
  return true;
}
// scores() (brandubh_gs.cc:441-482): 0 = not over, else 1 + index of the winner (2 = defenders, 3 = draw)
AZ_HD u32 br_terminal(const BrState& s) {
  if (s.rep >= 3) return 1u + s.player;
  if (s.king & kBrCorners) return 2;
  if (s.king == 0) return 1;
  if (!br_has_moves(s)) return 1u + (s.player ^ 1u);
  if (s.turn >= s.max_turns) return 3;
  return 0;
}
// canonicalized() element e in [0, 343): plane = e / 49, cell = e % 49 (brandubh_gs.cc:484-537)
AZ_HD float br_canon_elem(const BrState& s, u32 e) {
  const u32 c = e / 49u, cell = e % 49u;
  if (c == 0) return (float)((s.king >> cell) & 1ULL);
  if (c == 1) return (float)((s.def >> cell) & 1ULL);
  if (c == 2) return (float)((s.atk >> cell) & 1ULL);
  if (c < 5) return (c - 3u == s.player) ? 1.0f : 0.0f;
  if (c == 5) return (s.rep == 1 || s.rep > 2) ? 1.0f : 0.0f;
  return (s.rep >= 2) ? 1.0f : 0.0f;
}
// int8[3][7][7] <-> bitboards (the reference's BoardTensor / to_bytes layout, brandubh_gs.cc:11-40)
AZ_HD void br_to_board(const BrState& s, signed char* board147) {
  for (int c = 0; c < kBrCells; ++c) {
    board147[c] = (signed char)((s.king >> c) & 1ULL);
    board147[49 + c] = (signed char)((s.def >> c) & 1ULL);
    board147[98 + c] = (signed char)((s.atk >> c) & 1ULL);
  }
}
AZ_HD void br_from_board(BrState& s, const signed char* board147) {
  s.king = s.def = s.atk = 0;
  for (int c = 0; c < kBrCells; ++c) {
    if (board147[c]) s.king |= 1ULL << c;
    if (board147[49 + c]) s.def |= 1ULL << c;
    if (board147[98 + c]) s.atk |= 1ULL << c;
  }
}

}  // namespace b2az
