// az_connect4.h — Connect4 rules on two 64-bit bitboards (host+device).
//
// Replaces the reference's int8[2][6][7] board (connect4_gs.h:16-18, connect4_gs.cc:39-149) with
// one u64 per player. Bit (h, w) = 7*w + (5 - h): each column owns 7 bits, bit 0 of the column is
// the BOTTOM row (reference row h = 5), bit 6 is a sentinel that stays 0 so shifted-AND win tests
// cannot wrap between columns. Reference semantics kept:
//   valid_moves : column playable <=> its TOP cell (h = 0) is empty        (connect4_gs.cc:39-46)
//   play_move   : stone drops to the lowest empty cell of the column; full column is an error
//                 (connect4_gs.cc:48-58). Arbitrary (non-gravity) boards are legal inputs.
//   scores      : player 0 four-in-a-row first, then player 1, then draw when no column is
//                 playable (connect4_gs.cc:60-129)
//   canonical   : planes 0/1 = absolute player stones, plane 2+player = 1, other = 0, no
//                 perspective flip (connect4_gs.cc:131-149)
#pragma once

#include "az_common.h"

namespace b2az {

struct C4State {
  u64 p[2];
  u32 turn;
  u8 player;
};

#define C4_W 7
#define C4_H 6
#define C4_A 7
#define C4_CANON 168  // 4*6*7

AZ_HD int c4_bit(int h, int w) { return 7 * w + (5 - h); }
AZ_HD u64 c4_top_mask() {  // the top cell (h = 0) of every column
  return (1ULL << 5) | (1ULL << 12) | (1ULL << 19) | (1ULL << 26) | (1ULL << 33) | (1ULL << 40) | (1ULL << 47);
}
AZ_HD void c4_init(C4State& s) { s.p[0] = s.p[1] = 0; s.turn = 0; s.player = 0; }

// bit w set <=> column w playable
AZ_HD u32 c4_valid_mask(const C4State& s) {
  const u64 free_top = ~(s.p[0] | s.p[1]) & c4_top_mask();
  u32 m = 0;
#pragma unroll
  for (int w = 0; w < 7; ++w) m |= (u32)((free_top >> (7 * w + 5)) & 1ULL) << w;
  return m;
}
// returns false (state untouched) when the column has no empty cell
AZ_HD bool c4_play(C4State& s, u32 w) {
  const u64 occ = s.p[0] | s.p[1];
  const u64 free_cells = ~occ & (0x3FULL << (7 * w));
  if (!free_cells) return false;
  const u64 stone = free_cells & (0 - free_cells);  // lowest empty cell
  // selects, not s.p[s.player]: a dynamically indexed member would live in local memory on the device
  s.p[0] |= s.player == 0 ? stone : 0ULL;
  s.p[1] |= s.player == 0 ? 0ULL : stone;
  s.player ^= 1;
  ++s.turn;
  return true;
}
AZ_HD bool c4_has4(u64 m) {
  u64 t = m & (m >> 1);  // vertical
  if (t & (t >> 2)) return true;
  t = m & (m >> 7);      // horizontal
  if (t & (t >> 14)) return true;
  t = m & (m >> 6);      // diagonal
  if (t & (t >> 12)) return true;
  t = m & (m >> 8);      // anti-diagonal
  if (t & (t >> 16)) return true;
  return false;
}
// 0 = game not over, 1 = player 0 won, 2 = player 1 won, 3 = draw (one-hot index + 1)
AZ_HD u32 c4_terminal(const C4State& s) {
  if (c4_has4(s.p[0])) return 1;
  if (c4_has4(s.p[1])) return 2;
  if ((~(s.p[0] | s.p[1]) & c4_top_mask()) == 0) return 3;
  return 0;
}
// canonical element e in [0,168): plane c = e / 42, h = (e % 42) / 7, w = e % 7
AZ_HD float c4_canon_elem(u64 p0, u64 p1, u32 player, u32 e) {
  const u32 c = e / 42u, rem = e % 42u, h = rem / 7u, w = rem % 7u;
  if (c < 2u) return (float)(((c == 0u ? p0 : p1) >> (7u * w + (5u - h))) & 1ULL);
  return (c - 2u == player) ? 1.0f : 0.0f;
}
// index of the canonical element that the mirror image shows at position e (Connect4GS::symmetries,
// connect4_gs.cc:156-162: mirror(f, h, w) = base(f, h, 6 - w))
AZ_HD u32 c4_mirror_elem(u32 e) { return e - (e % 7u) + (6u - e % 7u); }
// int8[2][6][7] (the reference's to_bytes/from-board layout, connect4_gs.cc:172-178) -> bitboards
AZ_HD void c4_from_board(C4State& s, const signed char* board84, int player, int turn) {
  s.p[0] = s.p[1] = 0;
  for (int p = 0; p < 2; ++p)
    for (int h = 0; h < 6; ++h)
      for (int w = 0; w < 7; ++w)
        if (board84[p * 42 + h * 7 + w] != 0) s.p[p] |= 1ULL << c4_bit(h, w);
  s.player = (u8)player;
  s.turn = (u32)turn;
}
AZ_HD void c4_to_board(const C4State& s, signed char* board84) {
  for (int p = 0; p < 2; ++p)
    for (int h = 0; h < 6; ++h)
      for (int w = 0; w < 7; ++w) board84[p * 42 + h * 7 + w] = (signed char)((s.p[p] >> c4_bit(h, w)) & 1ULL);
}
// 64-bit position key for the device cache: equality class = (board, player) like
// Connect4GS::hash (connect4_gs.cc:33-37); the value itself is free (absl hashes are salted).
AZ_HD u64 c4_hash(const C4State& s) {
  u64 x = s.p[0] * 0x9E3779B97F4A7C15ULL ^ (s.p[1] + 0xD1B54A32D192ED03ULL + (u64)s.player);
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL;
  x ^= (s.p[1] << 7) ^ (s.p[0] >> 11);
  x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 32;
  return x;
}

}  // namespace b2az
