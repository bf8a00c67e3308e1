// az_stargambit.h — Star Gambit (four variants + the Unified 13x13 view) on a unit list with hex bitboards
// (host+device), one rule set parameterised by a small config record.
//
// Replaces the reference's per-candidate std::vector walks (star_gambit_gs.cc:372-442: every move / fire / deploy test
// rebuilds the list of occupied hexes and scans it) with three 169-bit occupancy sets over the 13x13 axial grid
// (bit (q + 6) * 13 + (r + 6)): all alive units, and one per player. A candidate action is then a handful of bit
// tests; every unit's ten action slots are independent of the other units' (one lane per unit on the device).
// Reference semantics kept (SURVEY.md Appendix B "Star Gambit"):
//   units        {type, player, slot, hp, facing, anchor q, r, moves_left, cannons_fired}, 9 bytes, in creation order
//                (the two portals first); destroyed units stay in the list with hp 0 (star_gambit_gs.h:359-371)
//   shapes       fighter 1 hex; cruiser anchor = front + the hex behind it; dreadnought anchor + rear-right + rear
//                (star_gambit_gs.cc:88-120); portal 3 fixed hexes per player (122-141)
//   valid_moves  (784-923) nothing when the game is over; turns 1 and 2 are deploy-only; per own alive ship: moves
//                while moves_left > 0 (new hexes in bounds and free of OTHER units), one fire slot per unfired cannon
//                with an enemy as the first thing in range 1-2; deploys (reserve left, facing allowed for the seat,
//                hexes free); END_TURN after at least one action. Player 1 acts in the frame rotated by 180 degrees
//                (row, col mirrored, deploy facing + 3; the facing-relative slot is not remapped).
//   play_move    (1093-1238) the ship is found by its anchor; a move is executed when it stays in bounds (no
//                collision test, moves_left decremented modulo 256 — only legal moves are ever played by the search);
//                a fire marks the cannon, hits the first unit in range 1-2 of ANY side but itself (damage 2 / 1), a
//                destroyed unit triggers check_game_end (1313-1345); every spatial action ends with the repetition
//                check (1246-1261: the position key is appended, three equal keys = draw); a deploy clears the key
//                history and ends the turn; END_TURN (1263-1290): switch player, ++turn, turn > 200 = draw,
//                repetition check, refresh the mover's ships, no legal action = the mover loses.
//   position key (1365-1382) XOR of per-unit products; excludes moves_left / cannons_fired / reserves
//   canonical    (1384-1669) 32 planes in the mover's frame; Unified (2586-2616): small boards sit in the 13x13 canvas
//                with a one-cell border, planes 32-35 = variant one-hot on the board's hexes
//   symmetries   (1671-1805, 2623-2727) identity + mirror about the NW axis
#pragma once

#include "az_common.h"

namespace b2az {

#define B2AZ_SG_SKIRMISH 0
#define B2AZ_SG_SHOWDOWN 1
#define B2AZ_SG_CLASH 2
#define B2AZ_SG_BATTLE 3

constexpr int kSGMaxUnits = 20;     // Battle: 2 x (4 fighters + 3 cruisers + 2 dreadnoughts + portal)
constexpr int kSGMaxTurns = 200;
constexpr int kSGUnifiedDim = 13, kSGUnifiedMoves = 13 * 13 * 10 + 18 + 1;  // 1709
constexpr int kSGPlanes = 32, kSGUnifiedPlanes = 36;
enum { SG_FIGHTER = 0, SG_CRUISER = 1, SG_DREAD = 2, SG_PORTAL = 3 };

struct SGUnit {  // star_gambit_gs.h:359-371 (the serialised layout of to_bytes as well)
  u8 type, player, slot, hp, facing;
  int8_t q, r;
  u8 moves_left, fired;
};
struct SGState {
  SGUnit units[kSGMaxUnits];
  u8 n_units;
  u8 reserves[2][4];
  u8 player, acted, over;
  int8_t winner;  // -1 none, 0 / 1, 2 = draw
  u8 variant;     // B2AZ_SG_*
  u8 pad_[2];
  u32 turn;
};

// the action / observation frame a state is presented in: its own (2*side+1)^2 grid, or the Unified 13x13 canvas
struct SGSpace {
  int side, dim;   // the variant's own board
  int udim, off;   // target grid and the offset of the board inside it
  AZ_HD int deploy_offset() const { return udim * udim * 10; }
  AZ_HD int end_turn() const { return udim * udim * 10 + 18; }
  AZ_HD int num_moves() const { return udim * udim * 10 + 19; }
  AZ_HD int planes(bool unified) const { return unified ? kSGUnifiedPlanes : kSGPlanes; }
};
AZ_HD int sg_side(int variant) { return variant == B2AZ_SG_BATTLE ? 6 : 5; }
AZ_HD int sg_start(int variant, int type) {  // STARTING_FIGHTERS / CRUISERS / DREADNOUGHTS (star_gambit_gs.h:22-60)
  // packed nibbles: type 0 | type 1 << 4 | type 2 << 8
  const u32 t = variant == B2AZ_SG_SKIRMISH ? 0x013u : variant == B2AZ_SG_SHOWDOWN ? 0x104u
              : variant == B2AZ_SG_CLASH ? 0x123u : 0x234u;
  return type < 3 ? (int)((t >> (4 * type)) & 15u) : 0;
}
AZ_HD SGSpace sg_space(int variant, bool unified) {
  SGSpace sp;
  sp.side = sg_side(variant);
  sp.dim = 2 * sp.side + 1;
  sp.udim = unified ? kSGUnifiedDim : sp.dim;
  sp.off = (sp.udim - sp.dim) / 2;
  return sp;
}

// hex directions E, NE, NW, W, SW, SE (star_gambit_gs.h:251-258)
AZ_HD int sg_dq(int d) { return (int)((0x41Au >> (2 * d)) & 3u) - 1; }
AZ_HD int sg_dr(int d) { return (int)((0xA41u >> (2 * d)) & 3u) - 1; }
AZ_HD int sg_rot(int d, int steps) { return (d + steps + 6) % 6; }
AZ_HD bool sg_inb(int q, int r, int side) {
  const int s = -q - r;
  return q >= -side && q <= side && r >= -side && r <= side && s >= -side && s <= side;
}
AZ_HD int sg_max_hp(int type) { return type == SG_FIGHTER ? 3 : type == SG_CRUISER ? 4 : type == SG_DREAD ? 6 : 5; }
AZ_HD int sg_max_moves(int type) { return type == SG_FIGHTER ? 2 : type == SG_PORTAL ? 0 : 1; }
AZ_HD int sg_num_cannons(int type) { return type == SG_FIGHTER ? 1 : type == SG_CRUISER ? 3 : type == SG_DREAD ? 4 : 0; }

struct SGOcc {  // a set of hexes of the 13x13 axial grid
  u64 w[3];
};
AZ_HD int sg_cell(int q, int r) { return (q + 6) * 13 + (r + 6); }
AZ_HD void sg_occ_clear(SGOcc& o) { o.w[0] = o.w[1] = o.w[2] = 0; }
AZ_HD void sg_occ_set(SGOcc& o, int q, int r) {  // callers pass hexes of placed units (in bounds by construction)
  if (q < -6 || q > 6 || r < -6 || r > 6) return;
  const int c = sg_cell(q, r);
  o.w[c >> 6] |= 1ULL << (c & 63);
}
AZ_HD bool sg_occ_test(const SGOcc& o, int q, int r) {
  if (q < -6 || q > 6 || r < -6 || r > 6) return false;
  const int c = sg_cell(q, r);
  return ((o.w[c >> 6] >> (c & 63)) & 1ULL) != 0;
}

// the hexes of a unit (get_unit_hexes / get_portal_hexes, star_gambit_gs.cc:88-141); returns their number
AZ_HD int sg_hexes(int type, int player, int q, int r, int facing, int side, int* hq, int* hr) {
  if (type == SG_PORTAL) {
    const int sg = player == 0 ? 1 : -1;  // player 0 at the bottom (positive r)
    hq[0] = 0; hr[0] = sg * side;
    hq[1] = sg; hr[1] = sg * (side - 1);
    hq[2] = -sg; hr[2] = sg * side;
    return 3;
  }
  hq[0] = q; hr[0] = r;
  if (type == SG_FIGHTER) return 1;
  const int rear = (facing + 3) % 6;
  if (type == SG_CRUISER) {
    hq[1] = q + sg_dq(rear); hr[1] = r + sg_dr(rear);
    return 2;
  }
  const int rr = sg_rot(rear, 1);
  hq[1] = q + sg_dq(rr); hr[1] = r + sg_dr(rr);
  hq[2] = q + sg_dq(rear); hr[2] = r + sg_dr(rear);
  return 3;
}
AZ_HD int sg_unit_hexes(const SGUnit& u, int side, int* hq, int* hr) {
  return sg_hexes(u.type, u.player, u.q, u.r, u.facing, side, hq, hr);
}

struct SGBoards {  // occupancy of the alive units: all, and per player
  SGOcc all, pl[2];
};
AZ_HD_CALL void sg_boards(const SGState& s, int side, SGBoards& b) {
  sg_occ_clear(b.all); sg_occ_clear(b.pl[0]); sg_occ_clear(b.pl[1]);
  for (int i = 0; i < (int)s.n_units; ++i) {
    const SGUnit& u = s.units[i];
    if (u.hp == 0) continue;
    int hq[3], hr[3];
    const int n = sg_unit_hexes(u, side, hq, hr);
    for (int j = 0; j < n; ++j) {
      sg_occ_set(b.all, hq[j], hr[j]);
      sg_occ_set(b.pl[u.player & 1], hq[j], hr[j]);
    }
  }
}

// compute_{fighter,cruiser,dreadnought}_move (star_gambit_gs.cc:448-600): `dir` is the unit type's own move index.
// Returns false when a hex of the new placement leaves the board.
AZ_HD_CALL bool sg_compute_move(const SGUnit& u, int dir, int side, int& nq, int& nr, int& nf) {
  const int f = u.facing, q = u.q, r = u.r;
  if (u.type == SG_FIGHTER) {  // 0 forward, 1 forward-left, 2 forward-right; faces where it moves
    if (dir < 0 || dir > 2) return false;
    nf = dir == 0 ? f : dir == 1 ? sg_rot(f, 1) : sg_rot(f, -1);
    nq = q + sg_dq(nf); nr = r + sg_dr(nf);
    return sg_inb(nq, nr, side);
  }
  if (u.type == SG_CRUISER) {  // 0 rotate-left, 1 forward-left, 2 forward, 3 forward-right, 4 rotate-right
    if (dir < 0 || dir > 4) return false;
    if (dir == 0 || dir == 4) {  // the rear hex stays, the front pivots
      const int rear = (f + 3) % 6, rq = q + sg_dq(rear), rr = r + sg_dr(rear);
      nf = sg_rot(f, dir == 0 ? 1 : -1);
      nq = rq + sg_dq(nf); nr = rr + sg_dr(nf);
    } else {
      nf = dir == 1 ? sg_rot(f, 1) : dir == 2 ? f : sg_rot(f, -1);
      nq = q + sg_dq(nf); nr = r + sg_dr(nf);
    }
  } else if (u.type == SG_DREAD) {  // 0 pivot left, 1 slide forward-left, 2 slide forward-right, 3 pivot right
    if (dir < 0 || dir > 3) return false;
    const int rear = (f + 3) % 6;
    if (dir == 0) {  // about the rear hex
      const int pq = q + sg_dq(rear), pr = r + sg_dr(rear), nd = sg_rot((rear + 3) % 6, 1);
      nq = pq + sg_dq(nd); nr = pr + sg_dr(nd);
      nf = sg_rot(f, 1);
    } else if (dir == 1) {
      const int d = sg_rot(f, 1);
      nq = q + sg_dq(d); nr = r + sg_dr(d);
      nf = f;
    } else if (dir == 2) {
      nq = q + sg_dq(f); nr = r + sg_dr(f);
      nf = f;
    } else {  // about the rear-right hex
      const int rrd = sg_rot(rear, 1), pq = q + sg_dq(rrd), pr = r + sg_dr(rrd), nd = sg_rot((rrd + 3) % 6, -1);
      nq = pq + sg_dq(nd); nr = pr + sg_dr(nd);
      nf = sg_rot(f, -1);
    }
  } else {
    return false;
  }
  int hq[3], hr[3];
  const int n = sg_hexes(u.type, u.player, nq, nr, nf, side, hq, hr);
  for (int j = 0; j < n; ++j)
    if (!sg_inb(hq[j], hr[j], side)) return false;
  return true;
}

// SpatialAction slot (star_gambit_gs.h:454-465) -> the unit type's own move index (slots 0-4) or cannon index
// (slots 5-9), -1 when the type has no such action (the play_move switch, star_gambit_gs.cc:1130-1218)
AZ_HD int sg_slot_index(int type, int slot) {
  // nibble tables per type, 0xF = none
  const u64 F = 0xFFFF0FF210ULL;            // slots 0,1,2 -> moves 0,1,2; slot 5 -> cannon 0
  const u64 C = 0xFF20140312ULL;            // 0->2 1->1 2->3 3->0 4->4 | 5->1 6->0 7->2
  const u64 D = 0x3021F3021FULL;            // 1->1 2->2 3->0 4->3 | 6->1 7->2 8->0 9->3
  const u64 t = type == SG_FIGHTER ? F : type == SG_CRUISER ? C : type == SG_DREAD ? D : ~0ULL;
  const int v = (int)((t >> (4 * slot)) & 15ULL);
  return v == 15 ? -1 : v;
}
// cannon (direction offset, source hex index) per type (get_cannon_info, star_gambit_gs.cc:201-231)
AZ_HD void sg_cannon(int type, int idx, int& doff, int& src) {
  if (type == SG_FIGHTER) { doff = 0; src = 0; }
  else if (type == SG_CRUISER) { doff = idx == 0 ? 1 : idx == 1 ? 0 : -1; src = 0; }
  else { doff = idx <= 1 ? 1 : 0; src = idx == 0 ? 2 : idx == 3 ? 1 : 0; }
}

// the ten action slots of unit `ui` that valid_moves() would mark (bit = SpatialAction slot)
AZ_HD_CALL u32 sg_unit_slots(const SGState& s, const SGBoards& b, int ui, int side) {
  const SGUnit& u = s.units[ui];
  if (u.player != s.player || u.hp == 0 || u.type == SG_PORTAL) return 0;
  int hq[3], hr[3];
  const int nh = sg_unit_hexes(u, side, hq, hr);
  u32 mask = 0;
  if (u.moves_left > 0) {
    SGOcc other = b.all;  // alive units never overlap: the others = all minus this unit's own hexes
    for (int j = 0; j < nh; ++j) {
      const int c = sg_cell(hq[j], hr[j]);
      other.w[c >> 6] &= ~(1ULL << (c & 63));
    }
    for (int slot = 0; slot < 5; ++slot) {
      const int dir = sg_slot_index(u.type, slot);
      if (dir < 0) continue;
      int nq, nr, nf;
      if (!sg_compute_move(u, dir, side, nq, nr, nf)) continue;
      int mq[3], mr[3];
      const int n = sg_hexes(u.type, u.player, nq, nr, nf, side, mq, mr);
      bool free_ = true;
      for (int j = 0; j < n; ++j) free_ = free_ && !sg_occ_test(other, mq[j], mr[j]);
      if (free_) mask |= 1u << slot;
    }
  }
  const SGOcc& enemy = b.pl[1 - (u.player & 1)];
  for (int slot = 5; slot < 10; ++slot) {
    const int ci = sg_slot_index(u.type, slot);
    if (ci < 0 || ((u.fired >> ci) & 1)) continue;
    int doff, src;
    sg_cannon(u.type, ci, doff, src);
    const int d = sg_rot(u.facing, doff);
    const int q1 = hq[src] + sg_dq(d), r1 = hr[src] + sg_dr(d), q2 = q1 + sg_dq(d), r2 = r1 + sg_dr(d);
    // has_target_in_range (star_gambit_gs.cc:669-713): an enemy at range 1, or nothing at range 1 and an enemy at
    // range 2 (a hex outside the board is never occupied, and the hex beyond it is outside as well)
    const bool hit = (sg_inb(q1, r1, side) && sg_occ_test(enemy, q1, r1)) ||
                     (!sg_occ_test(b.all, q1, r1) && sg_inb(q2, r2, side) && sg_occ_test(enemy, q2, r2));
    if (hit) mask |= 1u << slot;
  }
  return mask;
}

// is_deploy_valid (star_gambit_gs.cc:729-770) for every (type, facing): bit type * 6 + facing (absolute facing)
AZ_HD bool sg_deploy_anchor(int type, int player, int facing, int side, int& aq, int& ar) {
  // facings: dreadnought P0 {0,1,2,3} P1 {0,3,4,5}; fighter / cruiser P0 {1,2,3} P1 {4,5,0} (171-195)
  const u32 ok = type == SG_DREAD ? (player == 0 ? 0x0Fu : 0x39u) : (player == 0 ? 0x0Eu : 0x31u);
  if (!((ok >> facing) & 1u)) return false;
  const int dq0 = 0, dr0 = player == 0 ? side - 1 : -side + 1;  // get_deploy_hex (143-152)
  if (type == SG_DREAD) {  // one rear hex on the deploy hex (get_dreadnought_anchor_dir, 157-169)
    const int ad = player == 0 ? (facing == 0 ? 1 : facing == 3 ? 3 : 2) : (facing == 0 ? 0 : facing == 3 ? 4 : 5);
    aq = dq0 + sg_dq(ad); ar = dr0 + sg_dr(ad);
  } else if (type == SG_CRUISER) {  // the rear on the deploy hex
    aq = dq0 + sg_dq(facing); ar = dr0 + sg_dr(facing);
  } else {
    aq = dq0; ar = dr0;
  }
  return true;
}
AZ_HD_CALL bool sg_deploy_ok(const SGState& s, const SGBoards& b, int side, int type, int f) {
  const int p = s.player & 1;
  if (s.reserves[p][type] == 0) return false;
  int aq, ar;
  if (!sg_deploy_anchor(type, p, f, side, aq, ar)) return false;
  int hq[3], hr[3];
  const int n = sg_hexes(type, p, aq, ar, f, side, hq, hr);
  bool ok = true;
  for (int j = 0; j < n; ++j) ok = ok && sg_inb(hq[j], hr[j], side) && !sg_occ_test(b.all, hq[j], hr[j]);
  return ok;
}
AZ_HD u32 sg_deploy_mask(const SGState& s, const SGBoards& b, int side) {
  u32 mask = 0;
  for (int type = 0; type < 3; ++type)
    for (int f = 0; f < 6; ++f)
      if (sg_deploy_ok(s, b, side, type, f)) mask |= 1u << (type * 6 + f);
  return mask;
}
AZ_HD bool sg_turn_one(const SGState& s) { return s.turn == 1 || s.turn == 2; }

// valid_moves(): emit(id) for every legal action id of the space, in unit order (NOT ascending); returns the count
template <class Emit>
AZ_HD int sg_valid_moves(const SGState& s, const SGSpace& sp, Emit&& emit) {
  if (s.over) return 0;
  SGBoards b;
  sg_boards(s, sp.side, b);
  const bool p1 = s.player == 1;
  int count = 0;
  if (!sg_turn_one(s)) {
    for (int i = 0; i < (int)s.n_units; ++i) {
      const u32 m = sg_unit_slots(s, b, i, sp.side);
      if (!m) continue;
      int row = s.units[i].q + sp.side, col = s.units[i].r + sp.side;
      if (p1) { row = sp.dim - 1 - row; col = sp.dim - 1 - col; }
      const int base = ((row + sp.off) * sp.udim + (col + sp.off)) * 10;
      for (int slot = 0; slot < 10; ++slot)
        if ((m >> slot) & 1u) { emit(base + slot); ++count; }
    }
  }
  const u32 dm = sg_deploy_mask(s, b, sp.side);
  for (int i = 0; i < 18; ++i)
    if ((dm >> i) & 1u) {
      const int type = i / 6, f = i % 6;
      emit(sp.deploy_offset() + type * 6 + (p1 ? (f + 3) % 6 : f));
      ++count;
    }
  if (!sg_turn_one(s) && s.acted) { emit(sp.end_turn()); ++count; }
  return count;
}
struct SGNoEmit {
  AZ_HD void operator()(int) const {}
};

// compute_position_hash (star_gambit_gs.cc:1365-1382)
AZ_HD_CALL u64 sg_position_key(const SGState& s) {
  u64 h = (u64)s.player * 0x9e3779b97f4a7c15ULL;
  for (int i = 0; i < (int)s.n_units; ++i) {
    const SGUnit& u = s.units[i];
    if (u.hp == 0) continue;
    const u64 uh = (u64)u.type ^ ((u64)u.player << 8) ^ ((u64)u.hp << 12) ^ ((u64)u.facing << 20) ^
                   ((u64)(u.q + 10) << 28) ^ ((u64)(u.r + 10) << 36);
    h ^= uh * 0x517cc1b727220a95ULL;
  }
  return h;
}

// A flat key history (position_history_): the host classes and the replay kernels. `H` only needs clear(),
// push_count(key) -> number of equal keys including the new one, and count(key).
struct SGHistFlat {
  u64* keys;
  u32 len, cap;
  bool overflow;
  AZ_HD void clear() { len = 0; }
  AZ_HD int count(u64 k) const {
    int c = 0;
    for (u32 i = 0; i < len; ++i) c += keys[i] == k ? 1 : 0;
    return c;
  }
  AZ_HD int push_count(u64 k) {
    if (len < cap) keys[len++] = k; else overflow = true;
    return count(k) + (overflow ? 1 : 0);
  }
};

AZ_HD_CALL void sg_init(SGState& s, int variant) {  // StarGambitGS() (star_gambit_gs.cc:251-290); the caller pushes the first key
  for (int i = 0; i < (int)sizeof(SGState); ++i) ((u8*)&s)[i] = 0;
  s.variant = (u8)variant;
  const int side = sg_side(variant);
  for (int p = 0; p < 2; ++p)
    for (int t = 0; t < 3; ++t) s.reserves[p][t] = (u8)sg_start(variant, t);
  for (int p = 0; p < 2; ++p) {
    SGUnit& u = s.units[p];
    u.type = SG_PORTAL; u.player = (u8)p; u.slot = 0; u.hp = 5; u.facing = p == 0 ? 2 : 5;
    u.q = 0; u.r = (int8_t)(p == 0 ? side : -side);
    u.moves_left = 0; u.fired = 0;
  }
  s.n_units = 2;
  s.player = 0; s.turn = 1; s.acted = 0; s.over = 0; s.winner = -1;
}

AZ_HD_CALL void sg_check_game_end(SGState& s) {  // star_gambit_gs.cc:1313-1345
  for (int i = 0; i < (int)s.n_units; ++i)
    if (s.units[i].type == SG_PORTAL && s.units[i].hp == 0) {
      s.over = 1; s.winner = (int8_t)(1 - s.units[i].player);
      return;
    }
  for (int p = 0; p < 2; ++p) {
    bool ships = false;
    for (int i = 0; i < (int)s.n_units; ++i)
      ships = ships || (s.units[i].player == p && s.units[i].hp > 0 && s.units[i].type != SG_PORTAL);
    const bool res = s.reserves[p][0] > 0 || s.reserves[p][1] > 0 || s.reserves[p][2] > 0;
    if (!ships && !res) { s.over = 1; s.winner = (int8_t)(1 - p); return; }
  }
}
template <class H>
AZ_HD bool sg_check_repetition(SGState& s, H& hist) {  // star_gambit_gs.cc:1246-1261
  if (hist.push_count(sg_position_key(s)) >= 3) { s.over = 1; s.winner = 2; return true; }
  return false;
}
// `any_valid(s)`: does the side to move have a legal action (valid_moves().sum() != 0)? The scalar default walks the
// units; the device passes a lane-parallel test.
struct SGAnyValidScalar {
  AZ_HD bool operator()(const SGState& s, const SGSpace& sp) const { return sg_valid_moves(s, sp, SGNoEmit()) != 0; }
};
template <class H, class AnyValid>
AZ_HD_CALL void sg_end_turn(SGState& s, H& hist, const SGSpace& sp, AnyValid&& any_valid) {  // execute_end_turn (1263-1290)
  s.player = (u8)(1 - s.player);
  ++s.turn;
  s.acted = 0;
  if (s.turn > (u32)kSGMaxTurns) { s.over = 1; s.winner = 2; return; }
  if (sg_check_repetition(s, hist)) return;
  for (int i = 0; i < (int)s.n_units; ++i) {  // reset_turn_state
    SGUnit& u = s.units[i];
    if (u.player == s.player && u.hp > 0) { u.moves_left = (u8)sg_max_moves(u.type); u.fired = 0; }
  }
  if (!any_valid(s, sp)) { s.over = 1; s.winner = (int8_t)(1 - s.player); }
}
AZ_HD_CALL int sg_unit_at(const SGState& s, int q, int r, int side) {  // find_unit_at_hex (413-433)
  for (int i = 0; i < (int)s.n_units; ++i) {
    if (s.units[i].hp == 0) continue;
    int hq[3], hr[3];
    const int n = sg_unit_hexes(s.units[i], side, hq, hr);
    for (int j = 0; j < n; ++j)
      if (hq[j] == q && hr[j] == r) return i;
  }
  return -1;
}
// execute_fire's target search (978-1045): the first unit of ANY side (but the shooter) in range 1 then 2
AZ_HD_CALL bool sg_fire_target(const SGState& s, int ui, int ci, int side, int& target, int& damage) {
  const SGUnit& u = s.units[ui];
  if (ci >= sg_num_cannons(u.type)) return false;
  int doff, src, hq[3], hr[3];
  sg_cannon(u.type, ci, doff, src);
  const int nh = sg_unit_hexes(u, side, hq, hr);
  if (src >= nh) return false;
  const int d = sg_rot(u.facing, doff);
  int tq = hq[src], tr = hr[src];
  for (int range = 1; range <= 2; ++range) {
    tq += sg_dq(d); tr += sg_dr(d);
    if (!sg_inb(tq, tr, side)) continue;
    if (range == 2 && sg_unit_at(s, tq - sg_dq(d), tr - sg_dr(d), side) >= 0) break;  // line of sight
    const int t = sg_unit_at(s, tq, tr, side);
    if (t >= 0 && t != ui) { target = t; damage = range == 1 ? 2 : 1; return true; }
  }
  return false;
}

// play_move (star_gambit_gs.cc:1093-1238). Returns false only for ids outside the space.
template <class H, class AnyValid>
AZ_HD_CALL bool sg_play(SGState& s, H& hist, const SGSpace& sp, u32 move, AnyValid&& any_valid) {
  if (move >= (u32)sp.num_moves()) return false;
  if (move < (u32)sp.deploy_offset()) {
    const int slot = (int)(move % 10u), pos = (int)(move / 10u);
    int row = pos / sp.udim - sp.off, col = pos % sp.udim - sp.off;
    if (row < 0 || row >= sp.dim || col < 0 || col >= sp.dim) return true;  // no hex of this board: nothing happens
    if (s.player == 1) { row = sp.dim - 1 - row; col = sp.dim - 1 - col; }
    const int q = row - sp.side, r = col - sp.side;
    int ui = -1;
    for (int i = 0; i < (int)s.n_units && ui < 0; ++i) {
      const SGUnit& u = s.units[i];
      if (u.player == s.player && u.hp > 0 && u.q == q && u.r == r && u.type != SG_PORTAL) ui = i;
    }
    if (ui < 0) return true;  // no ship anchored here: not even a repetition entry
    SGUnit& u = s.units[ui];
    const int idx = sg_slot_index(u.type, slot);
    if (idx >= 0 && slot < 5) {
      int nq, nr, nf;
      if (sg_compute_move(u, idx, sp.side, nq, nr, nf)) {
        u.q = (int8_t)nq; u.r = (int8_t)nr; u.facing = (u8)nf;
        --u.moves_left;
        s.acted = 1;
      }
    } else if (idx >= 0) {
      u.fired |= (u8)(1u << idx);
      s.acted = 1;
      int target, damage;
      if (sg_fire_target(s, ui, idx, sp.side, target, damage)) {
        SGUnit& t = s.units[target];
        if (damage >= (int)t.hp) { t.hp = 0; sg_check_game_end(s); }
        else t.hp = (u8)(t.hp - damage);
      }
    }
    sg_check_repetition(s, hist);
  } else if (move < (u32)sp.end_turn()) {
    const int rel = (int)move - sp.deploy_offset(), type = rel / 6;
    int facing = rel % 6;
    if (s.player == 1) facing = (facing + 3) % 6;
    hist.clear();  // reserves change: earlier positions cannot repeat (execute_deploy, 1051-1087)
    int aq = 0, ar = 0;
    // the anchor formulas of execute_deploy do not test the facing; an illegal facing is never played by the search
    if (!sg_deploy_anchor(type, s.player, facing, sp.side, aq, ar)) {
      if (type == SG_CRUISER) { aq = sg_dq(facing); ar = (s.player == 0 ? sp.side - 1 : -sp.side + 1) + sg_dr(facing); }
      else { aq = 0; ar = s.player == 0 ? sp.side - 1 : -sp.side + 1; }
    }
    if (s.n_units < kSGMaxUnits) {
      int next = -1;  // get_next_slot (360-369): destroyed units count
      for (int i = 0; i < (int)s.n_units; ++i)
        if (s.units[i].player == s.player && s.units[i].type == type && (int)s.units[i].slot > next) next = s.units[i].slot;
      SGUnit& u = s.units[s.n_units++];
      u.type = (u8)type; u.player = s.player; u.slot = (u8)(next + 1); u.hp = (u8)sg_max_hp(type); u.facing = (u8)facing;
      u.q = (int8_t)aq; u.r = (int8_t)ar; u.moves_left = 0; u.fired = (u8)((1u << sg_num_cannons(type)) - 1u);
    }
    --s.reserves[s.player & 1][type];
    sg_end_turn(s, hist, sp, any_valid);
  } else {
    sg_end_turn(s, hist, sp, any_valid);
  }
  return true;
}
template <class H>
AZ_HD bool sg_play(SGState& s, H& hist, const SGSpace& sp, u32 move) { return sg_play(s, hist, sp, move, SGAnyValidScalar()); }
AZ_HD u32 sg_terminal(const SGState& s) {  // 0 running, 1 + winner, 3 draw (scores(), 1347-1363)
  if (!s.over) return 0;
  return s.winner == 2 ? 3u : s.winner == 0 ? 1u : s.winner == 1 ? 2u : 4u;  // 4: over without a winner (all zeros)
}

// ---- canonicalized() (star_gambit_gs.cc:1384-1669, Unified 2586-2616), element by element.
// Stage 1 (once per position): the unit standing on every cell of the MOVER'S frame + the broadcast values.
// Stage 2: out[ch][urow][ucol] = sg_canon_elem(...), any order (the device splits the elements over the lanes).
struct SGCanonCtx {
  u8 cell_unit[13 * 13];  // [row * dim + col] in the mover's frame: unit index + 1, 0 = empty
  float acted, rep, res[6], portal[2];
};
template <class H>
AZ_HD void sg_canon_ctx(const SGState& s, const H& hist, const SGSpace& sp, SGCanonCtx& c) {
  for (int i = 0; i < sp.dim * sp.dim; ++i) c.cell_unit[i] = 0;
  const bool p1 = s.player == 1;
  for (int i = 0; i < (int)s.n_units; ++i) {
    const SGUnit& u = s.units[i];
    if (u.hp == 0) continue;
    int hq[3], hr[3];
    const int n = sg_unit_hexes(u, sp.side, hq, hr);
    for (int j = 0; j < n; ++j) {
      const int q = p1 ? -hq[j] : hq[j], r = p1 ? -hr[j] : hr[j];
      if (q < -sp.side || q > sp.side || r < -sp.side || r > sp.side) continue;
      c.cell_unit[(q + sp.side) * sp.dim + (r + sp.side)] = (u8)(i + 1);
    }
  }
  c.acted = s.acted ? 1.0f : 0.0f;
  const int rc = hist.count(sg_position_key(s));
  c.rep = rc == 0 ? 0.0f : rc == 1 ? 0.5f : 1.0f;
  const int my = s.player & 1, opp = 1 - my;
  for (int t = 0; t < 3; ++t) {
    const int st = sg_start(s.variant, t);
    c.res[t] = st > 0 ? fdiv((float)s.reserves[my][t], (float)st) : 0.0f;
    c.res[3 + t] = st > 0 ? fdiv((float)s.reserves[opp][t], (float)st) : 0.0f;
  }
  c.portal[0] = c.portal[1] = 0.0f;
  for (int side_ = 0; side_ < 2; ++side_) {
    const int pl = side_ == 0 ? my : opp;
    for (int i = 0; i < (int)s.n_units; ++i) {  // find_unit_by_slot(player, PORTAL, 0): the first alive one
      const SGUnit& u = s.units[i];
      if (u.player == pl && u.type == SG_PORTAL && u.slot == 0 && u.hp > 0) { c.portal[side_] = fdiv((float)u.hp, 5.0f); break; }
    }
  }
}
AZ_HD float sg_canon_elem(const SGState& s, const SGCanonCtx& c, const SGSpace& sp, int ch, int urow, int ucol) {
  const int row = urow - sp.off, col = ucol - sp.off;
  if (row < 0 || row >= sp.dim || col < 0 || col >= sp.dim) return 0.0f;
  if (!sg_inb(row - sp.side, col - sp.side, sp.side)) return 0.0f;
  if (ch == 0) return 1.0f;
  if (ch >= 22) {
    if (ch == 22) return c.acted;
    if (ch == 23) return c.rep;
    if (ch < 30) return c.res[ch - 24];
    if (ch < 32) return c.portal[ch - 30];
    return (int)s.variant == ch - 32 ? 1.0f : 0.0f;
  }
  const int ui = c.cell_unit[row * sp.dim + col];
  if (!ui) return 0.0f;
  const SGUnit& u = s.units[ui - 1];
  const bool p1 = s.player == 1;
  if (ch <= 8) return (ch - 1) == ((u.player == s.player ? 0 : 4) + (int)u.type) ? 1.0f : 0.0f;
  if (ch == 15) return fdiv((float)u.hp, (float)sg_max_hp(u.type));
  if (u.type == SG_PORTAL) return 0.0f;
  if (ch <= 14) return (p1 ? (u.facing + 3) % 6 : (int)u.facing) == ch - 9 ? 1.0f : 0.0f;
  if (ch == 16) return fdiv((float)u.moves_left, (float)sg_max_moves(u.type));
  // 17-21: unfired cannons, on the anchor hex only; observation slots forward, fl, fr, rl, rr
  const int aq = p1 ? -u.q : u.q, ar = p1 ? -u.r : u.r;
  if (aq + sp.side != row || ar + sp.side != col) return 0.0f;
  const int sl = ch - 17;
  int ci = -1;
  if (u.type == SG_FIGHTER) ci = sl == 0 ? 0 : -1;
  else if (u.type == SG_CRUISER) ci = sl == 0 ? 1 : sl == 1 ? 0 : sl == 2 ? 2 : -1;
  else ci = sl == 1 ? 1 : sl == 2 ? 2 : sl == 3 ? 0 : sl == 4 ? 3 : -1;
  if (ci < 0) return 0.0f;
  return ((u.fired >> ci) & 1) ? 0.0f : 1.0f;
}

// ---- symmetries(): the NW-axis mirror (star_gambit_gs.cc:1671-1805, 2623-2727) as index maps: destination -> source.
// canonical: dst (ch, row, col) reads src (ch', BD-1-row', ...) — the forward map (row, col) -> (BD-1-row, row+col-S) is
// an involution on the cells it keeps in range, so the gather uses the same formula; cells nothing maps to stay 0.
AZ_HD int sg_mirror_canon_src(int dim, int side, int planes, int dst) {  // -1: zero
  const int cells = dim * dim, ch = dst / cells, rem = dst % cells, row = rem / dim, col = rem % dim;
  const int srow = dim - 1 - row, scol = col - srow + side;  // forward: new_col = row + col - side
  if (scol < 0 || scol >= dim) return -1;
  int sch = ch;
  if (ch >= 9 && ch <= 14) { const int m[6] = {4, 3, 2, 1, 0, 5}; sch = 9 + m[ch - 9]; }
  else if (ch >= 17 && ch <= 21) { const int m[5] = {0, 2, 1, 4, 3}; sch = 17 + m[ch - 17]; }
  (void)planes;
  return sch * cells + srow * dim + scol;
}
// policy: returns the SOURCE index whose value lands on dst, or -1 (zero). Off-board cells copy themselves; on-board
// cells move with the mirror and swap left / right slots (SLOT_MAP); deploys mirror their facing.
AZ_HD int sg_mirror_pi_src(int dim, int side, int dst) {
  const int spatial = dim * dim * 10;
  if (dst >= spatial + 18) return dst;
  if (dst >= spatial) {
    const int d = dst - spatial, type = d / 6, f = d % 6;
    const int md[6] = {3, 2, 1, 0, 5, 4}, mm[6] = {4, 3, 2, 1, 0, 5};  // both self-inverse
    return spatial + type * 6 + (type == 2 ? md[f] : mm[f]);
  }
  const int slot = dst % 10, pos = dst / 10, row = pos / dim, col = pos % dim;
  const bool on = sg_inb(row - side, col - side, side);
  const int sm[10] = {0, 2, 1, 4, 3, 5, 7, 6, 9, 8};
  // a valid hex maps to a valid hex (the mirror is a board symmetry), so dst on-board <=> its source on-board
  if (!on) return dst;
  const int srow = dim - 1 - row, scol = col - srow + side;
  if (scol < 0 || scol >= dim) return -1;
  return (srow * dim + scol) * 10 + sm[slot];
}

}  // namespace b2az
