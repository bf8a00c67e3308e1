// az_tafl_kernels.h — batched tafl game kernels behind the C ABI (b2az_tafl_replay). Included at the end of
// az_engine.cu (same translation unit: it uses that file's error / launch helpers).
//
// One WARP per game transcript. Every lane replays the moves on its own register copy of the 3-bitboard state
// (identical work, no divergence), so no state has to be broadcast; the lanes then split the OUTPUT of each
// position between them: the board bytes, the legal-move mask (2S bytes per source square, squares lane,
// lane + 32, ...), the canonical floats (coalesced 4 B stores), and a shuffle reduction for the legal-move
// count. The repetition history (48 B keys since the last capture) lives in HBM, one row per game.
// HBM bytes per position (what the kernel is bound by): 4*PLANES*S*S canonical + 2*S^3 mask + 3*S^2 board + 14
// = 2,219 B (Brandubh), 6,911 B (OpenTafl), 6,427 B (Tawlbwrdd).
#pragma once

#include "az_tafl.h"

namespace b2az {

struct TaflReplayArgs {
  u32 n, max_len, max_turns;
  const u16* moves;   // [n][max_len]
  const u32* lens;    // [n]
  TaflKey* hist;      // [n][max_len + 2]
  // outputs, [n][max_len + 1][...]; any may be null
  signed char* boards;  // [..][3*S*S]
  u8* players;
  u32* turns;
  u8* reps;
  u8* terminal;
  u32* n_valid;
  u8* valid;          // [..][2*S^3]
  float* canonical;   // [..][PLANES*S*S]
  i32* status;        // [n]: 0, or B2AZ_EMOVE at the first move the reference would have thrown on
};

// lane's share of the outputs of one position (lane in [0, nl))
template <int GAME>
AZ_HD void tafl_emit(const TaflReplayArgs& a, size_t row, const TaflState& s, u32 lane, u32 nl) {
  typedef Tafl<GAME> T;
  if (a.boards)
    for (u32 e = lane; e < (u32)T::BOARD_BYTES; e += nl) a.boards[row * T::BOARD_BYTES + e] = T::board_byte(s, e);
  if (a.valid)
    for (u32 c = lane; c < (u32)T::CELLS; c += nl) T::valid_bytes(s, (int)c, a.valid + row * T::A + c * (2u * T::S));
  if (a.canonical)
    for (u32 e = lane; e < (u32)T::CANON; e += nl) a.canonical[row * T::CANON + e] = T::canon_elem(s, e);
  if (lane == 0) {
    if (a.players) a.players[row] = s.player;
    if (a.turns) a.turns[row] = s.turn;
    if (a.reps) a.reps[row] = s.rep;
    if (a.terminal) a.terminal[row] = (u8)T::terminal(s);
  }
}

#ifndef B2AZ_HOST_EMU
__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
// Everything one position writes, by a whole warp (`a`'s pointers; row = output row):
//   legal moves  the per-square slide masks are computed once (squares lane, lane + 32, ...) into shared memory,
//                then the A mask bytes leave as coalesced 16-bit stores (2S is even: a byte pair never straddles
//                two source squares and every row starts 2-byte aligned); the shuffle-reduced count also answers
//                scores()'s "side to move has no legal move"
//   canonical    plane by plane, consecutive lanes = consecutive floats; planes 3.. are constants per position
//   board bytes, terminal code, scalars
// Per-warp shared memory of the emit code: the occupancy of every row and column as S-bit lines, and the
// legal-move mask of the position as bytes, staged so that it leaves in aligned 4-byte words.
template <int GAME>
struct TaflWarpSmem {
  u32 lines[2 * Tafl<GAME>::S + 2];
  u32 bytes[(Tafl<GAME>::A + 8) / 4];
};
template <int GAME>
__device__ __forceinline__ void tafl_emit_warp(const TaflReplayArgs& a, size_t row, const TaflState& s,
                                                TaflWarpSmem<GAME>& sm, u32 lane) {
  typedef Tafl<GAME> T;
  constexpr int S = T::S, CELLS = T::CELLS, CHUNKS = (CELLS + 31) / 32;
  // (1) occupancy lines: lane r < S extracts row r, lane S + c gathers column c
  const B128 occ = s.king | s.def | s.atk;
  if (lane < (u32)S) {
    sm.lines[lane] = b128_bits(occ, S * (int)lane, S);
  } else if (lane < 2u * S) {
    const int c = (int)lane - S;
    u32 v = 0;
#pragma unroll
    for (int h = 0; h < S; ++h) v |= (b128_test(occ, S * h + c) ? 1u : 0u) << h;
    sm.lines[lane] = v;
  }
  __syncwarp();
  // (2) legal moves: squares lane, lane + 32, ...: the 2S mask bytes of a square into the staging row, at the
  // same 4-byte phase as their place in global memory (rows are 2-byte aligned: 2S^3 is even)
  u8* vdst = a.valid ? a.valid + row * T::A : nullptr;
  const u32 mis = a.valid ? (u32)(reinterpret_cast<size_t>(vdst) & 3u) : 0u;
  unsigned short* stage = reinterpret_cast<unsigned short*>(sm.bytes) + (mis >> 1);
  const B128 mine = T::own(s);
  u32 cnt = 0;
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j) {
    const u32 c = 32u * j + lane;
    if (c < (u32)CELLS) {
      u32 r = 0, cl = 0;
      if ((b128_word(mine, j) >> lane) & 1u) {
        const int h = (int)(c / (u32)S), w = (int)(c % (u32)S);
        T::slides_lines(((b128_word(s.king, j) >> lane) & 1u) != 0, h, w, sm.lines[h], sm.lines[S + w], r, cl);
      }
      cnt += (u32)__popc(r) + (u32)__popc(cl);
      if (vdst) {
        const u32 bits = r | (cl << S);
#pragma unroll
        for (int t = 0; t < S; ++t)
          stage[c * S + t] = (unsigned short)(((bits >> (2 * t)) & 1u) | (((bits >> (2 * t + 1)) & 1u) << 8));
      }
    }
  }
  __syncwarp();
  if (vdst) {
    const u32 end = mis + (u32)T::A, first = (mis + 3u) >> 2, last = end >> 2;
    u32* gw = reinterpret_cast<u32*>(vdst - mis);
    for (u32 wi = first + lane; wi < last; wi += 32u) gw[wi] = sm.bytes[wi];
    if (lane == 0) {
      const unsigned short* s16 = reinterpret_cast<const unsigned short*>(sm.bytes);
      unsigned short* g16 = reinterpret_cast<unsigned short*>(vdst - mis);
      if (mis & 3u) g16[mis >> 1] = s16[mis >> 1];
      if (end & 3u) g16[last * 2u] = s16[last * 2u];
    }
  }
  __syncwarp();
  cnt = warp_sum(cnt);
  // (3) canonical planes and board bytes: consecutive lanes = consecutive squares; planes 3.. are constants
  if (a.canonical) {
    float* out = a.canonical + row * T::CANON;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) {
      const u32 c = 32u * j + lane;
      if (c < (u32)CELLS) {
        out[c] = (float)((b128_word(s.king, j) >> lane) & 1u);
        out[CELLS + c] = (float)((b128_word(s.def, j) >> lane) & 1u);
        out[2 * CELLS + c] = (float)((b128_word(s.atk, j) >> lane) & 1u);
      }
    }
#pragma unroll
    for (int pl = 3; pl < T::PLANES; ++pl) {
      const float v = T::canon_elem(s, (u32)(pl * CELLS));  // constant over the plane
#pragma unroll
      for (int j = 0; j < CHUNKS; ++j) {
        const u32 c = 32u * j + lane;
        if (c < (u32)CELLS) out[pl * CELLS + c] = v;
      }
    }
  }
  if (a.boards) {
    signed char* out = a.boards + row * T::BOARD_BYTES;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) {
      const u32 c = 32u * j + lane;
      if (c < (u32)CELLS) {
        out[c] = (signed char)((b128_word(s.king, j) >> lane) & 1u);
        out[CELLS + c] = (signed char)((b128_word(s.def, j) >> lane) & 1u);
        out[2 * CELLS + c] = (signed char)((b128_word(s.atk, j) >> lane) & 1u);
      }
    }
  }
  const u32 pre = T::terminal_pre(s);  // uniform over the warp
  if (lane == 0) {
    if (a.n_valid) a.n_valid[row] = cnt;
    if (a.terminal) a.terminal[row] = (u8)(pre ? pre : T::terminal_post(s, cnt != 0));
    if (a.players) a.players[row] = s.player;
    if (a.turns) a.turns[row] = s.turn;
    if (a.reps) a.reps[row] = s.rep;
  }
}

#ifndef B2AZ_TAFL_MINB
#define B2AZ_TAFL_MINB 8  /* 64 registers: 32 warps per SM (measured: OpenTafl 407 -> 451 M positions/s) */
#endif
template <int GAME>
__global__ void __launch_bounds__(128, B2AZ_TAFL_MINB) k_tafl_replay(TaflReplayArgs a) {
  typedef Tafl<GAME> T;
  __shared__ TaflWarpSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const u32 warp = GLOBAL_TID >> 5, nwarps = GLOBAL_NT >> 5;
  for (u32 g = warp; g < a.n; g += nwarps) {
    // every lane keeps its own register copy of the state and replays the move (uniform, no broadcast needed)
    TaflState s;
    T::init(s, a.max_turns);
    TaflKey* hist = a.hist + (size_t)g * (a.max_len + 2u);
    u32 hist_len = 0;
    const u32 len = a.lens[g] < a.max_len ? a.lens[g] : a.max_len;
    i32 st = 0;
    for (u32 k = 0; k <= len; ++k) {
      if (k > 0) {
        // play_move incl. the repetition table (Tafl::play_hist), the history scan split over the lanes
        const u32 mv = a.moves[(size_t)g * a.max_len + (k - 1u)];
        if (s.turn == 0) {
          if (lane == 0) hist[0] = T::key(s);
          hist_len = 1;
          __syncwarp();
        }
        bool cap;
        if (!T::play(s, mv, &cap)) {
          st = B2AZ_EMOVE;
          break;
        }
        if (cap) hist_len = 0;
        const TaflKey key = T::key(s);
        u32 same = 0;
        for (u32 i = lane; i < hist_len; i += 32u) same += T::key_eq(hist[i], key) ? 1u : 0u;
        same = warp_sum(same) + 1u;
        if (lane == 0) hist[hist_len] = key;
        ++hist_len;
        __syncwarp();
        s.rep = (u8)(same > 255u ? 255u : same);
      }
      tafl_emit_warp<GAME>(a, (size_t)g * (a.max_len + 1u) + k, s, sm[wib], lane);
    }
    if (lane == 0 && a.status) a.status[g] = st;
  }
}
#endif

template <int GAME>
void tafl_replay_host_one(const TaflReplayArgs& a, u32 g) {  // the same work by one thread (host-emulation build)
  typedef Tafl<GAME> T;
  TaflState s;
  T::init(s, a.max_turns);
  TaflKey* hist = a.hist + (size_t)g * (a.max_len + 2u);
  u32 hist_len = 0;
  const u32 len = a.lens[g] < a.max_len ? a.lens[g] : a.max_len;
  i32 st = 0;
  for (u32 k = 0; k <= len; ++k) {
    if (k > 0 && !T::play_hist(s, a.moves[(size_t)g * a.max_len + (k - 1u)], hist, hist_len)) {
      st = B2AZ_EMOVE;
      break;
    }
    const size_t row = (size_t)g * (a.max_len + 1u) + k;
    tafl_emit<GAME>(a, row, s, 0u, 1u);
    if (a.n_valid) a.n_valid[row] = T::moves(s, nullptr);
    // cross-check of the device kernel's line-based move generation against the square walk (CPU test-suite)
    const B128 occ = s.king | s.def | s.atk;
    u32 lines[2 * T::S];
    for (int i = 0; i < T::S; ++i) {
      lines[i] = b128_bits(occ, T::S * i, T::S);
      u32 v = 0;
      for (int h = 0; h < T::S; ++h) v |= (b128_test(occ, T::S * h + i) ? 1u : 0u) << h;
      lines[T::S + i] = v;
    }
    for (int c = 0; c < T::CELLS; ++c)
      if (b128_test(T::own(s), c)) {
        u32 r0, c0, r1, c1;
        T::slides(s, c / T::S, c % T::S, r0, c0);
        T::slides_lines(b128_test(s.king, c), c / T::S, c % T::S, lines[c / T::S], lines[T::S + c % T::S], r1, c1);
        if (r0 != r1 || c0 != c1) st = -99;
      }
  }
  if (a.status) a.status[g] = st;
}

template <int GAME>
int tafl_replay_impl(int device, uint32_t n, uint32_t max_len, uint32_t max_turns, const uint16_t* moves,
                     const uint32_t* lens, int8_t* boards, uint8_t* players, uint32_t* turns, uint8_t* reps,
                     uint8_t* terminal, uint32_t* n_valid, uint8_t* valid, float* canonical, int32_t* status) {
  typedef Tafl<GAME> T;
  const size_t rows = (size_t)n * (max_len + 1u);
#ifdef B2AZ_HOST_EMU
  (void)device; (void)rows;
  std::vector<TaflKey> hist((size_t)n * (max_len + 2u));
  TaflReplayArgs a{n, max_len, max_turns, moves, lens, hist.data(), reinterpret_cast<signed char*>(boards), players,
                   turns, reps, terminal, n_valid, valid, canonical, status};
  for (u32 g = 0; g < n; ++g) tafl_replay_host_one<GAME>(a, g);
  return 0;
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  TaflReplayArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.max_len = max_len; a.max_turns = max_turns;
  std::vector<void*> owned;
  bool oom = false;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { oom = true; return nullptr; }
    owned.push_back(d);
    if (src) cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    else cudaMemset(d, 0, bytes);
    return d;
  };
  a.moves = static_cast<const u16*>(up(moves, (size_t)n * max_len * 2));
  a.lens = static_cast<const u32*>(up(lens, (size_t)n * 4));
  a.hist = static_cast<TaflKey*>(up(nullptr, (size_t)n * (max_len + 2u) * sizeof(TaflKey)));
  a.boards = boards ? static_cast<signed char*>(up(nullptr, rows * T::BOARD_BYTES)) : nullptr;
  a.players = players ? static_cast<u8*>(up(nullptr, rows)) : nullptr;
  a.turns = turns ? static_cast<u32*>(up(nullptr, rows * 4)) : nullptr;
  a.reps = reps ? static_cast<u8*>(up(nullptr, rows)) : nullptr;
  a.terminal = terminal ? static_cast<u8*>(up(nullptr, rows)) : nullptr;
  a.n_valid = n_valid ? static_cast<u32*>(up(nullptr, rows * 4)) : nullptr;
  a.valid = valid ? static_cast<u8*>(up(nullptr, rows * T::A)) : nullptr;
  a.canonical = canonical ? static_cast<float*>(up(nullptr, rows * T::CANON * 4)) : nullptr;
  a.status = status ? static_cast<i32*>(up(nullptr, (size_t)n * 4)) : nullptr;
  int rc = 0;
  if (oom) rc = fail(B2AZ_ENOMEM, "b2az_tafl_replay: cudaMalloc failed");
  if (!rc) {
    const u32 warps_per_cta = 4;
    const u32 ctas = std::max(1u, std::min((n + warps_per_cta - 1u) / warps_per_cta, 148u * 8u));
    k_tafl_replay<GAME><<<ctas, warps_per_cta * 32u>>>(a);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) rc = fail(B2AZ_ECUDA, std::string("k_tafl_replay: ") + cudaGetErrorString(err));
  }
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (dst && src && !rc) cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
  };
  down(boards, a.boards, rows * T::BOARD_BYTES);
  down(players, a.players, rows);
  down(turns, a.turns, rows * 4);
  down(reps, a.reps, rows);
  down(terminal, a.terminal, rows);
  down(n_valid, a.n_valid, rows * 4);
  down(valid, a.valid, rows * T::A);
  down(canonical, a.canonical, rows * T::CANON * 4);
  down(status, a.status, (size_t)n * 4);
  for (void* d : owned) cudaFree(d);
  return rc;
#endif
}

// ---- arbitrary positions (b2az_tafl_positions): one warp per position
struct TaflPosArgs {
  u32 n, max_turns;
  const signed char* boards;  // [n][3*S*S]
  const u8* players;
  const u32* turns;
  const u8* reps;
  const u32* moves;           // may be null; 0xFFFFFFFF = no move
  TaflReplayArgs out;         // terminal / n_valid / valid / canonical of the INPUT position, row = position
  signed char* boards_out;    // [n][3*S*S] after the move
  u8* captured_any;           // [n]
  i32* status;                // [n]
};
template <int GAME>
AZ_HD void tafl_load_position(const TaflPosArgs& a, u32 i, TaflState& s) {
  typedef Tafl<GAME> T;
  s.king = s.def = s.atk = b128(0, 0);
  const signed char* b = a.boards + (size_t)i * T::BOARD_BYTES;
  for (int c = 0; c < T::CELLS; ++c) {
    if (b[c]) s.king = s.king | b128_bit(c);
    if (b[T::CELLS + c]) s.def = s.def | b128_bit(c);
    if (b[2 * T::CELLS + c]) s.atk = s.atk | b128_bit(c);
  }
  s.player = a.players[i];
  s.turn = a.turns[i];
  s.max_turns = (u16)a.max_turns;
  s.rep = a.reps[i];
}
template <int GAME>
AZ_HD void tafl_position_move(const TaflPosArgs& a, u32 i, TaflState& s, u32 lane, u32 nl) {
  typedef Tafl<GAME> T;
  i32 st = 0;
  if (a.moves && a.moves[i] != 0xFFFFFFFFu) {
    bool cap = false;
    if (!T::play(s, a.moves[i], &cap)) st = B2AZ_EMOVE;
    if (st == 0 && a.boards_out)
      for (u32 e = lane; e < (u32)T::BOARD_BYTES; e += nl) a.boards_out[(size_t)i * T::BOARD_BYTES + e] = T::board_byte(s, e);
    if (lane == 0 && a.captured_any) a.captured_any[i] = cap ? 1 : 0;
  }
  if (lane == 0 && a.status) a.status[i] = st;
}
template <int GAME>
void tafl_position_one(const TaflPosArgs& a, u32 i) {  // host-emulation build
  typedef Tafl<GAME> T;
  TaflState s;
  tafl_load_position<GAME>(a, i, s);
  tafl_emit<GAME>(a.out, i, s, 0u, 1u);
  if (a.out.n_valid) a.out.n_valid[i] = T::moves(s, nullptr);
  tafl_position_move<GAME>(a, i, s, 0u, 1u);
}
#ifndef B2AZ_HOST_EMU
template <int GAME>
__global__ void __launch_bounds__(128) k_tafl_positions(TaflPosArgs a) {
  typedef Tafl<GAME> T;
  __shared__ TaflWarpSmem<GAME> sm[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  for (u32 i = GLOBAL_TID >> 5; i < a.n; i += GLOBAL_NT >> 5) {
    TaflState s;
    tafl_load_position<GAME>(a, i, s);
    tafl_emit_warp<GAME>(a.out, i, s, sm[wib], lane);
    tafl_position_move<GAME>(a, i, s, lane, 32u);
  }
}
#endif
template <int GAME>
int tafl_positions_impl(int device, uint32_t n, uint32_t max_turns, const int8_t* boards, const uint8_t* players,
                        const uint32_t* turns, const uint8_t* reps, const uint32_t* moves, uint8_t* terminal,
                        uint32_t* n_valid, uint8_t* valid, float* canonical, int8_t* boards_out, uint8_t* captured_any,
                        int32_t* status) {
  typedef Tafl<GAME> T;
  TaflPosArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.max_turns = max_turns;
#ifdef B2AZ_HOST_EMU
  (void)device;
  a.boards = reinterpret_cast<const signed char*>(boards); a.players = players; a.turns = turns; a.reps = reps;
  a.moves = moves;
  a.out.terminal = terminal; a.out.n_valid = n_valid; a.out.valid = valid; a.out.canonical = canonical;
  a.boards_out = reinterpret_cast<signed char*>(boards_out); a.captured_any = captured_any; a.status = status;
  for (u32 i = 0; i < n; ++i) tafl_position_one<GAME>(a, i);
  return 0;
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  std::vector<void*> owned;
  bool oom = false;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { oom = true; return nullptr; }
    owned.push_back(d);
    if (src) cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    else cudaMemset(d, 0, bytes);
    return d;
  };
  a.boards = static_cast<const signed char*>(up(boards, (size_t)n * T::BOARD_BYTES));
  a.players = static_cast<const u8*>(up(players, n));
  a.turns = static_cast<const u32*>(up(turns, (size_t)n * 4));
  a.reps = static_cast<const u8*>(up(reps, n));
  a.moves = moves ? static_cast<const u32*>(up(moves, (size_t)n * 4)) : nullptr;
  a.out.terminal = terminal ? static_cast<u8*>(up(nullptr, n)) : nullptr;
  a.out.n_valid = n_valid ? static_cast<u32*>(up(nullptr, (size_t)n * 4)) : nullptr;
  a.out.valid = valid ? static_cast<u8*>(up(nullptr, (size_t)n * T::A)) : nullptr;
  a.out.canonical = canonical ? static_cast<float*>(up(nullptr, (size_t)n * T::CANON * 4)) : nullptr;
  a.boards_out = boards_out ? static_cast<signed char*>(up(nullptr, (size_t)n * T::BOARD_BYTES)) : nullptr;
  a.captured_any = captured_any ? static_cast<u8*>(up(nullptr, n)) : nullptr;
  a.status = status ? static_cast<i32*>(up(nullptr, (size_t)n * 4)) : nullptr;
  int rc = 0;
  if (oom) rc = fail(B2AZ_ENOMEM, "b2az_tafl_positions: cudaMalloc failed");
  if (!rc) {
    k_tafl_positions<GAME><<<std::max(1u, std::min((n + 3u) / 4u, 148u * 8u)), 128>>>(a);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) rc = fail(B2AZ_ECUDA, std::string("k_tafl_positions: ") + cudaGetErrorString(err));
  }
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (dst && src && !rc) cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
  };
  down(terminal, a.out.terminal, n);
  down(n_valid, a.out.n_valid, (size_t)n * 4);
  down(valid, a.out.valid, (size_t)n * T::A);
  down(canonical, a.out.canonical, (size_t)n * T::CANON * 4);
  down(boards_out, a.boards_out, (size_t)n * T::BOARD_BYTES);
  down(captured_any, a.captured_any, n);
  down(status, a.status, (size_t)n * 4);
  for (void* d : owned) cudaFree(d);
  return rc;
#endif
}

template <int GAME>
int tafl_replay_device_impl(const TaflReplayArgs& a, void* stream) {
#ifdef B2AZ_HOST_EMU
  (void)stream;
  for (u32 g = 0; g < a.n; ++g) tafl_replay_host_one<GAME>(a, g);
  return 0;
#else
  const u32 ctas = std::max(1u, std::min((a.n + 3u) / 4u, 148u * 8u));
  k_tafl_replay<GAME><<<ctas, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  CUDA_TRY(cudaGetLastError());
  return 0;
#endif
}

}  // namespace b2az

// ---- tafl_helper::eightSym (tafl_helper.h:16-149): the eight symmetric images of a training sample, in the
// reference's order [base, rot90, rot180, rot270, mirror(base), mirror(rot90), mirror(rot180), mirror(rot270)]
// (rot = rot90Clockwise, mirror = mirrorWidth). Pure data movement: every output element gathers its source.
//   rot90Clockwise  canonical out(c, h, w) = in(c, S-1-w, h)
//                   pi  out[(h,w) row-slide to w'] = in[(S-1-w, h) column-slide to S-1-w']
//                       out[(h,w) column-slide to h'] = in[(S-1-w, h) row-slide to h']
//   mirrorWidth     canonical out(c, h, w) = in(c, h, S-1-w)
//                   pi  out[(h,w) row-slide to w'] = in[(h, S-1-w) row-slide to S-1-w'];  column slides keep h'
// Symmetry i = i%4 rotations then (i >= 4) a mirror, so an output index is pulled back through the mirror first and
// then through the rotations.
namespace b2az {
template <int S>
AZ_HD u32 tafl_sym_src_cell(u32 sym, u32 h, u32 w) {  // source (h, w) of output cell (h, w), as h * S + w
  if (sym >= 4u) w = (u32)S - 1u - w;
  for (u32 r = 0; r < (sym & 3u); ++r) {
    const u32 nh = (u32)S - 1u - w, nw = h;
    h = nh; w = nw;
  }
  return h * (u32)S + w;
}
template <int S>
AZ_HD u32 tafl_sym_src_move(u32 sym, u32 mv) {  // source move id of output move id
  u32 loc = mv % (u32)(2 * S), cell = mv / (u32)(2 * S);
  u32 h = cell / (u32)S, w = cell % (u32)S;
  bool col = loc >= (u32)S;
  u32 t = col ? loc - (u32)S : loc;
  if (sym >= 4u) {
    w = (u32)S - 1u - w;
    if (!col) t = (u32)S - 1u - t;
  }
  for (u32 r = 0; r < (sym & 3u); ++r) {
    const u32 nh = (u32)S - 1u - w, nw = h;
    h = nh; w = nw;
    if (!col) { col = true; t = (u32)S - 1u - t; }  // a row slide was a column slide before the rotation
    else col = false;
  }
  return (h * (u32)S + w) * (u32)(2 * S) + (col ? (u32)S + t : t);
}
#ifndef B2AZ_HOST_EMU
template <int S>
__global__ void k_tafl_symmetries(u32 n, u32 planes, const float* __restrict__ canon, const float* __restrict__ v,
                                  const float* __restrict__ pi, float* __restrict__ canon_out, float* __restrict__ v_out,
                                  float* __restrict__ pi_out) {
  const size_t C = (size_t)planes * S * S, A = (size_t)2 * S * S * S, per = C + A + 3;
  const size_t total = (size_t)n * 8u * per;
  for (size_t i = GLOBAL_TID; i < total; i += GLOBAL_NT) {
    const size_t row = i / per, e = i % per;  // row = sample * 8 + sym
    const u32 sample = (u32)(row / 8u), sym = (u32)(row % 8u);
    if (e < C) {
      const u32 c = (u32)(e / (S * S)), cell = (u32)(e % (S * S));
      canon_out[row * C + e] = canon[(size_t)sample * C + (size_t)c * S * S + tafl_sym_src_cell<S>(sym, cell / S, cell % S)];
    } else if (e < C + A) {
      const u32 mv = (u32)(e - C);
      pi_out[row * A + mv] = pi[(size_t)sample * A + tafl_sym_src_move<S>(sym, mv)];
    } else {
      const u32 j = (u32)(e - C - A);
      v_out[row * 3 + j] = v[(size_t)sample * 3 + j];
    }
  }
}
#endif
}  // namespace b2az

extern "C" int b2az_tafl_symmetries(int device, uint32_t game, uint32_t n, const float* canon_host, const float* v_host,
                                    const float* pi_host, float* canon_out, float* v_out, float* pi_out) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!canon_host || !v_host || !pi_host || !canon_out || !v_out || !pi_out) return fail(B2AZ_EINVAL, "null argument");
  if (game > B2AZ_TAFL_TAWLBWRDD) return fail(B2AZ_EINVAL, "unknown tafl game");
  const u32 S = game == B2AZ_TAFL_BRANDUBH ? 7u : 11u, planes = game == B2AZ_TAFL_OPENTAFL ? 8u : 7u;
  const size_t C = (size_t)planes * S * S, A = (size_t)2 * S * S * S;
#ifdef B2AZ_HOST_EMU
  (void)device;
  for (size_t row = 0; row < (size_t)n * 8u; ++row) {
    const u32 sample = (u32)(row / 8u), sym = (u32)(row % 8u);
    for (size_t e = 0; e < C; ++e) {
      const u32 c = (u32)(e / (S * S)), cell = (u32)(e % (S * S));
      const u32 src = S == 7u ? tafl_sym_src_cell<7>(sym, cell / S, cell % S) : tafl_sym_src_cell<11>(sym, cell / S, cell % S);
      canon_out[row * C + e] = canon_host[(size_t)sample * C + (size_t)c * S * S + src];
    }
    for (size_t mv = 0; mv < A; ++mv)
      pi_out[row * A + mv] = pi_host[(size_t)sample * A + (S == 7u ? tafl_sym_src_move<7>(sym, (u32)mv) : tafl_sym_src_move<11>(sym, (u32)mv))];
    for (int j = 0; j < 3; ++j) v_out[row * 3 + j] = v_host[(size_t)sample * 3 + j];
  }
  return 0;
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  float *dc = nullptr, *dv = nullptr, *dp = nullptr, *oc = nullptr, *ov = nullptr, *op = nullptr;
  int rc = 0;
  auto al = [&](float** p, size_t count) { if (!rc && cudaMalloc(p, count * 4) != cudaSuccess) rc = fail(B2AZ_ENOMEM, "b2az_tafl_symmetries: cudaMalloc failed"); };
  al(&dc, n * C); al(&dv, (size_t)n * 3); al(&dp, n * A); al(&oc, (size_t)n * 8 * C); al(&ov, (size_t)n * 24); al(&op, (size_t)n * 8 * A);
  if (!rc) {
    cudaMemcpy(dc, canon_host, n * C * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, v_host, (size_t)n * 12, cudaMemcpyHostToDevice);
    cudaMemcpy(dp, pi_host, n * A * 4, cudaMemcpyHostToDevice);
    if (S == 7u) k_tafl_symmetries<7><<<148 * 8, 256>>>(n, planes, dc, dv, dp, oc, ov, op);
    else k_tafl_symmetries<11><<<148 * 8, 256>>>(n, planes, dc, dv, dp, oc, ov, op);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) rc = fail(B2AZ_ECUDA, std::string("k_tafl_symmetries: ") + cudaGetErrorString(err));
  }
  if (!rc) {
    cudaMemcpy(canon_out, oc, (size_t)n * 8 * C * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(v_out, ov, (size_t)n * 24 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(pi_out, op, (size_t)n * 8 * A * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(dc); cudaFree(dv); cudaFree(dp); cudaFree(oc); cudaFree(ov); cudaFree(op);
  return rc;
#endif
}

extern "C" int b2az_tafl_replay_device(uint32_t game, uint32_t n, uint32_t max_len, uint32_t max_turns,
                                       const uint16_t* moves_dev, const uint32_t* lens_dev, void* hist_dev,
                                       int8_t* boards_dev, uint8_t* terminal_dev, uint32_t* n_valid_dev,
                                       uint8_t* valid_dev, float* canonical_dev, int32_t* status_dev, void* stream) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!moves_dev || !lens_dev || !hist_dev || max_len == 0) return fail(B2AZ_EINVAL, "null argument");
  if (max_turns == 0 || max_turns > 65535u) return fail(B2AZ_EINVAL, "max_turns must fit uint16_t");
  TaflReplayArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.max_len = max_len; a.max_turns = max_turns;
  a.moves = moves_dev; a.lens = lens_dev; a.hist = static_cast<TaflKey*>(hist_dev);
  a.boards = reinterpret_cast<signed char*>(boards_dev); a.terminal = terminal_dev; a.n_valid = n_valid_dev;
  a.valid = valid_dev; a.canonical = canonical_dev; a.status = status_dev;
  switch (game) {
    case B2AZ_TAFL_BRANDUBH: return tafl_replay_device_impl<B2AZ_TAFL_BRANDUBH>(a, stream);
    case B2AZ_TAFL_OPENTAFL: return tafl_replay_device_impl<B2AZ_TAFL_OPENTAFL>(a, stream);
    case B2AZ_TAFL_TAWLBWRDD: return tafl_replay_device_impl<B2AZ_TAFL_TAWLBWRDD>(a, stream);
  }
  return fail(B2AZ_EINVAL, "unknown tafl game");
}

extern "C" int b2az_tafl_positions(int device, uint32_t game, uint32_t n, uint32_t max_turns, const int8_t* boards,
                                   const uint8_t* players, const uint32_t* turns, const uint8_t* reps,
                                   const uint32_t* moves, uint8_t* terminal, uint32_t* n_valid, uint8_t* valid,
                                   float* canonical, int8_t* boards_out, uint8_t* captured_any, int32_t* status) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!boards || !players || !turns || !reps) return fail(B2AZ_EINVAL, "null argument");
  if (max_turns == 0 || max_turns > 65535u) return fail(B2AZ_EINVAL, "max_turns must fit uint16_t");
  switch (game) {
    case B2AZ_TAFL_BRANDUBH:
      return tafl_positions_impl<B2AZ_TAFL_BRANDUBH>(device, n, max_turns, boards, players, turns, reps, moves, terminal,
                                                     n_valid, valid, canonical, boards_out, captured_any, status);
    case B2AZ_TAFL_OPENTAFL:
      return tafl_positions_impl<B2AZ_TAFL_OPENTAFL>(device, n, max_turns, boards, players, turns, reps, moves, terminal,
                                                     n_valid, valid, canonical, boards_out, captured_any, status);
    case B2AZ_TAFL_TAWLBWRDD:
      return tafl_positions_impl<B2AZ_TAFL_TAWLBWRDD>(device, n, max_turns, boards, players, turns, reps, moves, terminal,
                                                      n_valid, valid, canonical, boards_out, captured_any, status);
  }
  return fail(B2AZ_EINVAL, "unknown tafl game");
}

extern "C" int b2az_tafl_replay(int device, uint32_t game, uint32_t n, uint32_t max_len, uint32_t max_turns,
                                const uint16_t* moves, const uint32_t* lens, int8_t* boards, uint8_t* players,
                                uint32_t* turns, uint8_t* reps, uint8_t* terminal, uint32_t* n_valid, uint8_t* valid,
                                float* canonical, int32_t* status) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!moves || !lens || max_len == 0) return fail(B2AZ_EINVAL, "null argument");
  if (max_turns == 0 || max_turns > 65535u) return fail(B2AZ_EINVAL, "max_turns must fit uint16_t");
  switch (game) {
    case B2AZ_TAFL_BRANDUBH:
      return tafl_replay_impl<B2AZ_TAFL_BRANDUBH>(device, n, max_len, max_turns, moves, lens, boards, players, turns, reps,
                                                  terminal, n_valid, valid, canonical, status);
    case B2AZ_TAFL_OPENTAFL:
      return tafl_replay_impl<B2AZ_TAFL_OPENTAFL>(device, n, max_len, max_turns, moves, lens, boards, players, turns, reps,
                                                  terminal, n_valid, valid, canonical, status);
    case B2AZ_TAFL_TAWLBWRDD:
      return tafl_replay_impl<B2AZ_TAFL_TAWLBWRDD>(device, n, max_len, max_turns, moves, lens, boards, players, turns,
                                                   reps, terminal, n_valid, valid, canonical, status);
  }
  return fail(B2AZ_EINVAL, "unknown tafl game");
}
