// az_stargambit_kernels.h — Star Gambit on the device, ONE WARP PER GAME (included by az_engine.cu).
//
// Every lane keeps its own copy of the position (SGState, 200 B: registers + L1-resident local memory) and replays
// moves on it — uniform control flow, no broadcast, exactly the scalar rule code of az_stargambit.h. The lanes split
// whatever is data parallel:
//   repetition   the key history (HBM) is scanned 32 keys at a time (SGHistWarp)
//   legal moves  lane i evaluates unit i's ten action slots against the occupancy sets (built with warp OR
//                reductions), lanes 0-17 one deploy (type, facing) each; the bits land in a 1709-bit map in shared
//                memory and come out in ascending id order through a popcount prefix sum — the order
//                Node::add_children needs (mcts.cc:93-101)
//   canonical    lane c handles board cell c: one unit lookup, then the 32 / 36 plane values of that cell — every
//                plane row is written as a coalesced run of 4 B stores (24 KB per Unified position: the dominant
//                HBM term of this game)
// k_sg_replay: a batch of transcripts replayed from the start position (the game kernels' parity + throughput
// entry point: b2az_sg_replay / b2az_sg_replay_device). The wide-tree search uses the same device functions.
#pragma once

#include "az_stargambit.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace b2az {

constexpr int kSGMapWords = (kSGUnifiedMoves + 31) / 32;  // 54
constexpr int kSGMaxK = 128;                               // most legal actions of a position (10 ships x 9 + 18 + 1)

#ifndef B2AZ_HOST_EMU
struct SGHistWarp {  // `base` (read-only, e.g. a search root's history) followed by `keys` (appended to)
  const u64* base;
  u32 base_len;
  u64* keys;
  u32 len, cap, lane;
  bool overflow;
  __device__ __forceinline__ void clear() { base_len = 0; len = 0; }
  __device__ __forceinline__ int count(u64 k) const {
    u32 c = 0;
    for (u32 i = lane; i < base_len; i += 32u) c += base[i] == k ? 1u : 0u;
    for (u32 i = lane; i < len; i += 32u) c += keys[i] == k ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    return (int)c;
  }
  __device__ __forceinline__ int push_count(u64 k) {
    const int c = count(k);
    if (len < cap) { if (lane == 0) keys[len] = k; ++len; } else overflow = true;
    __syncwarp();
    return c + 1;
  }
};

struct SGWarpSmem {
  // The working position: ONE copy per warp. Every lane runs the scalar rule code on it redundantly — reads are
  // broadcasts, and the lanes of a converged warp store identical values (the control flow depends on the position
  // only; lane-dependent sections end in warp-synchronising primitives). A private copy per lane would be 32 x 200 B of
  // local memory per warp, which falls out of L1 at a few warps per SM (measured: 9.5 M simulations/s regardless of the
  // number of game slots).
  SGState st;
  u32 map[kSGMapWords];
  u16 moves[kSGMaxK];
  u8 cell_unit[13 * 13 + 3];
};
// cooperative copies between the warp's working position and a position record in global memory
__device__ __forceinline__ void sg_warp_load(SGState& dst, const SGState* src, u32 lane) {
  __syncwarp();
  for (u32 i = lane; i < (u32)(sizeof(SGState) / 4u); i += 32u) reinterpret_cast<u32*>(&dst)[i] = reinterpret_cast<const u32*>(src)[i];
  __syncwarp();
}
__device__ __forceinline__ void sg_warp_store(SGState* dst, const SGState& src, u32 lane) {
  __syncwarp();
  for (u32 i = lane; i < (u32)(sizeof(SGState) / 4u); i += 32u) reinterpret_cast<u32*>(dst)[i] = reinterpret_cast<const u32*>(&src)[i];
  __syncwarp();
}

// occupancy sets by warp reduction: lane i contributes unit i
__device__ __noinline__ void sg_warp_boards(const SGState& s, int side, u32 lane, SGBoards& b) {
  SGOcc mine;
  sg_occ_clear(mine);
  u32 pl = 0;
  if (lane < (u32)s.n_units && s.units[lane].hp > 0) {
    int hq[3], hr[3];
    const int n = sg_unit_hexes(s.units[lane], side, hq, hr);
    for (int j = 0; j < n; ++j) sg_occ_set(mine, hq[j], hr[j]);
    pl = s.units[lane].player & 1u;
  }
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    const u32 lo = (u32)mine.w[w], hi = (u32)(mine.w[w] >> 32);
    const u32 alo = __reduce_or_sync(0xFFFFFFFFu, lo), ahi = __reduce_or_sync(0xFFFFFFFFu, hi);
    const u32 olo = __reduce_or_sync(0xFFFFFFFFu, pl ? lo : 0u), ohi = __reduce_or_sync(0xFFFFFFFFu, pl ? hi : 0u);
    b.all.w[w] = ((u64)ahi << 32) | alo;
    b.pl[1].w[w] = ((u64)ohi << 32) | olo;
    b.pl[0].w[w] = b.all.w[w] & ~b.pl[1].w[w];
  }
}
struct SGAnyValidWarp {  // valid_moves().sum() != 0, one unit / one deploy per lane
  u32 lane;
  __device__ __noinline__ bool operator()(const SGState& s, const SGSpace& sp) const {
    if (s.over) return false;
    SGBoards b;
    sg_warp_boards(s, sp.side, lane, b);
    bool any = !sg_turn_one(s) && s.acted;
    if (!sg_turn_one(s) && lane < (u32)s.n_units) any = any || sg_unit_slots(s, b, (int)lane, sp.side) != 0;
    if (lane < 18u) any = any || sg_deploy_ok(s, b, sp.side, (int)lane / 6, (int)lane % 6);
    return __any_sync(0xFFFFFFFFu, any);
  }
};
// valid_moves() into sm.map (one bit per action id) and sm.moves (ascending ids); returns their number
__device__ __noinline__ u32 sg_warp_legal(const SGState& s, const SGSpace& sp, SGWarpSmem& sm, u32 lane) {
  for (u32 w = lane; w < (u32)kSGMapWords; w += 32u) sm.map[w] = 0;
  __syncwarp();
  if (!s.over) {
    SGBoards b;
    sg_warp_boards(s, sp.side, lane, b);
    const bool p1 = s.player == 1;
    if (!sg_turn_one(s) && lane < (u32)s.n_units) {
      const u32 m = sg_unit_slots(s, b, (int)lane, sp.side);
      if (m) {
        int row = s.units[lane].q + sp.side, col = s.units[lane].r + sp.side;
        if (p1) { row = sp.dim - 1 - row; col = sp.dim - 1 - col; }
        const u32 base = (u32)(((row + sp.off) * sp.udim + (col + sp.off)) * 10);
        const u64 bits = (u64)m << (base & 31u);  // ten bits: at most two words
        atomicOr(&sm.map[base >> 5], (u32)bits);
        if ((u32)(bits >> 32)) atomicOr(&sm.map[(base >> 5) + 1u], (u32)(bits >> 32));
      }
    }
    if (lane < 18u && sg_deploy_ok(s, b, sp.side, (int)lane / 6, (int)lane % 6)) {
      const int type = (int)lane / 6, f = (int)lane % 6;
      const u32 id = (u32)(sp.deploy_offset() + type * 6 + (p1 ? (f + 3) % 6 : f));
      atomicOr(&sm.map[id >> 5], 1u << (id & 31u));
    }
    if (lane == 0 && !sg_turn_one(s) && s.acted) {
      const u32 id = (u32)sp.end_turn();
      atomicOr(&sm.map[id >> 5], 1u << (id & 31u));
    }
  }
  __syncwarp();
  u32 total = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {  // 54 words in two rounds of 32
    const u32 w = 32u * r + lane;
    u32 bits = w < (u32)kSGMapWords ? sm.map[w] : 0u;
    const u32 cnt = (u32)__popc(bits);
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((int)lane >= o) incl += up;
    }
    u32 off = total + incl - cnt;
    for (; bits; bits &= bits - 1u, ++off)
      if (off < (u32)kSGMaxK) sm.moves[off] = (u16)(32u * w + (u32)(__ffs((int)bits) - 1));
    total += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  __syncwarp();
  return total;
}

// canonicalized() written by the warp: out[plane][udim][udim]
// (`rc` = how often the position's key occurs in its key history: position_history_ scanned by the caller)
__device__ __noinline__ void sg_warp_canon(const SGState& s, int rc, const SGSpace& sp, bool unified,
                                              SGWarpSmem& sm, u32 lane, float* out) {
  const int cells = sp.dim * sp.dim, ucells = sp.udim * sp.udim;
  for (int i = (int)lane; i < cells; i += 32) sm.cell_unit[i] = 0;
  __syncwarp();
  const bool p1 = s.player == 1;
  if (lane < (u32)s.n_units && s.units[lane].hp > 0) {
    int hq[3], hr[3];
    const int n = sg_unit_hexes(s.units[lane], sp.side, hq, hr);
    for (int j = 0; j < n; ++j) {
      const int q = p1 ? -hq[j] : hq[j], r = p1 ? -hr[j] : hr[j];
      if (q >= -sp.side && q <= sp.side && r >= -sp.side && r <= sp.side)
        sm.cell_unit[(q + sp.side) * sp.dim + (r + sp.side)] = (u8)(lane + 1u);
    }
  }
  // the broadcast planes' values (every lane computes the same ten numbers)
  float g[10];
  {
    g[0] = s.acted ? 1.0f : 0.0f;
    g[1] = rc == 0 ? 0.0f : rc == 1 ? 0.5f : 1.0f;
    const int my = s.player & 1, opp = 1 - my;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int st = sg_start(s.variant, t);
      g[2 + t] = st > 0 ? fdiv((float)s.reserves[my][t], (float)st) : 0.0f;
      g[5 + t] = st > 0 ? fdiv((float)s.reserves[opp][t], (float)st) : 0.0f;
    }
    g[8] = g[9] = 0.0f;
    for (int i = (int)s.n_units - 1; i >= 0; --i) {  // the FIRST alive portal of a player wins: walk backwards
      const SGUnit& u = s.units[i];
      if (u.type == SG_PORTAL && u.slot == 0 && u.hp > 0) g[u.player == my ? 8 : 9] = fdiv((float)u.hp, 5.0f);
    }
  }
  __syncwarp();
  const int planes = sp.planes(unified);
  for (int uc = (int)lane; uc < ucells; uc += 32) {
    const int row = uc / sp.udim - sp.off, col = uc % sp.udim - sp.off;
    const bool on = row >= 0 && row < sp.dim && col >= 0 && col < sp.dim && sg_inb(row - sp.side, col - sp.side, sp.side);
    u64 ones = 0;          // planes whose value is exactly 1
    float hp = 0.0f, mv = 0.0f;
    if (on) {
      ones = 1ULL;
      if (unified) ones |= 1ULL << (32 + s.variant);
      const int ui = sm.cell_unit[row * sp.dim + col];
      if (ui) {
        const SGUnit& u = s.units[ui - 1];
        ones |= 1ULL << (1 + (u.player == s.player ? 0 : 4) + u.type);
        hp = fdiv((float)u.hp, (float)sg_max_hp(u.type));
        if (u.type != SG_PORTAL) {
          ones |= 1ULL << (9 + (p1 ? (u.facing + 3) % 6 : (int)u.facing));
          mv = fdiv((float)u.moves_left, (float)sg_max_moves(u.type));
          const int aq = p1 ? -u.q : u.q, ar = p1 ? -u.r : u.r;
          if (aq + sp.side == row && ar + sp.side == col) {  // unfired cannons, on the anchor: forward, fl, fr, rl, rr
            const u32 nf = ~(u32)u.fired;
            u32 obs;
            if (u.type == SG_FIGHTER) obs = nf & 1u;
            else if (u.type == SG_CRUISER) obs = ((nf >> 1) & 1u) | ((nf & 1u) << 1) | (((nf >> 2) & 1u) << 2);
            else obs = (((nf >> 1) & 1u) << 1) | (((nf >> 2) & 1u) << 2) | ((nf & 1u) << 3) | (((nf >> 3) & 1u) << 4);
            ones |= (u64)obs << 17;
          }
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < kSGUnifiedPlanes; ++ch) {
      if (ch >= planes) break;
      float v = ((ones >> ch) & 1ULL) ? 1.0f : 0.0f;
      if (ch == 15) v = hp;
      if (ch == 16) v = mv;
      if (ch >= 22 && ch < 32) v = on ? g[ch - 22] : 0.0f;
      out[ch * ucells + uc] = v;
    }
  }
  __syncwarp();
}

struct SGReplayArgs {
  u32 n, max_len, game, hist_cap;
  const u16* moves;
  const u32* lens;
  u64* hist;       // [n][hist_cap]
  u8* states;      // [rows][sizeof(SGState)] or null
  u8* terminal;    // [rows]
  u32* n_valid;    // [rows]
  u8* valid;       // [rows][A]
  float* canonical;  // [rows][P][D][D]
  i32* status;     // [n]
};
__device__ __forceinline__ bool sg_game_unified(u32 game) { return game >= 20u; }
__device__ __forceinline__ int sg_game_variant(u32 game) { return (int)(game % 10u); }

__global__ void __launch_bounds__(128) k_sg_replay(SGReplayArgs a) {
  __shared__ SGWarpSmem smem[4];
  const u32 lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  SGWarpSmem& sm = smem[wib];
  const bool unified = sg_game_unified(a.game);
  const int variant = sg_game_variant(a.game);
  const SGSpace sp = sg_space(variant, unified);
  const u32 A = (u32)sp.num_moves(), C = (u32)(sp.planes(unified) * sp.udim * sp.udim);
  for (u32 gi = GLOBAL_TID >> 5; gi < a.n; gi += GLOBAL_NT >> 5) {
    SGState& s = sm.st;
    __syncwarp();
    sg_init(s, variant);
    __syncwarp();
    SGHistWarp hist;
    hist.base = nullptr; hist.base_len = 0; hist.keys = a.hist + (size_t)gi * a.hist_cap; hist.len = 0; hist.cap = a.hist_cap;
    hist.lane = lane; hist.overflow = false;
    hist.push_count(sg_position_key(s));
    const u32 len = a.lens[gi];
    for (u32 k = 0; k <= len; ++k) {
      if (k > 0) {
        const bool ok = sg_play(s, hist, sp, (u32)a.moves[(size_t)gi * a.max_len + (k - 1u)], SGAnyValidWarp{lane});
        __syncwarp();
        if (!ok) {
          if (lane == 0 && a.status) a.status[gi] = B2AZ_EMOVE;
          break;
        }
      }
      const size_t row = (size_t)gi * (a.max_len + 1u) + k;
      if (a.states)
        for (u32 i = lane; i < (u32)sizeof(SGState); i += 32u) a.states[row * sizeof(SGState) + i] = ((const u8*)&s)[i];
      if (a.terminal && lane == 0) a.terminal[row] = (u8)sg_terminal(s);
      if (a.n_valid || a.valid) {
        const u32 nv = sg_warp_legal(s, sp, sm, lane);
        if (a.n_valid && lane == 0) a.n_valid[row] = nv;
        if (a.valid)
          for (u32 m = lane; m < A; m += 32u) a.valid[row * A + m] = (u8)((sm.map[m >> 5] >> (m & 31u)) & 1u);
        __syncwarp();
      }
      if (a.canonical) sg_warp_canon(s, hist.count(sg_position_key(s)), sp, unified, sm, lane, a.canonical + row * C);
    }
    if (hist.overflow && lane == 0 && a.status) a.status[gi] = B2AZ_ENOMEM;
  }
}
// GameState::symmetries for Star Gambit (star_gambit_gs.cc:1671-1805, Unified 2623-2727): every sample followed by its
// NW-axis mirror image, gathered through the inverse index maps of az_stargambit.h; OUT = float or __half (the fp16 form
// game_runner.save_compressed stores, game_runner.py:200-210)
template <class OUT>
__device__ __forceinline__ OUT sg_out_cast(float x);
template <>
__device__ __forceinline__ float sg_out_cast<float>(float x) { return x; }
template <>
__device__ __forceinline__ __half sg_out_cast<__half>(float x) { return __float2half_rn(x); }
template <class OUT>
__global__ void k_sg_symmetries(u32 n, u32 planes, u32 dim, const float* __restrict__ canon, const float* __restrict__ v,
                                const float* __restrict__ pi, OUT* __restrict__ canon_out, OUT* __restrict__ v_out,
                                OUT* __restrict__ pi_out) {
  const u32 side = (dim - 1u) / 2u;
  const size_t C = (size_t)planes * dim * dim, A = (size_t)dim * dim * 10u + 19u, per = C + A + 3;
  const size_t total = (size_t)n * 2u * per;
  for (size_t i = GLOBAL_TID; i < total; i += GLOBAL_NT) {
    const size_t row = i / per, e = i % per;  // row = sample * 2 + sym
    const u32 sample = (u32)(row / 2u), sym = (u32)(row % 2u);
    if (e < C) {
      const int src = sym ? sg_mirror_canon_src((int)dim, (int)side, (int)planes, (int)e) : (int)e;
      canon_out[row * C + e] = sg_out_cast<OUT>(src < 0 ? 0.0f : canon[(size_t)sample * C + (size_t)src]);
    } else if (e < C + A) {
      const int mv = (int)(e - C);
      const int src = sym ? sg_mirror_pi_src((int)dim, (int)side, mv) : mv;
      pi_out[row * A + (size_t)mv] = sg_out_cast<OUT>(src < 0 ? 0.0f : pi[(size_t)sample * A + (size_t)src]);
    } else {
      const u32 j = (u32)(e - C - A);
      v_out[row * 3 + j] = sg_out_cast<OUT>(v[(size_t)sample * 3 + j]);
    }
  }
}
#endif  // !B2AZ_HOST_EMU

}  // namespace b2az

extern "C" int b2az_sg_symmetries(int device, uint32_t game, uint32_t n, const float* canon, const float* v, const float* pi,
                                  void* canon_out, void* v_out, void* pi_out, int device_pointers, int fp16_out, void* stream) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!canon || !v || !pi || !canon_out || !v_out || !pi_out) return fail(B2AZ_EINVAL, "null argument");
  if (!((game >= 10 && game <= 13) || (game >= 20 && game <= 24))) return fail(B2AZ_EINVAL, "unknown Star Gambit game");
#ifdef B2AZ_HOST_EMU
  (void)device; (void)device_pointers; (void)fp16_out; (void)stream;
  return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  const bool unified = game >= 20;
  const SGSpace sp = sg_space(game == 24u ? B2AZ_SG_BATTLE : (int)(game % 10u), unified);
  const u32 planes = (u32)sp.planes(unified), dim = (u32)sp.udim;
  const size_t C = (size_t)planes * dim * dim, A = (size_t)sp.num_moves(), osz = fp16_out ? 2 : 4;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float *dc = canon, *dv = v, *dp = pi;
  void *oc = canon_out, *ov = v_out, *op = pi_out;
  float *tc = nullptr, *tv = nullptr, *tp = nullptr;
  void *toc = nullptr, *tov = nullptr, *top = nullptr;
  int rc = 0;
  if (!device_pointers) {
    auto al = [&](void** p, size_t bytes) { if (!rc && cudaMalloc(p, bytes) != cudaSuccess) rc = fail(B2AZ_ENOMEM, "b2az_sg_symmetries: cudaMalloc failed"); };
    al((void**)&tc, n * C * 4); al((void**)&tv, (size_t)n * 12); al((void**)&tp, n * A * 4);
    al(&toc, (size_t)n * 2 * C * osz); al(&tov, (size_t)n * 6 * osz); al(&top, (size_t)n * 2 * A * osz);
    if (!rc) {
      cudaMemcpyAsync(tc, canon, n * C * 4, cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(tv, v, (size_t)n * 12, cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(tp, pi, n * A * 4, cudaMemcpyHostToDevice, s);
    }
    dc = tc; dv = tv; dp = tp; oc = toc; ov = tov; op = top;
  }
  if (!rc) {
    if (fp16_out) k_sg_symmetries<__half><<<148 * 8, 256, 0, s>>>(n, planes, dim, dc, dv, dp, (__half*)oc, (__half*)ov, (__half*)op);
    else k_sg_symmetries<float><<<148 * 8, 256, 0, s>>>(n, planes, dim, dc, dv, dp, (float*)oc, (float*)ov, (float*)op);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_sg_symmetries launch failed");
  }
  if (!device_pointers) {
    if (!rc) {
      cudaMemcpyAsync(canon_out, toc, (size_t)n * 2 * C * osz, cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(v_out, tov, (size_t)n * 6 * osz, cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(pi_out, top, (size_t)n * 2 * A * osz, cudaMemcpyDeviceToHost, s);
      const cudaError_t err = cudaStreamSynchronize(s);
      if (err != cudaSuccess) rc = fail(B2AZ_ECUDA, std::string("k_sg_symmetries: ") + cudaGetErrorString(err));
    }
    cudaFree(tc); cudaFree(tv); cudaFree(tp); cudaFree(toc); cudaFree(tov); cudaFree(top);
  }
  return rc;
#endif
}

extern "C" int b2az_sg_replay_device(uint32_t game, uint32_t n, uint32_t max_len, const uint16_t* moves_dev,
                                     const uint32_t* lens_dev, void* hist_dev, uint32_t hist_cap, uint8_t* states_dev,
                                     uint8_t* terminal_dev, uint32_t* n_valid_dev, uint8_t* valid_dev,
                                     float* canonical_dev, int32_t* status_dev, void* stream) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!moves_dev || !lens_dev || !hist_dev || max_len == 0 || hist_cap < 2) return fail(B2AZ_EINVAL, "null argument");
  if (!((game >= 10 && game <= 13) || (game >= 20 && game <= 23))) return fail(B2AZ_EINVAL, "unknown Star Gambit game");
#ifdef B2AZ_HOST_EMU
  (void)states_dev; (void)terminal_dev; (void)n_valid_dev; (void)valid_dev; (void)canonical_dev; (void)status_dev; (void)stream;
  return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
#else
  SGReplayArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.max_len = max_len; a.game = game; a.hist_cap = hist_cap;
  a.moves = moves_dev; a.lens = lens_dev; a.hist = static_cast<u64*>(hist_dev);
  a.states = states_dev; a.terminal = terminal_dev; a.n_valid = n_valid_dev; a.valid = valid_dev;
  a.canonical = canonical_dev; a.status = status_dev;
  const unsigned ctas = std::max(1u, std::min((n + 3u) / 4u, 148u * 8u));
  k_sg_replay<<<ctas, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  CUDA_TRY(cudaGetLastError());
  return 0;
#endif
}

extern "C" int b2az_sg_replay(int device, uint32_t game, uint32_t n, uint32_t max_len, const uint16_t* moves,
                              const uint32_t* lens, uint8_t* states, uint8_t* terminal, uint32_t* n_valid,
                              uint8_t* valid, float* canonical, int32_t* status) {
  using namespace b2az;
  if (n == 0) return 0;
  if (!moves || !lens || max_len == 0) return fail(B2AZ_EINVAL, "null argument");
  if (!((game >= 10 && game <= 13) || (game >= 20 && game <= 23))) return fail(B2AZ_EINVAL, "unknown Star Gambit game");
#ifdef B2AZ_HOST_EMU
  (void)device; (void)states; (void)terminal; (void)n_valid; (void)valid; (void)canonical; (void)status;
  return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  const bool unified = game >= 20;
  const SGSpace sp = sg_space((int)(game % 10u), unified);
  const size_t rows = (size_t)n * (max_len + 1u), A = (size_t)sp.num_moves(), C = (size_t)sp.planes(unified) * sp.udim * sp.udim;
  const uint32_t hist_cap = max_len + 2u;
  u16* dm = nullptr; u32 *dl = nullptr, *dnv = nullptr; u64* dh = nullptr; u8 *dst = nullptr, *dt = nullptr, *dv = nullptr;
  float* dc = nullptr; i32* ds = nullptr;
  int rc = 0;
  auto al = [&](void** p, size_t bytes) { if (!rc && cudaMalloc(p, bytes) != cudaSuccess) rc = fail(B2AZ_ENOMEM, "b2az_sg_replay: cudaMalloc failed"); };
  al((void**)&dm, (size_t)n * max_len * 2); al((void**)&dl, (size_t)n * 4); al((void**)&dh, (size_t)n * hist_cap * 8);
  al((void**)&ds, (size_t)n * 4);
  if (states) al((void**)&dst, rows * sizeof(SGState));
  if (terminal) al((void**)&dt, rows);
  if (n_valid) al((void**)&dnv, rows * 4);
  if (valid) al((void**)&dv, rows * A);
  if (canonical) al((void**)&dc, rows * C * 4);
  if (!rc) {
    cudaMemcpy(dm, moves, (size_t)n * max_len * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dl, lens, (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemset(ds, 0, (size_t)n * 4);
    rc = b2az_sg_replay_device(game, n, max_len, dm, dl, dh, hist_cap, dst, dt, dnv, dv, dc, ds, nullptr);
    if (!rc) {
      const cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) rc = fail(B2AZ_ECUDA, std::string("k_sg_replay: ") + cudaGetErrorString(err));
    }
  }
  if (!rc) {
    if (states) cudaMemcpy(states, dst, rows * sizeof(SGState), cudaMemcpyDeviceToHost);
    if (terminal) cudaMemcpy(terminal, dt, rows, cudaMemcpyDeviceToHost);
    if (n_valid) cudaMemcpy(n_valid, dnv, rows * 4, cudaMemcpyDeviceToHost);
    if (valid) cudaMemcpy(valid, dv, rows * A, cudaMemcpyDeviceToHost);
    if (canonical) cudaMemcpy(canonical, dc, rows * C * 4, cudaMemcpyDeviceToHost);
    if (status) cudaMemcpy(status, ds, (size_t)n * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(dm); cudaFree(dl); cudaFree(dh); cudaFree(ds); cudaFree(dst); cudaFree(dt); cudaFree(dnv); cudaFree(dv); cudaFree(dc);
  return rc;
#endif
}
