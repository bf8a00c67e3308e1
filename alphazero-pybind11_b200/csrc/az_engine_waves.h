// az_engine_waves.h — k_step_w: the fused self-play step as a persistent kernel scheduled in WAVES.
//
// Same pieces and the same per-game state machine as the queue kernel (az_engine_queue.h: q_level / q_leaf / q_move on
// a game's shared-memory record), but no dynamic queue: a CTA (one per SM, <= kQGames games) keeps three plain LISTS of
// game ids — games whose descent goes on (LEVEL), games at a step boundary (LEAF), games whose search budget is reached
// (MOVE) — and works in rounds separated by CTA barriers:
//
//   round:  the LEAF and MOVE lists are cut into chunks of 32 ids and handed to the first warps ("heavy" warps: a leaf
//           chunk is ~3x a level chunk); the remaining warps take the LEVEL list and run K level sub-phases over it
//           (a named barrier among the level warps between sub-phases), so both sides finish at about the same time.
//           Every piece appends each of its games to the list of the piece it needs next (ballot ranks + ONE shared
//           atomicAdd per warp and target list).
//   barrier, swap lists, next round.
//
// Every warp-level instruction of a chunk works for 32 games that need the same code; nobody polls, nothing is
// handed over through flags: between two barriers a list is either read or appended to, never both.
// Measured on the way here (profiles/r2_queue_vs_waves.md): the queue kernel spent 12 k cycles per batch in pop /
// push hand-shakes and idled half of its warps, for 3.6 k (level) / 10.8 k (leaf) cycles of work per batch.
#pragma once

#include "az_engine_queue.h"

#ifndef B2AZ_HOST_EMU

namespace b2az {

#ifndef B2AZ_W_WARPS
#define B2AZ_W_WARPS 16
#endif
#ifndef B2AZ_W_MOVE_WAIT
#define B2AZ_W_MOVE_WAIT 16  // rounds a game whose search is finished may wait for more movers to share its chunk
#endif
#ifndef B2AZ_W_K
#define B2AZ_W_K 3        // level sub-phases per round while heavy chunks are being worked on
#endif
constexpr int kWLevBufs = B2AZ_W_K + 1;

#if defined(B2AZ_W_PROF)
// experiment build: clock64 sums (lane 0 of every warp, shared accumulators, one global add per launch)
//   [0] level chunks  [1] leaf chunks  [2] move chunks  [3] waiting at barriers  [4] rounds (warp 0)
//   [5..7] chunks per kind  [8..10] games per kind
__device__ unsigned long long g_wprof[16];
__shared__ unsigned long long s_wprof[B2AZ_W_WARPS][16];
#define WPROF_ADD(i, v) do { if (lane == 0u) s_wprof[threadIdx.x >> 5][i] += (unsigned long long)(v); } while (0)
#define WPROF_CLK() clock64()
#else
#define WPROF_ADD(i, v) do { } while (0)
#define WPROF_CLK() 0ll
#endif

struct WShared {
  QGame game[kQGames];
  u16 lev[kWLevBufs][kQGames];  // LEVEL lists: this round's input, K - 1 scratch lists, next round's input
  u16 leaf[2][kQGames];
  u16 mov[2][kQGames];
  u32 n_lev[kWLevBufs];
  u32 n_leaf[2];
  u32 n_mov[2];
  u32 done;
};

// every lane with pred appends `id` to list (count in shared memory): ballot ranks, one atomicAdd per warp
__device__ __forceinline__ void w_append(u16* list, u32* count, bool pred, u32 id, u32 lane) {
  const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  if (m == 0u) return;
  const int leader = __ffs(m) - 1;
  u32 base = 0;
  if ((int)lane == leader) base = atomicAdd(count, (u32)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (pred) list[base + (u32)__popc(m & ((1u << lane) - 1u))] = (u16)id;
}

// One chunk of 32 ids of `list`: run piece `kind` for every game and hand each game to its next list.
template <bool GB>
__device__ __forceinline__ void w_chunk(const EngineView& E, WShared& S, u32 kind, const u16* list, u32 count, u32 chunk,
                                        u32 lev_out, u32 leaf_out, u32 mov_out, u32 g0, u32 lane) {
  const u32 i = chunk * 32u + lane;
  const bool have = i < count;
  const u32 id = have ? (u32)list[i] : 0u;
  u32 ns = 0xFFu;
  if (have) {
    QGame& q = S.game[id];
    const u32 g = g0 + id;
    if (kind == Q_LEVEL) ns = q_level<GB>(E, g, q);
    else if (kind == Q_LEAF) ns = q_leaf<GB>(E, g, q);
    else ns = q_move<GB>(E, g, q);
  }
  __syncwarp();
  w_append(S.lev[lev_out], &S.n_lev[lev_out], ns == Q_LEVEL, id, lane);
  w_append(S.leaf[leaf_out], &S.n_leaf[leaf_out], ns == Q_LEAF, id, lane);
  w_append(S.mov[mov_out], &S.n_mov[mov_out], ns == Q_MOVE, id, lane);
  const unsigned md = __ballot_sync(0xFFFFFFFFu, ns == Q_DONE);
  if (md && lane == 0u) atomicAdd(&S.done, (u32)__popc(md));
}

__device__ __forceinline__ void w_bar_level(u32 threads) {  // named barrier 1: the level warps of this round
  asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
}

template <bool GB>
__global__ void __launch_bounds__(B2AZ_W_WARPS * 32, 1) k_step_w(EngineView E, u32 n_steps, u32 games_per_group, u32 n_groups) {
  extern __shared__ __align__(16) unsigned char w_smem[];
  WShared& S = *reinterpret_cast<WShared*>(w_smem);
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr u32 W = B2AZ_W_WARPS;
#if defined(B2AZ_W_PROF)
  if (lane < 16u) s_wprof[warp][lane] = 0;
  __syncwarp();
#endif
  for (u32 grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const u32 g0 = grp * games_per_group;
    const u32 ng = (g0 >= E.G) ? 0u : (E.G - g0 < games_per_group ? E.G - g0 : games_per_group);
    if (tid < (u32)kWLevBufs) S.n_lev[tid] = 0;
    if (tid < 2u) { S.n_leaf[tid] = 0; S.n_mov[tid] = 0; }
    if (tid == 0u) S.done = 0;
    __syncthreads();
    // load the group's state; every active game starts at a step boundary (LEAF list 0)
    for (u32 i0 = 0; i0 < ng; i0 += blockDim.x) {
      const u32 i = i0 + tid;
      bool active = false;
      if (i < ng) {
        q_load_game(E, g0 + i, S.game[i], n_steps);
        active = S.game[i].gs.active != 0;
      }
      w_append(S.leaf[0], &S.n_leaf[0], active, i, lane);
      const unsigned md = __ballot_sync(0xFFFFFFFFu, i < ng && !active);
      if (md && lane == 0u) atomicAdd(&S.done, (u32)__popc(md));
    }
    __syncthreads();
    u32 done_snap = S.done;  // read between two barriers: warps that are already working on a round add to it
    __syncthreads();
    u32 cur = 0, fcur = 0, mcur = 0, move_wait = 0;
    for (u32 round = 0; round < (1u << 26); ++round) {  // bounded: a scheduling bug must not hang the GPU
      const u32 nL = S.n_lev[cur], nF = S.n_leaf[fcur];  // not written to during this round
      const u32 nM_all = S.n_mov[mcur];
      if (done_snap >= ng) break;
      if (nL + nF + nM_all == 0u) {  // cannot happen: every running game is in exactly one list
        if (tid == 0u) at_or(&E.glob->error, B2AZ_DEVERR_QUEUE);
        break;
      }
      // A MOVE chunk is long (acting rule, sample, two re-roots, maybe a compaction) and everybody waits for it at
      // the barrier, so movers are collected until a chunk is full — or they have waited B2AZ_W_MOVE_WAIT rounds, or
      // nothing else is left to do. The waiting games simply sit in the list; no lane idles for them.
      const bool do_moves = nM_all >= 32u || (nM_all > 0u && (move_wait >= (u32)B2AZ_W_MOVE_WAIT || nL + nF == 0u));
      const u32 nM = do_moves ? nM_all : 0u;
      const u32 mout = do_moves ? (mcur ^ 1u) : mcur;  // where this round's new movers go
      move_wait = (nM_all > 0u && !do_moves) ? move_wait + 1u : 0u;
      const u32 cL = (nL + 31u) >> 5, cF = (nF + 31u) >> 5, cM = (nM + 31u) >> 5;
      const u32 H = cF + cM;
      const u32 Wh = (cL == 0u) ? (H < W ? H : W) : (H < W - 1u ? H : W - 1u);
      const u32 Wl = (cL == 0u) ? 0u : (cL < W - Wh ? cL : W - Wh);
      const u32 K = (H == 0u) ? 1u : (u32)B2AZ_W_K;
      const u32 X = (cur + K) % (u32)kWLevBufs;  // next round's LEVEL list (its count is 0: reset at the end of a round)
      const long long t0 = WPROF_CLK();
      if (warp < Wh) {
        // heavy side: MOVE chunks first (the longest), then LEAF chunks
        for (u32 c = warp; c < H; c += Wh) {
          const long long c0 = WPROF_CLK();
          if (c < cM) {
            w_chunk<GB>(E, S, Q_MOVE, S.mov[mcur], nM, c, X, fcur ^ 1u, mout, g0, lane);
            WPROF_ADD(2, WPROF_CLK() - c0); WPROF_ADD(7, 1); WPROF_ADD(10, (nM - c * 32u) < 32u ? nM - c * 32u : 32u);
          } else {
            w_chunk<GB>(E, S, Q_LEAF, S.leaf[fcur], nF, c - cM, X, fcur ^ 1u, mout, g0, lane);
            WPROF_ADD(1, WPROF_CLK() - c0); WPROF_ADD(6, 1);
            WPROF_ADD(9, (nF - (c - cM) * 32u) < 32u ? nF - (c - cM) * 32u : 32u);
          }
        }
      } else if (warp < Wh + Wl) {
        const u32 idx = warp - Wh;
        for (u32 k = 0; k < K; ++k) {
          const u32 rd = (cur + k) % (u32)kWLevBufs, wr = (cur + k + 1u) % (u32)kWLevBufs;  // wr == X in the last one
          const u32 n = S.n_lev[rd];
          if (n == 0u) break;  // uniform over the level warps: they all read the count after the same barrier
          const u32 chunks = (n + 31u) >> 5;
          for (u32 c = idx; c < chunks; c += Wl) {
            const long long c0 = WPROF_CLK();
            // a game that goes on descending after the LAST sub-phase joins next round's list X, like the games the
            // heavy side starts on a new descent
            w_chunk<GB>(E, S, Q_LEVEL, S.lev[rd], n, c, (k + 1u == K) ? X : wr, fcur ^ 1u, mout, g0, lane);
            WPROF_ADD(0, WPROF_CLK() - c0); WPROF_ADD(5, 1); WPROF_ADD(8, (n - c * 32u) < 32u ? n - c * 32u : 32u);
          }
          if (k + 1u < K) w_bar_level(Wl * 32u);
        }
      }
      const long long t1 = WPROF_CLK();
      __syncthreads();
      // the lists consumed in this round are empty again; the scratch LEVEL lists too
      if (tid < (u32)kWLevBufs && tid != X) S.n_lev[tid] = 0;
      if (tid == 0u) {
        S.n_leaf[fcur] = 0;
        if (do_moves) S.n_mov[mcur] = 0;
      }
      done_snap = S.done;
      cur = X; fcur ^= 1u;
      if (do_moves) mcur ^= 1u;
      __syncthreads();
      WPROF_ADD(3, WPROF_CLK() - t1);
      if (warp == 0u) WPROF_ADD(4, 1);
      (void)t0;
    }
    __syncthreads();
    for (u32 i = tid; i < ng; i += blockDim.x) q_store_game(E, g0 + i, S.game[i]);
    __syncthreads();
  }
#if defined(B2AZ_W_PROF)
  __syncwarp();
  if (lane < 16u) atomicAdd(&g_wprof[lane], s_wprof[warp][lane]);
#endif
}

}  // namespace b2az

#endif  // !B2AZ_HOST_EMU
