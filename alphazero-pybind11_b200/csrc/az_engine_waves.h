// az_engine_waves.h — k_step_w: the fused self-play step as a persistent kernel scheduled in WAVES.
//
// Same pieces and the same per-game state machine as run_flat() (az_engine_logic.h) — hence the same results, game by
// game — but the lanes of a warp are not married to 32 fixed games. A CTA (one per SM) owns <= kQGames game slots whose
// working state lives in SHARED memory (QGame, az_engine_queue.h) and keeps three plain LISTS of game ids:
//     LEVEL     games whose descent goes on: one PUCT level (one 160 B block burst + 7 scores); a game whose descent
//               ends there is expanded in the same piece (descent_finish)
//     BOUNDARY  games whose leaf is answered: process_result (priors, backprop) and the start of the next descent
//     MOVE      games whose search budget is reached (play_move: acting rule, sample, two re-roots, game end)
// A round cuts the three lists into chunks of 32 ids, one chunk per warp; every piece appends each of its games to the
// list of the piece it needs next (ballot ranks + ONE shared atomicAdd per warp and target list); one CTA barrier; next
// round. Both common pieces are one dependent HBM round trip plus a few hundred instructions long, so the warps of a
// round finish together, and every warp-level instruction works for ~29 games that need the same code (the
// thread-per-game kernel: 9.6 of 32 lanes, profiles/r49_k_step_ncu_summary.json). Nobody polls and nothing is handed
// over through flags: between two barriers a list is either read or appended to, never both (three buffers per list:
// input of this round, output of this round, and the one thread 0 clears for the next round).
// The road here — a dynamic shared-memory queue (12 k cycles of pop / push hand-shakes per batch), rounds with a long
// leaf piece next to K level sub-phases (38 % of the warp time at barriers) — is recorded in profiles/r2_step_kernel_log.md.
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

#if defined(B2AZ_W_PROF)
// experiment build: clock64 sums (lane 0 of every warp, shared accumulators, one global add per launch)
//   [0] level chunks  [1] boundary chunks  [2] move chunks  [3] waiting at the barrier  [4] rounds (warp 0)
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
  u16 list[3][3][kQGames];  // [kind: Q_LEVEL, Q_PR, Q_MOVE][buffer][entry]
  u32 count[3][3];
};

// every lane with pred appends `id` to a list (count in shared memory): ballot ranks, one atomicAdd per warp
__device__ __forceinline__ void w_append(u16* list, u32* count, bool pred, u32 id, u32 lane) {
  const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  if (m == 0u) return;
  const int leader = __ffs(m) - 1;
  u32 base = 0;
  if ((int)lane == leader) base = atomicAdd(count, (u32)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (pred) list[base + (u32)__popc(m & ((1u << lane) - 1u))] = (u16)id;
}

// One chunk of 32 ids of list `kind` (buffer `in`): run the piece for every game, hand each game to its next list
// (buffer `out`).
template <bool GB>
__device__ __forceinline__ void w_chunk(const EngineView& E, WShared& S, u32 kind, u32 in, u32 out, u32 count, u32 chunk,
                                        u32 g0, u32 lane) {
  const u32 i = chunk * 32u + lane;
  const bool have = i < count;
  const u32 id = have ? (u32)S.list[kind][in][i] : 0u;
  u32 ns = 0xFFu;
  if (have) {
    QGame& q = S.game[id];
    const u32 g = g0 + id;
    if (kind == Q_LEVEL) {
      ns = q_level<GB>(E, g, q);
      if (ns == Q_LEAF) ns = q_finish<GB>(E, g, q);  // the descent ended: expand right here
    } else if (kind == Q_PR) {
      ns = q_boundary<GB>(E, g, q);
    } else {
      ns = q_move<GB>(E, g, q);
      if (ns == Q_LEAF) ns = q_finish<GB>(E, g, q);
    }
  }
  __syncwarp();
#pragma unroll
  for (u32 t = 0; t < 3u; ++t) w_append(S.list[t][out], &S.count[t][out], ns == t, id, lane);
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
    if (tid < 9u) (&S.count[0][0])[tid] = 0;
    __syncthreads();
    // load the group's state; every active game starts at a step boundary (BOUNDARY list, buffer 0)
    for (u32 i0 = 0; i0 < ng; i0 += blockDim.x) {
      const u32 i = i0 + tid;
      bool active = false;
      if (i < ng) {
        q_load_game(E, g0 + i, S.game[i], n_steps);
        active = S.game[i].gs.active != 0;
      }
      w_append(S.list[Q_PR][0], &S.count[Q_PR][0], active, i, lane);
    }
    __syncthreads();
    u32 move_wait = 0;
    for (u32 round = 0; round < (1u << 26); ++round) {  // bounded: a scheduling bug must not hang the GPU
      const u32 in = round % 3u, out = (round + 1u) % 3u, clr = (round + 2u) % 3u;
      // the input lists are not written to during this round
      const u32 nL = S.count[Q_LEVEL][in], nP = S.count[Q_PR][in], nM = S.count[Q_MOVE][in];
      if (nL + nP + nM == 0u) break;  // every game has finished its steps of this launch
      // buffer `clr` was the input of the previous round; it becomes the output of the next one
      if (tid < 3u) S.count[tid][clr] = 0;
      // A MOVE chunk is long (acting rule, sample, two re-roots, maybe a compaction) and everybody waits for it at the
      // barrier, so movers are collected until a chunk is full — or they have waited B2AZ_W_MOVE_WAIT rounds, or nothing
      // else is left to do. The waiting games are passed on from list to list; no lane idles for them.
      const bool do_moves = nM >= 32u || (nM > 0u && (move_wait >= (u32)B2AZ_W_MOVE_WAIT || nL + nP == 0u));
      move_wait = (nM > 0u && !do_moves) ? move_wait + 1u : 0u;
      const u32 cM = do_moves ? (nM + 31u) >> 5 : 0u, cP = (nP + 31u) >> 5, cL = (nL + 31u) >> 5;
      const u32 total = cM + cP + cL;
      const long long t0 = WPROF_CLK();
      if (nM > 0u && !do_moves && warp == W - 1u) {  // pass the waiting movers on (the last warp gets the fewest chunks)
        for (u32 i0 = 0; i0 < nM; i0 += 32u) {
          const u32 i = i0 + lane;
          w_append(S.list[Q_MOVE][out], &S.count[Q_MOVE][out], i < nM, i < nM ? (u32)S.list[Q_MOVE][in][i] : 0u, lane);
        }
      }
      for (u32 c = warp; c < total; c += W) {  // the longest pieces first
        const long long c0 = WPROF_CLK();
        if (c < cM) {
          w_chunk<GB>(E, S, Q_MOVE, in, out, nM, c, g0, lane);
          WPROF_ADD(2, WPROF_CLK() - c0); WPROF_ADD(7, 1); WPROF_ADD(10, (nM - c * 32u) < 32u ? nM - c * 32u : 32u);
        } else if (c < cM + cP) {
          const u32 cc = c - cM;
          w_chunk<GB>(E, S, Q_PR, in, out, nP, cc, g0, lane);
          WPROF_ADD(1, WPROF_CLK() - c0); WPROF_ADD(6, 1); WPROF_ADD(9, (nP - cc * 32u) < 32u ? nP - cc * 32u : 32u);
        } else {
          const u32 cc = c - cM - cP;
          w_chunk<GB>(E, S, Q_LEVEL, in, out, nL, cc, g0, lane);
          WPROF_ADD(0, WPROF_CLK() - c0); WPROF_ADD(5, 1); WPROF_ADD(8, (nL - cc * 32u) < 32u ? nL - cc * 32u : 32u);
        }
      }
      const long long t1 = WPROF_CLK();
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
