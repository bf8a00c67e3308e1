// az_engine_queue.h — k_step_q: the fused self-play step as a PERSISTENT kernel with shared-memory work queues.
//
// Same per-game state machine as run_flat() (az_engine_logic.h) — hence the same results, game by game — but
// the lanes of a warp are no longer married to 32 fixed games. One CTA per SM owns a group of <= kQGames game
// slots whose working state (GameSlot, the mover's TreeHdr, the Descent, the selection path) lives in SHARED
// memory for the whole launch, and three queues of game ids say what each game needs next:
//     LEVEL  one PUCT level   (descent_level: one 160 B block burst + 7 scores)
//     LEAF   the step boundary (descent_finish -> [leaf batch] -> process_result -> descent_begin)
//     MOVE   PlayManager's per-move work (play_move: acting rule, sample, re-root both trees, game end)
// A warp pops up to 32 ids of ONE queue, runs that piece for all of them, and pushes every game to the queue of
// its next piece (ballot ranks + one shared atomicAdd per target queue). Every warp-level instruction therefore
// works for ~32 games doing the same thing, where the thread-per-game kernel ran the level code for the ~24
// descending lanes AND the leaf code for the ~8 lanes at a step boundary in every iteration
// (profiles/r49_k_step_ncu_summary.json: 9.6 of 32 lanes active per issued instruction).
// This is the device form of the reference's awaiting_mcts_ queue (play_manager.h:66-72, concurrent_queue.h):
// there the queue decouples games from worker THREADS, here it decouples them from SIMT lanes.
//
// Per-game order of operations is untouched (a game is in exactly one queue, or in exactly one warp's hands), so
// everything that is bit-exact in B2AZ_RNG_PER_GAME mode stays bit-exact.
#pragma once

#include "az_engine_logic.h"

#ifndef B2AZ_HOST_EMU

namespace b2az {

#define B2AZ_Q_WARPS_DEFAULT_MAX 32
#ifndef B2AZ_Q_WARPS
#define B2AZ_Q_WARPS 16   // warps per CTA; the pop policy below limits how many of them hold a batch at a time
#endif
#ifndef B2AZ_Q_NAP_MAX
#define B2AZ_Q_NAP_MAX 2048  // longest sleep (ns) of a warp that finds no batch
#endif
#ifndef B2AZ_Q_MIN
#define B2AZ_Q_MIN 32     // batch-size floor of a full group (kQGames games); smaller groups scale it down (QShared::qmin)
#endif

#if defined(B2AZ_Q_PROF)
// experiment build: where the warps' time goes (clock64 sums over all warps of all CTAs; accumulated in shared
// memory per warp and added to the global totals once per launch, so the accounting does not perturb the run)
//   [0..2] cycles in LEVEL / LEAF / MOVE work   [3] cycles in failed pops + sleeping   [4] cycles in successful pop + push
//   [5..7] batches per queue   [8..10] games per queue   [11] failed pops
//   [12] pop: decision loop   [13] pop: slot reads + fence   [14] push: fence
__device__ unsigned long long g_qprof[16];
__shared__ unsigned long long s_qprof[B2AZ_Q_WARPS_DEFAULT_MAX][16];
#define QPROF_ADD(i, v) do { if (lane == 0u) s_qprof[threadIdx.x >> 5][i] += (unsigned long long)(v); } while (0)
#define QPROF_CLK() clock64()
#else
#define QPROF_ADD(i, v) do { } while (0)
#define QPROF_CLK() 0ll
#endif

constexpr int kQCap = 512;    // ring capacity: a power of two > kQGames
constexpr u32 kQEmpty = 0xFFFFu;
enum : u32 { Q_LEVEL = 0, Q_LEAF = 1, Q_MOVE = 2, Q_DONE = 3 };

struct __attribute__((aligned(16))) QGame {  // 17 x 16 B: an odd number of 16 B vectors, so the LDS.128 / STS.128 of
  GameSlot gs;                               // consecutive games fall into different bank groups
  TreeHdr T;        // the tree of the side to move
  Descent D;
  PathSm path;
  u32 sims;         // simulations finished since the state was loaded
  u32 left;         // steps of this launch still to run
  u32 hits;         // cache hits answered inside this launch
  u32 in_descent;   // D describes a descent that reached its leaf (descent_finish is due)
};
static_assert(sizeof(PathSm) == 64, "PathSm layout");
static_assert(sizeof(QGame) == 272, "QGame must stay an odd number of 16 B vectors");

struct QShared {
  QGame game[kQGames];
  u32 head[3];      // items claimed by consumers
  u32 tail[3];      // items reserved by producers
  u32 inflight;     // games popped and not yet pushed back
  u32 done;         // games that finished this launch
  u32 ng;           // games of the current group
  u32 pad_;
  u16 ring[3][kQCap];
};

// CTA-scope fence between a game's state and the publication of its id. __threadfence_block() is fence.sc.cta
// (MEMBAR.SC.CTA); release / acquire order is all the hand-off needs.
__device__ __forceinline__ void q_fence() {
#if defined(B2AZ_Q_FENCE_SC)
  __threadfence_block();
#else
  asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
}

// Hand every lane's game (id >= 0) to queue `ns` (Q_DONE: the game is finished for this launch).
__device__ __forceinline__ void q_push(QShared& S, int id, u32 ns, u32 lane, u32 n_taken) {
  const long long p0 = QPROF_CLK();
  q_fence();  // the game's state (shared and global) is written before its id is published
  QPROF_ADD(14, QPROF_CLK() - p0);
#pragma unroll
  for (u32 t = 0; t < 3u; ++t) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, id >= 0 && ns == t);
    if (m == 0u) continue;
    const int leader = __ffs(m) - 1;
    u32 base = 0;
    if ((int)lane == leader) base = atomicAdd(&S.tail[t], (u32)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (id >= 0 && ns == t) {
      volatile u16* r = &S.ring[t][(base + (u32)__popc(m & ((1u << lane) - 1u))) & (u32)(kQCap - 1)];
#if !defined(B2AZ_Q_NO_PUSH_SPIN)
      for (u32 spin = 0; *r != (u16)kQEmpty && spin < (1u << 26); ++spin) {}  // the slot's previous item has been taken
#endif
      *r = (u16)id;
    }
  }
  QPROF_ADD(15, QPROF_CLK() - p0);
  const unsigned md = __ballot_sync(0xFFFFFFFFu, id >= 0 && ns == Q_DONE);
  if (lane == 0u) {
    if (md) atomicAdd(&S.done, (u32)__popc(md));
    if (n_taken) atomicSub(&S.inflight, n_taken);
  }
}

// Take a batch: returns the queue (or -1 = nothing to do right now, -2 = the group is finished); lanes < n get an id.
__device__ __forceinline__ int q_pop(QShared& S, u32 lane, int& id, u32& n) {
  int q = -1;
  u32 base = 0;
  n = 0;
  const long long p0 = QPROF_CLK();
  if (lane == 0u) {
    volatile u32* head = S.head;
    volatile u32* tail = S.tail;
    for (;;) {
      u32 h[3], a[3];
#pragma unroll
      for (int t = 0; t < 3; ++t) { h[t] = head[t]; a[t] = tail[t] - h[t]; }
      // full batches while the group is busy; as games finish their steps of this launch the floor comes down with
      // the number of games still running, so the tail of a launch keeps many warps going with smaller batches
      const u32 rem = S.ng - *(volatile u32*)&S.done;
      u32 qmin = rem * (u32)B2AZ_Q_MIN / 384u;
      qmin = qmin < 1u ? 1u : (qmin > 32u ? 32u : qmin);
      int pick = -1;
      if (a[Q_MOVE] >= qmin) pick = Q_MOVE;
      else if (a[Q_LEAF] >= qmin && a[Q_LEAF] >= a[Q_LEVEL]) pick = Q_LEAF;
      else if (a[Q_LEVEL] >= qmin) pick = Q_LEVEL;
      else if (a[Q_LEAF] >= qmin) pick = Q_LEAF;
      else if (*(volatile u32*)&S.inflight == 0u) {
        // nobody is working, so nothing more will arrive: take what there is (the tail of a launch)
        u32 best = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t)
          if (a[t] > best) { best = a[t]; pick = t; }
        if (pick < 0) {
          q = (*(volatile u32*)&S.done >= S.ng) ? -2 : -1;
          break;
        }
      } else {
        break;  // wait for a full batch
      }
      const u32 take = a[pick] < 32u ? a[pick] : 32u;
      if (atomicCAS(&S.head[pick], h[pick], h[pick] + take) == h[pick]) {
        atomicAdd(&S.inflight, take);
        q = pick; base = h[pick]; n = take;
        break;
      }
    }
  }
  q = __shfl_sync(0xFFFFFFFFu, q, 0);
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  n = __shfl_sync(0xFFFFFFFFu, n, 0);
  const long long p1 = QPROF_CLK();
  if (q >= 0) QPROF_ADD(12, p1 - p0);
  id = -1;
  if (q >= 0 && lane < n) {
    volatile u16* r = &S.ring[q][(base + lane) & (u32)(kQCap - 1)];
    u32 v = kQEmpty;
    for (u32 spin = 0; (v = *r) == kQEmpty && spin < (1u << 26); ++spin) {}  // reserved by a producer about to write it
    if (v == kQEmpty) v = 0;  // watchdog expired (cannot happen): keep the launch finite
    *r = (u16)kQEmpty;
    id = (int)v;
  }
  q_fence();  // ids are read before the games' state
  if (q >= 0) QPROF_ADD(13, QPROF_CLK() - p1);
  return q;
}

AZ_D void q_load_game(const EngineView& E, u32 g, QGame& q, u32 n_steps) {
  q.gs = E.games[g];
  q.T = E.trees[(size_t)g * kP + q.gs.player];
  q.path.valid = 0;  // the pending leaf's path (if any) is in HBM
  q.sims = 0;
  q.left = n_steps;
  q.hits = 0;
  q.in_descent = 0;
}
AZ_D void q_store_game(const EngineView& E, u32 g, QGame& q) {
  path_flush(E, g, q.path, (u32)q.T.path_len);
  E.trees[(size_t)g * kP + q.gs.player] = q.T;
  E.games[g] = q.gs;
  if (q.sims) E.cold[g].sims += q.sims;
  q.sims = 0;
}

// LEVEL: one PUCT level for a game whose descent goes on.
template <bool GB>
AZ_D u32 q_level(const EngineView& E, u32 g, QGame& q) {
  Descent D = q.D;
  const bool ok = descent_level<GB>(E, g, D, q.path);
  q.D = D;
  return (ok && descent_more(D)) ? Q_LEVEL : Q_LEAF;
}
// start the next descent; a root that has never been visited is its own leaf
template <bool GB>
AZ_D u32 q_begin(const EngineView& E, u32 g, QGame& q, const TreeHdr& T, GameSlot& gs) {
  Descent D;
  descent_begin<GB>(E, g, T, gs, D, q.path, gs.rng);
  q.D = D;
  q.in_descent = 1;
  return descent_more(D) ? Q_LEVEL : Q_LEAF;
}
// LEAF: the end of one simulation's descent and the step boundary that follows it (run_flat's loop body from
// descent_finish to descent_begin).
template <bool GB>
AZ_D u32 q_leaf(const EngineView& E, u32 g, QGame& q) {
  GameSlot gs = q.gs;
  TreeHdr T = q.T;
  u32 left = q.left;
  u32 ns = Q_DONE;
  if (q.in_descent) {
    const Descent D = q.D;
    descent_finish(E, g, T, gs, gs.rng, D);
    q.in_descent = 0;
    bool hit = false;
    if (E.eval_type == 0) {
      // a cache hit is an answered leaf: the game goes on with its next simulation in the same launch
      // (play_manager.cc:589-594); bounded so a launch stays short
      hit = leaf_emit(E, g, gs, D.s, q.hits < E.hit_cap);
      if (hit) ++q.hits;
    }
    if (!hit) --left;
  }
  if (left > 0 && gs.active) {
    bool move = false;
    if (gs.initialized) {
      const u32 cp = gs.player;
      const bool noise = (E.epsilon > 0.0f) && !gs.capped;
      process_result(E, g, T, gs, gs.rng, noise, q.path);
      ++q.sims;
      const u32 goal = seat_budget(E, g, cp, gs.capped != 0);
      move = T.depth >= goal;
    } else {
      gs.initialized = 1;
      gs.capped = (E.playout_cap && rng_uniform01(gs.rng) < E.playout_cap_percent) ? 1 : 0;
      if (GB && E.gumbel_enabled) gumbel_arm(E, g, gs.player, gs.capped != 0);
    }
    ns = move ? (u32)Q_MOVE : q_begin<GB>(E, g, q, T, gs);
  }
  q.gs = gs;
  q.T = T;
  q.left = left;
  return ns;
}
// The two halves of q_leaf as separate pieces (the waves kernel runs them in different rounds): FINISH = the end of a
// descent (expansion, leaf batch), BOUNDARY = process_result .. descent_begin. Q_PR = "needs BOUNDARY next".
constexpr u32 Q_PR = Q_LEAF;
template <bool GB>
AZ_D u32 q_finish(const EngineView& E, u32 g, QGame& q) {
  GameSlot gs = q.gs;
  TreeHdr T = q.T;
  const Descent D = q.D;
  descent_finish(E, g, T, gs, gs.rng, D);
  q.in_descent = 0;
  bool hit = false;
  if (E.eval_type == 0) {
    hit = leaf_emit(E, g, gs, D.s, q.hits < E.hit_cap);
    if (hit) ++q.hits;
  }
  u32 left = q.left;
  if (!hit) --left;
  q.left = left;
  q.gs = gs;
  q.T = T;
  return (left > 0 && gs.active) ? Q_PR : Q_DONE;
}
template <bool GB>
AZ_D u32 q_boundary(const EngineView& E, u32 g, QGame& q) {
  GameSlot gs = q.gs;
  TreeHdr T = q.T;
  bool move = false;
  if (gs.initialized) {
    const u32 cp = gs.player;
    const bool noise = (E.epsilon > 0.0f) && !gs.capped;
    process_result(E, g, T, gs, gs.rng, noise, q.path);
    ++q.sims;
    const u32 goal = seat_budget(E, g, cp, gs.capped != 0);
    move = T.depth >= goal;
  } else {
    gs.initialized = 1;
    gs.capped = (E.playout_cap && rng_uniform01(gs.rng) < E.playout_cap_percent) ? 1 : 0;
    if (GB && E.gumbel_enabled) gumbel_arm(E, g, gs.player, gs.capped != 0);
  }
  u32 ns = Q_MOVE;
  if (!move) ns = q_begin<GB>(E, g, q, T, gs);
  q.gs = gs;
  q.T = T;
  if (ns == Q_LEAF) ns = q_finish<GB>(E, g, q);  // a root that was never visited is its own leaf
  return ns;
}
// MOVE: the search budget is reached. play_move() works on the slot's state in HBM.
template <bool GB>
AZ_D u32 q_move(const EngineView& E, u32 g, QGame& q) {
  q_store_game(E, g, q);
  const bool retired = play_move(E, g);
  q.gs = E.games[g];
  q.T = E.trees[(size_t)g * kP + q.gs.player];
  q.path.valid = 0;
  if (retired) return Q_DONE;
  GameSlot gs = q.gs;
  const u32 ns = q_begin<GB>(E, g, q, q.T, gs);
  q.gs = gs;
  return ns;
}

template <bool GB>
__global__ void __launch_bounds__(B2AZ_Q_WARPS * 32, 1) k_step_q(EngineView E, u32 n_steps, u32 games_per_group, u32 n_groups) {
  extern __shared__ __align__(16) unsigned char q_smem[];
  QShared& S = *reinterpret_cast<QShared*>(q_smem);
  const u32 tid = threadIdx.x, lane = tid & 31u;
#if defined(B2AZ_Q_PROF)
  if (lane < 16u) s_qprof[tid >> 5][lane] = 0;
  __syncwarp();
#endif
  for (u32 grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const u32 g0 = grp * games_per_group;
    const u32 ng = (g0 >= E.G) ? 0u : (E.G - g0 < games_per_group ? E.G - g0 : games_per_group);
    if (tid < 3u) { S.head[tid] = 0; S.tail[tid] = 0; }
    if (tid == 0u) {
      S.inflight = 0; S.done = 0; S.ng = ng;
    }
    for (u32 i = tid; i < 3u * (u32)kQCap; i += blockDim.x) (&S.ring[0][0])[i] = (u16)kQEmpty;
    __syncthreads();
    // load the group's state; every active game starts at a step boundary
    for (u32 i0 = 0; i0 < ng; i0 += blockDim.x) {
      const u32 i = i0 + tid;
      int id = -1;
      u32 ns = Q_DONE;
      if (i < ng) {
        QGame& q = S.game[i];
        q_load_game(E, g0 + i, q, n_steps);
        id = (int)i;
        ns = q.gs.active ? (u32)Q_LEAF : (u32)Q_DONE;
      }
      q_push(S, id, ns, lane, 0u);
    }
    __syncthreads();
    u32 idle = 0, nap = 64;
    for (;;) {
      int id;
      u32 n;
      const long long t0 = QPROF_CLK();
      const int qsel = q_pop(S, lane, id, n);
      if (qsel == -2) break;
      if (qsel < 0) {
        // watchdog: a scheduling bug must end the launch with an error, not hang the GPU (seconds of idling)
        if (++idle > (1u << 22)) {
          at_or(&E.glob->error, B2AZ_DEVERR_QUEUE);
          break;
        }
        __nanosleep(nap);  // back off: idle warps must not fight the working ones for the shared-memory pipe
        if (nap < (u32)B2AZ_Q_NAP_MAX) nap *= 2u;
        QPROF_ADD(3, QPROF_CLK() - t0);
        QPROF_ADD(11, 1);
        continue;
      }
      idle = 0;
      nap = 64;
      const long long t1 = QPROF_CLK();
      u32 ns = Q_DONE;
      if (id >= 0) {
        QGame& q = S.game[id];
        const u32 g = g0 + (u32)id;
        if (qsel == (int)Q_LEVEL) ns = q_level<GB>(E, g, q);
        else if (qsel == (int)Q_LEAF) ns = q_leaf<GB>(E, g, q);
        else ns = q_move<GB>(E, g, q);
      }
      __syncwarp();
      const long long t2 = QPROF_CLK();
      q_push(S, id, ns, lane, n);
      QPROF_ADD(qsel, t2 - t1);
      QPROF_ADD(4, (t1 - t0) + (QPROF_CLK() - t2));
      QPROF_ADD(5 + qsel, 1);
      QPROF_ADD(8 + qsel, n);
    }
    __syncthreads();
    for (u32 i = tid; i < ng; i += blockDim.x) q_store_game(E, g0 + i, S.game[i]);
    __syncthreads();
  }
#if defined(B2AZ_Q_PROF)
  __syncwarp();
  if (lane < 16u) atomicAdd(&g_qprof[lane], s_qprof[tid >> 5][lane]);
#endif
}

}  // namespace b2az

#endif  // !B2AZ_HOST_EMU
