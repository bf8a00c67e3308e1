// az_engine.cu — kernels + host side of the C ABI declared in include/b2az.h.
//
// Kernels (sm_100a):
//   k_step           one THREAD per game slot; per step runs game_step() =
//                    process_result -> [move] -> find_leaf (az_engine_logic.h). With RANDOM eval the
//                    evaluator is inline and n_steps are fused into one launch: the game slot, the
//                    mover's tree header and the RNG stay in registers across the steps.
//   k_step_serial    B2AZ_RNG_GLOBAL: one thread walks the slots in ascending order so the single
//                    pcg32 stream is consumed in the reference's order (bit-exact parity mode).
//   k_canonicalize   compact leaf positions -> dense float32[B][4][6][7] for the torch net.
//   k_hist_expand    compact finished samples -> canonical / v / pi arrays.
//   k_peek, k_stats, k_c4_batch: inspection + the batched bitboard game kernels.
//
// The file also compiles as plain C++ with -DB2AZ_HOST_EMU (kernels become loops over W = 1
// groups, device memory becomes calloc). That build exists ONLY for the CPU test-suite
// (tests/cpp/, build/libb2az_hostemu.so); the product library never defines the macro and every
// entry point fails with B2AZ_ECUDA when no CUDA device is usable.
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b2az.h"
#include "az_engine_logic.h"
#include "az_engine_queue.h"
#include "az_engine_waves.h"

#ifndef B2AZ_HOST_EMU
#include <cuda_runtime.h>
#endif

using namespace b2az;

namespace {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#ifndef B2AZ_HOST_EMU
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t err__ = (expr);                                                              \
    if (err__ != cudaSuccess)                                                                \
      return fail(B2AZ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));       \
  } while (0)
typedef cudaStream_t stream_t;
#else
#define CUDA_TRY(expr) \
  do {                 \
    (void)(expr);      \
  } while (0)
typedef void* stream_t;
#endif

// ------------------------------------------------------------------------------------ memory shims
template <typename T>
int dev_alloc(T** p, size_t count) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  CUDA_TRY(cudaMemset(*p, 0, count * sizeof(T)));
#else
  *p = static_cast<T*>(calloc(count ? count : 1, sizeof(T)));
  if (!*p) return fail(B2AZ_ENOMEM, "calloc failed");
#endif
  return 0;
}
template <typename T>
int dev_alloc_raw(T** p, size_t count) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
#else
  *p = static_cast<T*>(malloc((count ? count : 1) * sizeof(T)));
  if (!*p) return fail(B2AZ_ENOMEM, "malloc failed");
#endif
  return 0;
}
template <typename T>
void dev_free(T* p) {
  if (!p) return;
#ifndef B2AZ_HOST_EMU
  cudaFree(const_cast<typename std::remove_const<T>::type*>(p));
#else
  free(const_cast<typename std::remove_const<T>::type*>(p));
#endif
}
int copy_h2d(void* dst, const void* src, size_t bytes, stream_t s) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
#else
  (void)s;
  memcpy(dst, src, bytes);
#endif
  return 0;
}
int copy_d2h(void* dst, const void* src, size_t bytes, stream_t s) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
#else
  (void)s;
  memcpy(dst, src, bytes);
#endif
  return 0;
}
int copy_d2d(void* dst, const void* src, size_t bytes, stream_t s) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
#else
  (void)s;
  memcpy(dst, src, bytes);
#endif
  return 0;
}
int dev_zero(void* dst, size_t bytes, stream_t s) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaMemsetAsync(dst, 0, bytes, s));
#else
  (void)s;
  memset(dst, 0, bytes);
#endif
  return 0;
}
int stream_sync(stream_t s) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaStreamSynchronize(s));
#else
  (void)s;
#endif
  return 0;
}

// ------------------------------------------------------------------------------------ kernels
#ifndef B2AZ_HOST_EMU
#define KERNEL __global__
#define GLOBAL_TID (blockIdx.x * blockDim.x + threadIdx.x)
#define GLOBAL_NT (gridDim.x * blockDim.x)
#else
#define KERNEL
#endif

struct InitArgs {
  u64 seed;
  u32 rng_mode;
};

#ifndef B2AZ_HOST_EMU
__global__ void k_init_pool(EngineView E) {
  // every page starts as a one-page chain in its own ring slot
  for (u32 p = GLOBAL_TID; p < E.num_pages; p += GLOBAL_NT) {
    E.page_next[p] = kNil;
    E.ring[p] = p;
  }
}
__global__ void k_init_games(EngineView E, InitArgs a) {
  for (u32 g = GLOBAL_TID; g < E.G; g += GLOBAL_NT) {
    GameSlot gs;
    memset(&gs, 0, sizeof(gs));
    gs.active = 1;
    pcg32_seed(gs.rng, a.seed + (u64)g);
    E.games[g] = gs;
    GameCold gc;
    memset(&gc, 0, sizeof(gc));
    E.cold[g] = gc;
    TreeHdr T;
    memset(&T, 0, sizeof(T));
    tree_reset(T);
    T.region = g / E.region_games;
    E.trees[(size_t)g * kP + 0] = T;
    E.trees[(size_t)g * kP + 1] = T;
  }
  for (u32 r = GLOBAL_TID; r < E.n_regions; r += GLOBAL_NT) {
    E.ring_tickets[2u * r] = 0;                    // pop tickets
    E.ring_tickets[2u * r + 1u] = E.region_pages;  // push tickets: the ring starts full
  }
  if (GLOBAL_TID == 0) {
    Globals* G = E.glob;
    memset(G, 0, sizeof(Globals));
    G->games_started = E.G;
    G->active_games = E.G;
    pcg32_seed(G->global_rng, a.seed);
  }
}

// One thread per game slot. <= 144 registers so that seven 64-thread CTAs (14 warps) fit an SM: at the
// BASELINE size (65,536 games over 148 SMs = 443 threads per SM) every game is resident at once.
template <bool GB>
__global__ void __launch_bounds__(64, 7) k_step(EngineView E, u32 n_steps) {
  const u32 g = GLOBAL_TID;
  if (g >= E.G) return;
  Ctx c;
  ctx_load(E, g, c);
  if (!c.gs.active) return;
  run_flat<GB>(E, g, c, n_steps);
  ctx_store(E, g, c);
}
// The same, with the 32 games of a warp in lock step (run_sync): no lane leaves early, the warp votes. The selection
// path lives in shared memory (one column per thread): an indexed access is one LDS / STS instead of a chain of selects.
template <bool GB, bool PX = false>
__global__ void __launch_bounds__(64, 7) k_step_sync(EngineView E, u32 n_steps) {
  const u32 g = GLOBAL_TID;
  const bool in_range = g < E.G;
  const u32 gg = in_range ? g : 0u;
  Ctx c;
  ctx_load(E, gg, c);
#if defined(B2AZ_SYNC_PATH_SMEM)
  __shared__ u32 s_blk[kPathSm][kPathColStride];
  __shared__ u8 s_slot[kPathSm][kPathColStride];
  PathCol pr;
  pr.blk = &s_blk[0][threadIdx.x];
  pr.slot = &s_slot[0][threadIdx.x];
  pr.valid = 0;  // the pending leaf's path (if any) is in HBM
  run_sync<GB, PathCol, PX>(E, gg, c, pr, n_steps, in_range);
  if (in_range && c.gs.active) {
    path_flush(E, gg, pr, (u32)c.T.path_len);
    ctx_store(E, gg, c);
  }
#else
  run_sync<GB, PathRegs, PX>(E, gg, c, c.pr, n_steps, in_range);
  if (in_range && c.gs.active) ctx_store(E, gg, c);
#endif
}
// update_inferences' cache half (play_manager.cc:619-642: insert_many of every evaluated leaf), as its own
// launch BEFORE the step kernel: during k_step the table is then read-only (plus frequency bumps), so lookups
// need neither fences nor locks.
__global__ void k_cache_insert(EngineView E) {
  const u32 count = E.glob->leaf_count;  // still the previous generation's: the step zeroes it after this launch
  for (u32 r = GLOBAL_TID; r < count; r += GLOBAL_NT)
    cache_insert(E, E.leaf_key[r], E.ev_v + (size_t)r * (kP + 1), E.ev_pi + (size_t)r * kA);
}
// b2az_cache_insert_host / b2az_cache_find_host: the position cache driven key by key IN ORDER (one thread), for tools and
// for the tests that hold it against the reference's S3FIFOCache
__global__ void k_cache_insert_keys(EngineView E, const u64* keys, const float* v, const float* pi, u32 n) {
  for (u32 i = 0; i < n; ++i) cache_insert(E, keys[i], v + (size_t)i * (kP + 1), pi + (size_t)i * kA);
}
__global__ void k_cache_find_keys(EngineView E, const u64* keys, u32 n, u8* found, float* v, float* pi) {
  for (u32 i = 0; i < n; ++i) {
    const u32 idx = cache_find(E, keys[i]);
    found[i] = idx != kNil;
    if (idx != kNil) {
      for (int j = 0; j < kP + 1; ++j) v[(size_t)i * (kP + 1) + j] = E.cache_vals[idx].v[j];
      for (int j = 0; j < kA; ++j) pi[(size_t)i * kA + j] = E.cache_vals[idx].pi[j];
    }
  }
}
__global__ void k_step_serial(EngineView E, u32 n_steps) {
  for (u32 s = 0; s < n_steps; ++s)
    for (u32 g = 0; g < E.G; ++g) {
      Ctx c;
      ctx_load(E, g, c);
      if (!c.gs.active) continue;
      game_step(E, g, c);
      ctx_store(E, g, c);
    }
}

__global__ void k_canonicalize(EngineView E, const u32* __restrict__ count_ptr, float* __restrict__ out) {
  const u32 count = *count_ptr;
  const size_t total = (size_t)count * C4_CANON;
  for (size_t i = GLOBAL_TID; i < total; i += GLOBAL_NT) {
    const u32 row = (u32)(i / C4_CANON), e = (u32)(i % C4_CANON);
    out[i] = c4_canon_elem(E.leaf_p0[row], E.leaf_p1[row], E.leaf_player[row], e);
  }
}
// GameState::valid_moves() of every leaf of the batch: uint8[B][7] next to the canonical batch (the mask half of the
// zero-copy feed; the reference's net masks its policy head with it)
__global__ void k_leaf_valid(EngineView E, const u32* __restrict__ count_ptr, u8* __restrict__ out) {
  const u32 count = *count_ptr;
  for (u32 row = GLOBAL_TID; row < count; row += GLOBAL_NT) {
    C4State s;
    s.p[0] = E.leaf_p0[row]; s.p[1] = E.leaf_p1[row]; s.turn = 0; s.player = 0;
    const u32 m = c4_valid_mask(s);
#pragma unroll
    for (int w = 0; w < 7; ++w) out[(size_t)row * 7 + w] = (u8)((m >> w) & 1u);
  }
}
// nsym = 1: one row per sample. nsym = 2: every sample followed by its mirror image (Connect4GS::symmetries,
// connect4_gs.cc:151-170: canonical(f, h, w) <- canonical(f, h, 6 - w), pi(w) <- pi(6 - w), v unchanged), the
// order game_runner.exploit_symmetries writes them in (game_runner.py:1083-1115).
__global__ void k_hist_expand(EngineView E, unsigned long long first, u32 count, u32 nsym, float* __restrict__ canon,
                              float* __restrict__ v, float* __restrict__ pi) {
  const size_t total = (size_t)count * nsym * C4_CANON;
  for (size_t i = GLOBAL_TID; i < total; i += GLOBAL_NT) {
    const u32 row = (u32)(i / C4_CANON), e = (u32)(i % C4_CANON);
    const u32 sample = row / nsym, mirror = row % nsym;
    const HistEntry& h = E.hist_out[(first + sample) % (unsigned long long)E.hist_capacity];
    canon[i] = c4_canon_elem(h.p0, h.p1, h.player, mirror ? c4_mirror_elem(e) : e);
    if (e < 3u) v[(size_t)row * 3 + e] = (h.result == e + 1u) ? 1.0f : 0.0f;
    if (e < (u32)kA) pi[(size_t)row * kA + e] = h.pi[mirror ? (u32)kA - 1u - e : e];
  }
}
#endif  // !B2AZ_HOST_EMU

struct PeekOut {
  u8 state[89];
  u32 counts[kA];
  float q[kA];
  float policy[kA];
  float root_value[3];
  u32 depth, root_n;
};
AZ_HD void peek_impl(const EngineView& E, u32 g, u32 seat, PeekOut* o) {
  const GameSlot& gs = E.games[g];
  C4State s;
  s.p[0] = gs.p0; s.p[1] = gs.p1; s.turn = gs.turn; s.player = gs.player;
  c4_to_board(s, reinterpret_cast<signed char*>(o->state));
  o->state[84] = gs.player;
  const u32 turn = gs.turn;
  o->state[85] = (u8)(turn & 0xFF); o->state[86] = (u8)((turn >> 8) & 0xFF);
  o->state[87] = (u8)((turn >> 16) & 0xFF); o->state[88] = (u8)((turn >> 24) & 0xFF);
  const TreeHdr& T = E.trees[(size_t)g * kP + seat];
  RootView R;
  root_view(E, T, R);
  for (int m = 0; m < kA; ++m) { o->counts[m] = 0; o->q[m] = 0.0f; o->policy[m] = 0.0f; }
  for (u32 j = 0; j < R.k; ++j) {
    o->counts[R.mv[j]] = R.n[j];
    o->q[R.mv[j]] = R.n[j] ? R.q[j] : 0.0f;
    o->policy[R.mv[j]] = R.pol[j];
  }
  mcts_root_value(T, R, o->root_value);
  o->depth = T.depth;
  o->root_n = T.n;
}
struct StatsOut {
  unsigned long long sims, moves;
};
#ifndef B2AZ_HOST_EMU
__global__ void k_peek(EngineView E, u32 g, u32 seat, PeekOut* o) { peek_impl(E, g, seat, o); }
__global__ void k_stats(EngineView E, StatsOut* o) {
  unsigned long long s = 0, m = 0;
  for (u32 g = GLOBAL_TID; g < E.G; g += GLOBAL_NT) { s += E.cold[g].sims; m += E.cold[g].nmoves; }
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_down_sync(0xFFFFFFFFu, s, off);
    m += __shfl_down_sync(0xFFFFFFFFu, m, off);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&o->sims, s); atomicAdd(&o->moves, m); }
}
__global__ void k_count_free_pages(EngineView E, unsigned long long* out) {
  // walks every chain in the ring; only meaningful while no step kernel is running
  unsigned long long c = 0;
  for (u32 i = GLOBAL_TID; i < E.num_pages; i += GLOBAL_NT) c += (E.ring[i] != kNil);
  for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
#endif

struct C4BatchArgs {
  u32 n;
  const signed char* boards;
  const u8* players;
  const u32* turns;
  const u32* moves;
  signed char* boards_out;
  u8* players_out;
  u8* valid;
  float* scores;
  u8* terminal;
  float* canonical;
  i32* status;
};
AZ_HD void c4_batch_one(const C4BatchArgs& a, u32 i) {
  C4State s;
  c4_from_board(s, a.boards + (size_t)i * 84, a.players[i], (int)a.turns[i]);
  i32 st = 0;
  if (a.moves && a.moves[i] != 0xFFFFFFFFu) {
    if (a.moves[i] >= (u32)kA || !c4_play(s, a.moves[i])) st = B2AZ_EMOVE;
  }
  if (a.status) a.status[i] = st;
  if (a.boards_out) c4_to_board(s, a.boards_out + (size_t)i * 84);
  if (a.players_out) a.players_out[i] = s.player;
  if (a.valid) {
    const u32 vm = c4_valid_mask(s);
    for (int w = 0; w < kA; ++w) a.valid[(size_t)i * kA + w] = (u8)((vm >> w) & 1u);
  }
  const u32 term = c4_terminal(s);
  if (a.terminal) a.terminal[i] = term ? 1 : 0;
  if (a.scores)
    for (u32 e = 0; e < 3u; ++e) a.scores[(size_t)i * 3 + e] = (term == e + 1u) ? 1.0f : 0.0f;
  if (a.canonical)
    for (u32 e = 0; e < (u32)C4_CANON; ++e)
      a.canonical[(size_t)i * C4_CANON + e] = c4_canon_elem(s.p[0], s.p[1], s.player, e);
}
#ifndef B2AZ_HOST_EMU
__global__ void k_c4_batch(C4BatchArgs a) {
  for (u32 i = GLOBAL_TID; i < a.n; i += GLOBAL_NT) c4_batch_one(a, i);
}
#endif

}  // namespace

// ------------------------------------------------------------------------------------ engine object
struct b2az_engine {
  b2az_params params;
  int device = 0;
  int num_sms = 148;
  uint32_t n_perms = 1;         // seat permutations (b2az_params.n_seat_perms; the tables live at view.perms)
  EngineView view;
  // owned device buffers
  float* canon_buf = nullptr;   // [G][168]
  u8* valid_buf = nullptr;      // [G][7] legal-move mask of the leaf batch (b2az_leaf_valid_device), allocated on first use
  float* ev_v_buf = nullptr;    // [G][3]   (legacy host path staging target)
  float* ev_pi_buf = nullptr;   // [G][7]
  PeekOut* peek_buf = nullptr;
  StatsOut* stats_buf = nullptr;
  unsigned long long* freepages_buf = nullptr;
  float* hist_canon = nullptr;  // drain staging when the destination is host memory
  float* hist_v = nullptr;
  float* hist_pi = nullptr;
  u32 hist_stage_cap = 0;
  // host-side bookkeeping of the current leaf batch
  bool leaves_pending = false;      // a step produced leaves that have not all been answered
  u32 leaf_count = 0;               // valid once leaf_count_known
  u32 device_error_seen = 0;        // Globals::error as read together with leaf_count
  bool leaf_count_known = false;
  bool canon_ready = false;
  u32 leaves_taken = 0;             // legacy build_batch cursor
  u32 evals_submitted = 0;
  bool evals_all = false;           // b2az_submit_eval_all: the whole batch was answered on the device
  std::vector<u32> leaf_ids_host;   // slot id of every leaf row (for submit_eval_host)
  std::vector<u32> row_of_game;     // slot id -> row
  std::vector<float> stage_v, stage_pi;
  bool started = false;
  u32 step_kernel = 0;              // B2AZ_STEP_*
  u32 group_games = 0, groups = 0;  // slots per persistent CTA (== pool region), number of such groups
  // overlapped history drain (b2az_history_mark / b2az_drain_history_marked)
  unsigned long long* hist_mark_host = nullptr;  // pinned: {hist_written, hist_read} as of the mark
  unsigned long long* hist_read_host = nullptr;  // pinned: the new read counter on its way to the device
  void* hist_event = nullptr;                    // cudaEvent_t recorded by the mark
  bool hist_marked = false;
};

namespace {

// The CUDA current device is per host thread (a fresh Python thread starts on device 0): every entry point that
// launches or copies binds the engine's device first.
int bind_device(b2az_engine* e) {
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaSetDevice(e->device));
#else
  (void)e;
#endif
  return 0;
}

int sync_leaf_count(b2az_engine* e, stream_t s) {
  if (e->leaf_count_known) return 0;
  u32 c[2] = {0, 0};  // Globals::error and ::leaf_count are neighbours: one copy, one synchronisation per generation
  static_assert(offsetof(Globals, leaf_count) == offsetof(Globals, error) + sizeof(u32), "Globals layout");
  if (int rc = copy_d2h(c, &e->view.glob->error, sizeof(c), s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  e->device_error_seen = c[0];
  e->leaf_count = c[1];
  e->leaf_count_known = true;
  return 0;
}

int ensure_canon(b2az_engine* e, stream_t s) {
  if (e->canon_ready) return 0;
#ifndef B2AZ_HOST_EMU
  const int blocks = e->num_sms * 8;
  k_canonicalize<<<blocks, 256, 0, s>>>(e->view, &e->view.glob->leaf_count, e->canon_buf);
  CUDA_TRY(cudaGetLastError());
#else
  const u32 count = e->view.glob->leaf_count;
  for (size_t i = 0; i < (size_t)count * C4_CANON; ++i) {
    const u32 row = (u32)(i / C4_CANON), el = (u32)(i % C4_CANON);
    e->canon_buf[i] = c4_canon_elem(e->view.leaf_p0[row], e->view.leaf_p1[row], e->view.leaf_player[row], el);
  }
#endif
  e->canon_ready = true;
  return 0;
}

int check_device_error(b2az_engine* e, stream_t s) {
  u32 err = 0;
  if (e->leaf_count_known) {
    err = e->device_error_seen;  // read together with the leaf count of this generation
  } else {
    if (int rc = copy_d2h(&err, &e->view.glob->error, sizeof(u32), s)) return rc;
    if (int rc = stream_sync(s)) return rc;
  }
  if (err & B2AZ_DEVERR_POOL) return fail(B2AZ_ENOMEM, "device tree-node pool exhausted: raise b2az_params.pool_nodes");
  if (err & B2AZ_DEVERR_MOVE) return fail(B2AZ_EMOVE, "device: update_root could not find the move / illegal move");
  if (err & B2AZ_DEVERR_DEPTH) return fail(B2AZ_ESTATE, "device: selection path exceeded the path buffer");
  if (err & B2AZ_DEVERR_QUEUE) return fail(B2AZ_ESTATE, "device: the step kernel's work queues stalled (watchdog)");
  if (err & B2AZ_DEVERR_HIST) return fail(B2AZ_ENOMEM, "device history ring overflowed: drain more often or raise history_capacity");
  return 0;
}

}  // namespace

extern "C" {

const char* b2az_last_error(void) { return g_last_error.c_str(); }

int b2az_params_default(b2az_params* p) {
  if (!p) return fail(B2AZ_EINVAL, "null params");
  memset(p, 0, sizeof(*p));
  p->game = B2AZ_GAME_CONNECT4;
  p->games_to_play = 1;
  p->concurrent_games = 1;
  p->max_batch_size = 0;
  p->mcts_visits[0] = p->mcts_visits[1] = 100;
  p->cpuct = 2.0f;          // PlayParams defaults (play_manager.h:83-102)
  p->start_temp = 1.0f;
  p->final_temp = 1.0f;
  p->temp_decay_half_life = 0.0f;
  p->tree_reuse = 1;
  p->epsilon = 0.0f;
  p->mcts_root_temp = 1.0f;
  p->playout_cap_depth = 25;
  p->playout_cap_percent = 0.75f;
  p->gumbel_m = 16;
  p->gumbel_c_visit = 50.0f;
  p->gumbel_c_scale = 1.0f;
  p->eval_type = B2AZ_EVAL_NN;
  p->rng_mode = B2AZ_RNG_PER_GAME;
  p->seed = 0;
  return 0;
}

#if defined(B2AZ_W_PROF) && !defined(B2AZ_HOST_EMU)
int b2az_debug_wprof(unsigned long long* out16) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, g_wprof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_wprof, z, sizeof(z));
  return 0;
}
#endif
#if defined(B2AZ_Q_PROF) && !defined(B2AZ_HOST_EMU)
// experiment builds only: read (and clear) the queue kernel's time accounting
int b2az_debug_qprof(unsigned long long* out16) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, g_qprof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_qprof, z, sizeof(z));
  return 0;
}
#endif

int b2az_destroy(b2az_engine* e) {
  if (!e) return 0;
#ifndef B2AZ_HOST_EMU
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
#endif
  EngineView& V = e->view;
  dev_free(V.blocks); dev_free(V.page_next); dev_free(V.ring); dev_free(V.ring_tickets);
  dev_free(V.trees); dev_free(V.games); dev_free(V.cold); dev_free(V.path); dev_free(V.pslot); dev_free(V.gum);
  dev_free(V.leaf_p0); dev_free(V.leaf_p1); dev_free(V.leaf_player); dev_free(V.leaf_game); dev_free(V.leaf_seat);
  dev_free(const_cast<PermTables*>(V.perms));
  dev_free(V.hist_partial); dev_free(V.hist_out); dev_free(V.glob);
  dev_free(V.cache_keys); dev_free(V.cache_meta); dev_free(V.cache_lock); dev_free(V.cache_vals);
  dev_free(V.cache_ghost); dev_free(V.leaf_key); dev_free(V.hit_val);
  dev_free(e->canon_buf); dev_free(e->valid_buf); dev_free(e->ev_v_buf); dev_free(e->ev_pi_buf);
  dev_free(e->peek_buf); dev_free(e->stats_buf); dev_free(e->freepages_buf);
  dev_free(e->hist_canon); dev_free(e->hist_v); dev_free(e->hist_pi);
#ifndef B2AZ_HOST_EMU
  if (e->hist_mark_host) cudaFreeHost(e->hist_mark_host);
  if (e->hist_read_host) cudaFreeHost(e->hist_read_host);
  if (e->hist_event) cudaEventDestroy((cudaEvent_t)e->hist_event);
#else
  free(e->hist_mark_host); free(e->hist_read_host);
#endif
  delete e;
  return 0;
}

int b2az_create(const b2az_params* p, int device, b2az_engine** out) {
  if (!p || !out) return fail(B2AZ_EINVAL, "null argument");
  *out = nullptr;
  if (p->game != B2AZ_GAME_CONNECT4) return fail(B2AZ_EINVAL, "only B2AZ_GAME_CONNECT4 is implemented");
  if (p->concurrent_games == 0) return fail(B2AZ_EINVAL, "concurrent_games must be > 0");
  if (p->games_to_play < p->concurrent_games)
    return fail(B2AZ_EINVAL, "games_to_play must be >= concurrent_games (every slot starts a game)");
  if (p->mcts_visits[0] == 0 || p->mcts_visits[1] == 0)
    return fail(B2AZ_EINVAL, "You must specify MCTS visits for each player");  // play_manager.cc:21
  if (p->gumbel_enabled && p->gumbel_m == 0) return fail(B2AZ_EINVAL, "gumbel_m must be > 0");
  if ((p->resign_percent != 0.0f || p->playout_cap_randomization) && p->rng_mode == B2AZ_RNG_GLOBAL)
    return fail(B2AZ_EINVAL, "resign_percent / playout_cap_randomization flip coins from an unseedable engine in the "
                             "reference: not available in B2AZ_RNG_GLOBAL (parity) mode");
  if (p->playout_cap_randomization && p->playout_cap_depth == 0) return fail(B2AZ_EINVAL, "playout_cap_depth must be > 0");
  if (p->eval_type != B2AZ_EVAL_NN && p->eval_type != B2AZ_EVAL_RANDOM) return fail(B2AZ_EINVAL, "bad eval_type");
  if (p->rng_mode != B2AZ_RNG_PER_GAME && p->rng_mode != B2AZ_RNG_GLOBAL) return fail(B2AZ_EINVAL, "bad rng_mode");
  if (p->max_cache_size != 0 && p->rng_mode == B2AZ_RNG_GLOBAL)
    return fail(B2AZ_EINVAL, "the position cache changes the evaluation order: not available in B2AZ_RNG_GLOBAL (parity) mode");
  if (p->step_kernel > B2AZ_STEP_QUEUE) return fail(B2AZ_EINVAL, "bad step_kernel");
  if (p->model_groups[0] > 1 || p->model_groups[1] > 1) return fail(B2AZ_EINVAL, "model_groups: a group index must be 0 or 1");
  if (p->n_seat_perms > (uint32_t)kMaxPerms) return fail(B2AZ_EINVAL, "seat_perms: at most 8 seat permutations");
  if (p->n_seat_perms > 1 && p->concurrent_games % p->n_seat_perms != 0)
    return fail(B2AZ_EINVAL, "seat_perms: concurrent_games must be a multiple of the number of seat permutations (slot g plays permutation g % n)");
  for (uint32_t pm = 0; pm < p->n_seat_perms && pm < (uint32_t)kMaxPerms; ++pm)
    if (p->seat_perms[pm][0] > 1 || p->seat_perms[pm][1] > 1) return fail(B2AZ_EINVAL, "seat_perms: a group index must be 0 or 1");
  if (p->n_seat_perms > 1 && p->rng_mode == B2AZ_RNG_GLOBAL)
    return fail(B2AZ_EINVAL, "seat_perms: not available in B2AZ_RNG_GLOBAL (parity) mode (the hand-out order is per slot here)");
  if (p->per_slot_quota && p->games_to_play % p->concurrent_games != 0)
    return fail(B2AZ_EINVAL, "per_slot_quota: games_to_play must be a multiple of concurrent_games");
  if (p->lanes_per_game > 1)
    return fail(B2AZ_EINVAL, "lanes_per_game must be 0 or 1: Connect4 runs one thread per game slot (DESIGN.md 3)");
#ifdef B2AZ_HOST_EMU
  (void)device;
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(B2AZ_EINVAL, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
#endif
  b2az_engine* e = new b2az_engine();
  e->params = *p;
  e->device = device;
  e->step_kernel = p->step_kernel == B2AZ_STEP_DEFAULT ? B2AZ_STEP_SYNC : p->step_kernel;
  if (const char* sk = getenv("B2AZ_STEP_KERNEL"))
    e->step_kernel = sk[0] == 'f' ? B2AZ_STEP_FLAT : sk[0] == 'w' ? B2AZ_STEP_WAVES : sk[0] == 's' ? B2AZ_STEP_SYNC : B2AZ_STEP_QUEUE;
#ifndef B2AZ_HOST_EMU
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  e->num_sms = prop.multiProcessorCount;
#endif
  const u32 G = p->concurrent_games;
  EngineView& V = e->view;
  memset(&V, 0, sizeof(V));
  V.G = G;
  V.games_to_play = p->games_to_play;
  // seat_visits_ / seat_cap_visits_ / seat_perms_ of every permutation (play_manager.cc:46-90); permutation 0 also fills
  // the view's own fields, which is all a one-seating run reads
  PermTables pt;
  memset(&pt, 0, sizeof(pt));
  pt.n_perms = std::max(1u, p->n_seat_perms);
  for (u32 pm = 0; pm < pt.n_perms; ++pm)
    for (int seat = 0; seat < 2; ++seat) {
      pt.visits[pm][seat] = p->perm_seat_visits[pm][seat] ? p->perm_seat_visits[pm][seat] : p->mcts_visits[seat];
      const u32 cv = p->perm_seat_cap_visits[pm][seat] ? p->perm_seat_cap_visits[pm][seat] : p->seat_cap_visits[seat];
      pt.cap_visits[pm][seat] = cv ? cv : p->playout_cap_depth;
      pt.seat_group[pm][seat] = p->n_seat_perms ? p->seat_perms[pm][seat] : p->model_groups[seat];
    }
  pt.random_groups = p->eval_type == B2AZ_EVAL_NN ? ((p->group_random[0] ? 1u : 0u) | (p->group_random[1] ? 2u : 0u)) : 0u;
  for (int seat = 0; seat < 2; ++seat) {
    V.visits[seat] = pt.visits[0][seat];
    V.cap_visits[seat] = pt.cap_visits[0][seat];
    V.seat_group[seat] = pt.seat_group[0][seat];
  }
  e->n_perms = pt.n_perms;
  V.cpuct = p->cpuct; V.fpu_reduction = p->fpu_reduction; V.epsilon = p->epsilon; V.root_temp = p->mcts_root_temp;
  V.start_temp = p->start_temp; V.final_temp = p->final_temp; V.half_life = p->temp_decay_half_life;
  V.playout_cap_percent = p->playout_cap_percent;
  V.history_enabled = p->history_enabled; V.tree_reuse = p->tree_reuse; V.root_fpu_zero = p->root_fpu_zero;
  V.shaped_dirichlet = p->shaped_dirichlet; V.policy_target_pruning = p->policy_target_pruning;
  V.playout_cap = p->playout_cap_randomization ? 1 : 0; V.eval_type = p->eval_type; V.rng_mode = p->rng_mode;
  V.resign_percent = p->resign_percent; V.resign_playthrough_percent = p->resign_playthrough_percent;
  V.gumbel_enabled = p->gumbel_enabled; V.gumbel_full = p->gumbel_full; V.fast_search_uses_gumbel = p->fast_search_uses_gumbel;
  V.gumbel_m = p->gumbel_m; V.gumbel_c_visit = p->gumbel_c_visit; V.gumbel_c_scale = p->gumbel_c_scale;
  V.hit_cap = 64u;
  if (const char* hc = getenv("B2AZ_HIT_CAP")) V.hit_cap = (u32)std::max(0, atoi(hc));  // experiment knob
  V.slot_quota = p->per_slot_quota ? p->games_to_play / p->concurrent_games : 0u;

  // ---- pool sizing: 192 B blocks (8 child nodes each), pages of 64 blocks
  u64 max_visits = (u64)std::max(p->mcts_visits[0], p->mcts_visits[1]);
  for (u32 pm = 0; pm < pt.n_perms; ++pm) max_visits = std::max<u64>(max_visits, std::max(pt.visits[pm][0], pt.visits[pm][1]));
  // a search adds <= one block per simulation; budget = 4 searches' worth + slack per tree
  const u64 tree_blocks = 4ull * max_visits + 2ull * kPageBlocks;
  u64 pool_blocks = p->pool_nodes ? (p->pool_nodes + kKMax - 1) / kKMax : tree_blocks * (u64)G * kP;
#ifndef B2AZ_HOST_EMU
  if (p->pool_nodes == 0) {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const u64 cap = (u64)(free_b * 0.6) / sizeof(Block);
    if (pool_blocks > cap) pool_blocks = cap;
  }
#endif
  u64 pages = (pool_blocks + kPageBlocks - 1) / kPageBlocks;
  const u64 min_pages = (u64)G * kP + 16ull;  // every tree can at least hold one page
  if (pages < min_pages) pages = min_pages;
  if (pages > (0xFFFFFFF0ull >> kPageLog2)) pages = 0xFFFFFFF0ull >> kPageLog2;  // block index must fit 32 bits
  // regions: one per group of slots that a persistent CTA owns (see EngineView::ring_tickets)
  {
    const u32 sms = (u32)e->num_sms;
    u32 groups;
    if (G <= sms * (u32)kQGames) groups = std::max(1u, std::min(sms, (G + 31u) / 32u));
    else groups = (((G + (u32)kQGames - 1u) / (u32)kQGames + sms - 1u) / sms) * sms;
    e->group_games = (G + groups - 1u) / groups;
    e->groups = (G + e->group_games - 1u) / e->group_games;
    V.region_games = e->group_games;
    V.n_regions = e->groups;
    V.region_pages = (u32)std::max<u64>(pages / V.n_regions, (u64)V.region_games * kP + 2ull);
    pages = (u64)V.region_pages * V.n_regions;
  }
  V.num_pages = (u32)pages;
  // compact a tree at a move once it holds more than half of its share of the pool
  V.compact_pages = p->compact_pages ? std::min<u32>(p->compact_pages, 60000u)
                                     : (u32)std::max<u64>(4, std::min<u64>(pages / ((u64)G * kP) / 2, 60000));
  const size_t nblocks = (size_t)pages * kPageBlocks;
  V.hist_capacity = p->history_capacity ? p->history_capacity : std::max<u32>(1u << 16, G * (u32)kMaxHist);

  int rc = 0;
  auto A = [&](int r) { if (r && !rc) rc = r; };
  A(dev_alloc_raw(&V.blocks, nblocks));  // every block is fully written before it is read: no memset
  A(dev_alloc(&V.page_next, pages)); A(dev_alloc(&V.ring, pages));
  A(dev_alloc(&V.ring_tickets, (size_t)V.n_regions * 2));
  A(dev_alloc(&V.trees, (size_t)G * kP)); A(dev_alloc(&V.games, (size_t)G)); A(dev_alloc(&V.cold, (size_t)G));
  if (p->gumbel_enabled) A(dev_alloc(&V.gum, (size_t)G * kP));  // zero = reset state, no sims target
  A(dev_alloc(&V.path, (size_t)G * kMaxPath)); A(dev_alloc(&V.pslot, (size_t)G * kMaxPath));
  A(dev_alloc(&V.leaf_p0, (size_t)G)); A(dev_alloc(&V.leaf_p1, (size_t)G));
  A(dev_alloc(&V.leaf_player, (size_t)G)); A(dev_alloc(&V.leaf_game, (size_t)G)); A(dev_alloc(&V.leaf_seat, (size_t)G));
  if (pt.n_perms > 1u || pt.random_groups != 0u) {
    PermTables* dpt = nullptr;
    A(dev_alloc(&dpt, 1));
    V.perms = dpt;
    A(copy_h2d(dpt, &pt, sizeof(pt), nullptr));
    A(stream_sync(nullptr));
  }
  A(dev_alloc(&V.hist_partial, p->history_enabled ? (size_t)G * kMaxHist : 1));
  A(dev_alloc(&V.hist_out, p->history_enabled ? (size_t)V.hist_capacity : 1));
  A(dev_alloc(&V.glob, 1));
  if (p->max_cache_size != 0 && p->eval_type == B2AZ_EVAL_NN) {
    // per-model-group cache of max_cache_size entries, ghost set of 9/10 of that (play_manager.cc:195-203)
    V.cache_buckets = std::max<u32>(1u, (p->max_cache_size + kCacheWays - 1) / kCacheWays);
    V.cache_ghost_slots = std::max<u32>(1u, (u32)((u64)p->max_cache_size * 9ull / 10ull));
    A(dev_alloc(&V.cache_keys, (size_t)V.cache_buckets * kCacheWays));
    A(dev_alloc(&V.cache_meta, (size_t)V.cache_buckets));
    A(dev_alloc(&V.cache_lock, (size_t)V.cache_buckets));
    A(dev_alloc(&V.cache_vals, (size_t)V.cache_buckets * kCacheWays));
    A(dev_alloc(&V.cache_ghost, (size_t)V.cache_ghost_slots));
    A(dev_alloc(&V.leaf_key, (size_t)G));
    A(dev_alloc(&V.hit_val, (size_t)G));
  }
  A(dev_alloc(&e->canon_buf, (size_t)G * C4_CANON));
  A(dev_alloc(&e->ev_v_buf, (size_t)G * (kP + 1))); A(dev_alloc(&e->ev_pi_buf, (size_t)G * kA));
  A(dev_alloc(&e->peek_buf, 1)); A(dev_alloc(&e->stats_buf, 1)); A(dev_alloc(&e->freepages_buf, 1));
  if (rc) { b2az_destroy(e); return rc; }
  V.ev_v = e->ev_v_buf;
  V.ev_pi = e->ev_pi_buf;

  // ---- init
  if (V.hit_val) {
#ifndef B2AZ_HOST_EMU
    CUDA_TRY(cudaMemset(V.hit_val, 0xFF, (size_t)G * sizeof(u32)));
#else
    memset(V.hit_val, 0xFF, (size_t)G * sizeof(u32));
#endif
  }
#ifndef B2AZ_HOST_EMU
  if (const char* cv = getenv("B2AZ_CARVEOUT"))  // experiment knob: shared-memory carveout of the step kernel, percent
    CUDA_TRY(cudaFuncSetAttribute(k_step<false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
  CUDA_TRY(cudaFuncSetAttribute(k_step_q<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QShared)));
  CUDA_TRY(cudaFuncSetAttribute(k_step_q<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QShared)));
  CUDA_TRY(cudaFuncSetAttribute(k_step_w<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WShared)));
  CUDA_TRY(cudaFuncSetAttribute(k_step_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WShared)));
  InitArgs ia{p->seed, p->rng_mode};
  k_init_pool<<<e->num_sms * 4, 256>>>(V);
  k_init_games<<<e->num_sms * 4, 256>>>(V, ia);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
#else
  for (u32 pg = 0; pg < V.num_pages; ++pg) {
    V.page_next[pg] = kNil;
    V.ring[pg] = pg;
  }
  for (u32 g = 0; g < G; ++g) {
    GameSlot gs;
    memset(&gs, 0, sizeof(gs));
    gs.active = 1;
    pcg32_seed(gs.rng, p->seed + (u64)g);
    V.games[g] = gs;
    memset(&V.cold[g], 0, sizeof(GameCold));
    TreeHdr T;
    memset(&T, 0, sizeof(T));
    tree_reset(T);
    T.region = g / V.region_games;
    V.trees[(size_t)g * kP + 0] = T;
    V.trees[(size_t)g * kP + 1] = T;
  }
  for (u32 r = 0; r < V.n_regions; ++r) {
    V.ring_tickets[2u * r] = 0;
    V.ring_tickets[2u * r + 1u] = V.region_pages;
  }
  memset(V.glob, 0, sizeof(Globals));
  V.glob->games_started = G;
  V.glob->active_games = G;
  pcg32_seed(V.glob->global_rng, p->seed);
#endif
  e->row_of_game.assign(G, 0);
  *out = e;
  return 0;
}

int b2az_step(b2az_engine* e, uint32_t n_steps, void* stream) {
  if (!e) return fail(B2AZ_EINVAL, "null engine");
  if (n_steps == 0) return 0;
  stream_t s = static_cast<stream_t>(stream);
  EngineView& V = e->view;
  if (int rc = bind_device(e)) return rc;
  if (V.eval_type == B2AZ_EVAL_NN) {
    if (n_steps != 1) return fail(B2AZ_EINVAL, "n_steps must be 1 with B2AZ_EVAL_NN");
    if (e->leaves_pending && !e->evals_all) {
      if (int rc = sync_leaf_count(e, s)) return rc;
      if (e->evals_submitted < e->leaf_count)
        return fail(B2AZ_ESTATE, "b2az_step: leaves of the previous step are still waiting for b2az_submit_eval");
    }
    if (V.cache_buckets && e->leaves_pending) {
      // the evaluations of the previous generation enter the cache before anybody looks anything up
#ifndef B2AZ_HOST_EMU
      k_cache_insert<<<std::max(1u, std::min((V.G + 127u) / 128u, (u32)e->num_sms * 8u)), 128, 0, s>>>(V);
      CUDA_TRY(cudaGetLastError());
#else
      for (u32 r = 0; r < V.glob->leaf_count; ++r)
        cache_insert(V, V.leaf_key[r], V.ev_v + (size_t)r * (kP + 1), V.ev_pi + (size_t)r * kA);
#endif
    }
    if (int rc = dev_zero(&V.glob->leaf_count, sizeof(u32), s)) return rc;
  }
#ifndef B2AZ_HOST_EMU
  if (V.rng_mode == B2AZ_RNG_GLOBAL) {
    k_step_serial<<<1, 1, 0, s>>>(V, n_steps);  // exactly ONE thread walks the slots
  } else if (e->step_kernel == B2AZ_STEP_WAVES) {
    // persistent CTAs (one per SM) owning groups of <= kQGames slots, scheduled in waves (az_engine_waves.h)
    const u32 grid = std::min(e->groups, (u32)e->num_sms);
    if (V.gumbel_enabled) k_step_w<true><<<grid, B2AZ_W_WARPS * 32, sizeof(WShared), s>>>(V, n_steps, e->group_games, e->groups);
    else k_step_w<false><<<grid, B2AZ_W_WARPS * 32, sizeof(WShared), s>>>(V, n_steps, e->group_games, e->groups);
  } else if (e->step_kernel == B2AZ_STEP_QUEUE) {
    // persistent CTAs (one per SM), each owning groups of <= kQGames slots whose state lives in shared memory
    const u32 grid = std::min(e->groups, (u32)e->num_sms);
    if (V.gumbel_enabled) k_step_q<true><<<grid, B2AZ_Q_WARPS * 32, sizeof(QShared), s>>>(V, n_steps, e->group_games, e->groups);
    else k_step_q<false><<<grid, B2AZ_Q_WARPS * 32, sizeof(QShared), s>>>(V, n_steps, e->group_games, e->groups);
  } else if (e->step_kernel == B2AZ_STEP_SYNC) {
    const u32 threads = V.G <= (u32)e->num_sms * 32u * 32u ? 32u : 64u;
    const u32 blocks = (V.G + threads - 1) / threads;
    const bool px = V.perms != nullptr;  // seat permutations / a RANDOM group next to an NN one
    if (px) {
      if (V.gumbel_enabled) k_step_sync<true, true><<<blocks, threads, 0, s>>>(V, n_steps);
      else k_step_sync<false, true><<<blocks, threads, 0, s>>>(V, n_steps);
    } else {
      if (V.gumbel_enabled) k_step_sync<true, false><<<blocks, threads, 0, s>>>(V, n_steps);
      else k_step_sync<false, false><<<blocks, threads, 0, s>>>(V, n_steps);
    }
  } else {
    // small CTAs spread the (one thread per game) population evenly over the SMs
    const u32 threads = V.G <= (u32)e->num_sms * 32u * 32u ? 32u : 64u;
    const u32 blocks = (V.G + threads - 1) / threads;
    if (V.gumbel_enabled) k_step<true><<<blocks, threads, 0, s>>>(V, n_steps);
    else k_step<false><<<blocks, threads, 0, s>>>(V, n_steps);
  }
  CUDA_TRY(cudaGetLastError());
#else
  // B2AZ_RNG_GLOBAL needs slot-major order inside a step (matches k_step_serial); per-game RNG is
  // order independent, so the same loop serves both.
  // Test build only. Default: slot-major inside every step (the order of k_step_serial, and the one in
  // which a single reference worker thread visits the slots). B2AZ_EMU_FLAT=1 runs the fused kernel's
  // flattened loop game by game instead (what one GPU thread does), to exercise run_flat() on the CPU.
  const char* flat = getenv("B2AZ_EMU_FLAT");
  if (!(flat && flat[0] == '1' && V.rng_mode == B2AZ_RNG_PER_GAME)) {
    for (u32 st = 0; st < n_steps; ++st)
      for (u32 g = 0; g < V.G; ++g) {
        Ctx c;
        ctx_load(V, g, c);
        if (!c.gs.active) continue;
        game_step(V, g, c);
        ctx_store(V, g, c);
      }
  } else {
    for (u32 g = 0; g < V.G; ++g) {
      Ctx c;
      ctx_load(V, g, c);
      if (!c.gs.active) continue;
      run_flat(V, g, c, n_steps);
      ctx_store(V, g, c);
    }
  }
#endif
  e->started = true;
  if (V.eval_type == B2AZ_EVAL_NN) {
    e->leaves_pending = true;
    e->leaf_count_known = false;
    e->canon_ready = false;
    e->leaves_taken = 0;
    e->evals_submitted = 0;
    e->evals_all = false;
    V.ev_v = e->ev_v_buf;
    V.ev_pi = e->ev_pi_buf;
  }
  return 0;
}

int b2az_leaf_batch(b2az_engine* e, void* stream, uint32_t* count, const float** canon_dev, const uint32_t** ids_dev) {
  if (!e || !count) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  if (e->view.eval_type != B2AZ_EVAL_NN) return fail(B2AZ_ESTATE, "leaf batches only exist with B2AZ_EVAL_NN");
  if (!e->leaves_pending) { *count = 0; return 0; }
  if (int rc = ensure_canon(e, s)) return rc;
  if (int rc = sync_leaf_count(e, s)) return rc;
  if (int rc = check_device_error(e, s)) return rc;
  *count = e->leaf_count;
  if (canon_dev) *canon_dev = e->canon_buf;
  if (ids_dev) *ids_dev = e->view.leaf_game;
  return 0;
}

int b2az_leaf_batch_host(b2az_engine* e, void* stream, uint32_t max, float* canon_host, uint32_t* ids_host,
                         uint32_t* count) {
  if (!e || !count || !canon_host || !ids_host) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  if (e->view.eval_type != B2AZ_EVAL_NN) return fail(B2AZ_ESTATE, "leaf batches only exist with B2AZ_EVAL_NN");
  *count = 0;
  if (!e->leaves_pending) return 0;
  if (int rc = ensure_canon(e, s)) return rc;
  if (int rc = sync_leaf_count(e, s)) return rc;
  if (int rc = check_device_error(e, s)) return rc;
  if (e->leaf_ids_host.size() != e->leaf_count || e->leaves_taken == 0) {
    e->leaf_ids_host.resize(e->leaf_count);
    if (e->leaf_count) {
      if (int rc = copy_d2h(e->leaf_ids_host.data(), e->view.leaf_game, (size_t)e->leaf_count * 4, s)) return rc;
      if (int rc = stream_sync(s)) return rc;
    }
    for (u32 r = 0; r < e->leaf_count; ++r) e->row_of_game[e->leaf_ids_host[r]] = r;
  }
  const u32 n = std::min(max, e->leaf_count - e->leaves_taken);
  if (n == 0) return 0;
  if (int rc = copy_d2h(canon_host, e->canon_buf + (size_t)e->leaves_taken * C4_CANON, (size_t)n * C4_CANON * 4, s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  memcpy(ids_host, e->leaf_ids_host.data() + e->leaves_taken, (size_t)n * 4);
  e->leaves_taken += n;
  *count = n;
  return 0;
}

int b2az_cache_insert_host(b2az_engine* e, void* stream, const uint64_t* keys_host, const float* v_host, const float* pi_host,
                           uint32_t n) {
  if (!e || (n && (!keys_host || !v_host || !pi_host))) return fail(B2AZ_EINVAL, "null argument");
  if (!e->view.cache_buckets) return fail(B2AZ_ESTATE, "the engine was created without a position cache (max_cache_size = 0)");
  for (u32 i = 0; i < n; ++i)
    if (keys_host[i] == 0) return fail(B2AZ_EINVAL, "key 0 is the empty-slot marker");
  stream_t s = static_cast<stream_t>(stream);
  if (int rc = bind_device(e)) return rc;
  if (n == 0) return 0;
#ifndef B2AZ_HOST_EMU
  u64* dk = nullptr; float *dv = nullptr, *dp = nullptr;
  int rc = dev_alloc_raw(&dk, n);
  if (!rc) rc = dev_alloc_raw(&dv, (size_t)n * (kP + 1));
  if (!rc) rc = dev_alloc_raw(&dp, (size_t)n * kA);
  if (!rc) rc = copy_h2d(dk, keys_host, (size_t)n * 8, s);
  if (!rc) rc = copy_h2d(dv, v_host, (size_t)n * (kP + 1) * 4, s);
  if (!rc) rc = copy_h2d(dp, pi_host, (size_t)n * kA * 4, s);
  if (!rc) {
    k_cache_insert_keys<<<1, 1, 0, s>>>(e->view, dk, dv, dp, n);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_cache_insert_keys launch failed");
  }
  if (!rc) rc = stream_sync(s);
  dev_free(dk); dev_free(dv); dev_free(dp);
  return rc;
#else
  (void)s;
  for (u32 i = 0; i < n; ++i) cache_insert(e->view, keys_host[i], v_host + (size_t)i * (kP + 1), pi_host + (size_t)i * kA);
  return 0;
#endif
}

int b2az_cache_find_host(b2az_engine* e, void* stream, const uint64_t* keys_host, uint32_t n, uint8_t* found_host, float* v_host,
                         float* pi_host) {
  if (!e || (n && (!keys_host || !found_host || !v_host || !pi_host))) return fail(B2AZ_EINVAL, "null argument");
  if (!e->view.cache_buckets) return fail(B2AZ_ESTATE, "the engine was created without a position cache (max_cache_size = 0)");
  stream_t s = static_cast<stream_t>(stream);
  if (int rc = bind_device(e)) return rc;
  if (n == 0) return 0;
#ifndef B2AZ_HOST_EMU
  u64* dk = nullptr; float *dv = nullptr, *dp = nullptr; u8* df = nullptr;
  int rc = dev_alloc_raw(&dk, n);
  if (!rc) rc = dev_alloc(&dv, (size_t)n * (kP + 1));
  if (!rc) rc = dev_alloc(&dp, (size_t)n * kA);
  if (!rc) rc = dev_alloc(&df, (size_t)n);
  if (!rc) rc = copy_h2d(dk, keys_host, (size_t)n * 8, s);
  if (!rc) {
    k_cache_find_keys<<<1, 1, 0, s>>>(e->view, dk, n, df, dv, dp);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2AZ_ECUDA, "k_cache_find_keys launch failed");
  }
  if (!rc) rc = copy_d2h(found_host, df, n, s);
  if (!rc) rc = copy_d2h(v_host, dv, (size_t)n * (kP + 1) * 4, s);
  if (!rc) rc = copy_d2h(pi_host, dp, (size_t)n * kA * 4, s);
  if (!rc) rc = stream_sync(s);
  dev_free(dk); dev_free(dv); dev_free(dp); dev_free(df);
  return rc;
#else
  (void)s;
  for (u32 i = 0; i < n; ++i) {
    const u32 idx = cache_find(e->view, keys_host[i]);
    found_host[i] = idx != kNil;
    if (idx != kNil) {
      memcpy(v_host + (size_t)i * (kP + 1), e->view.cache_vals[idx].v, sizeof(float) * (kP + 1));
      memcpy(pi_host + (size_t)i * kA, e->view.cache_vals[idx].pi, sizeof(float) * kA);
    }
  }
  return 0;
#endif
}

int b2az_leaf_seats_host(b2az_engine* e, void* stream, uint8_t* seats_host, uint32_t count) {
  if (!e || (count && !seats_host)) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (int rc = bind_device(e)) return rc;
  if (!e->leaves_pending) return fail(B2AZ_ESTATE, "no leaf batch");
  if (int rc = sync_leaf_count(e, s)) return rc;
  if (count > e->leaf_count) return fail(B2AZ_EINVAL, "more rows than leaves");
  if (count == 0) return 0;
  if (int rc = copy_d2h(seats_host, e->view.leaf_seat, count, s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  for (uint32_t i = 0; i < count; ++i) seats_host[i] &= 15u;  // (bits 4-7 carry the model group)
  return 0;
}
int b2az_leaf_groups_host(b2az_engine* e, void* stream, uint8_t* groups_host, uint32_t count) {
  if (!e || (count && !groups_host)) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (int rc = bind_device(e)) return rc;
  if (!e->leaves_pending) return fail(B2AZ_ESTATE, "no leaf batch");
  if (int rc = sync_leaf_count(e, s)) return rc;
  if (count > e->leaf_count) return fail(B2AZ_EINVAL, "more rows than leaves");
  if (count == 0) return 0;
  if (int rc = copy_d2h(groups_host, e->view.leaf_seat, count, s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  for (uint32_t i = 0; i < count; ++i) groups_host[i] >>= 4;
  return 0;
}

int b2az_submit_eval(b2az_engine* e, const float* v_dev, const float* pi_dev, uint32_t count) {
  if (!e || (count && (!v_dev || !pi_dev))) return fail(B2AZ_EINVAL, "null argument");
  if (!e->leaves_pending) return fail(B2AZ_ESTATE, "no leaf batch is waiting for evaluations");
  if (!e->leaf_count_known) return fail(B2AZ_ESTATE, "call b2az_leaf_batch before b2az_submit_eval");
  if (count != e->leaf_count) return fail(B2AZ_EINVAL, "b2az_submit_eval: count must equal the leaf count");
  e->view.ev_v = v_dev;
  e->view.ev_pi = pi_dev;
  e->evals_submitted = count;
  return 0;
}

int b2az_leaf_batch_device(b2az_engine* e, void* stream, const float** canon_dev, const uint32_t** ids_dev,
                           const uint32_t** count_dev) {
  if (!e) return fail(B2AZ_EINVAL, "null engine");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  if (e->view.eval_type != B2AZ_EVAL_NN) return fail(B2AZ_ESTATE, "leaf batches only exist with B2AZ_EVAL_NN");
  if (!e->leaves_pending) return fail(B2AZ_ESTATE, "call b2az_step first");
  if (int rc = ensure_canon(e, s)) return rc;
  if (canon_dev) *canon_dev = e->canon_buf;
  if (ids_dev) *ids_dev = e->view.leaf_game;
  if (count_dev) *count_dev = &e->view.glob->leaf_count;
  return 0;
}

int b2az_leaf_valid_device(b2az_engine* e, void* stream, const uint8_t** valid_dev) {
  if (!e || !valid_dev) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (int rc = bind_device(e)) return rc;
  if (e->view.eval_type != B2AZ_EVAL_NN) return fail(B2AZ_ESTATE, "leaf batches only exist with B2AZ_EVAL_NN");
  if (!e->leaves_pending) return fail(B2AZ_ESTATE, "call b2az_step first");
#ifdef B2AZ_HOST_EMU
  (void)s;
  return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
#else
  if (!e->valid_buf)
    if (int rc = dev_alloc(&e->valid_buf, (size_t)e->view.G * 7)) return rc;
  k_leaf_valid<<<e->num_sms * 4, 256, 0, s>>>(e->view, &e->view.glob->leaf_count, e->valid_buf);
  CUDA_TRY(cudaGetLastError());
  *valid_dev = e->valid_buf;
  return 0;
#endif
}

int b2az_submit_eval_all(b2az_engine* e, const float* v_dev, const float* pi_dev) {
  if (!e || !v_dev || !pi_dev) return fail(B2AZ_EINVAL, "null argument");
  if (!e->leaves_pending) return fail(B2AZ_ESTATE, "no leaf batch is waiting for evaluations");
  e->view.ev_v = v_dev;
  e->view.ev_pi = pi_dev;
  e->evals_all = true;
  return 0;
}

int b2az_submit_eval_host(b2az_engine* e, void* stream, const uint32_t* ids_host, const float* v_host,
                          const float* pi_host, uint32_t count) {
  if (!e || (count && (!ids_host || !v_host || !pi_host))) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  if (!e->leaves_pending || !e->leaf_count_known) return fail(B2AZ_ESTATE, "no leaf batch is waiting for evaluations");
  if (e->evals_submitted + count > e->leaf_count) return fail(B2AZ_EINVAL, "more evaluations than leaves");
  // contiguous run of rows? (the common case: ids come straight from b2az_leaf_batch_host)
  bool contiguous = true;
  for (u32 i = 0; i < count; ++i) {
    if (ids_host[i] >= e->view.G) return fail(B2AZ_EINVAL, "bad slot id");
    if (e->row_of_game[ids_host[i]] != e->row_of_game[ids_host[0]] + i) contiguous = false;
  }
  if (count == 0) return 0;
  if (contiguous) {
    const u32 r0 = e->row_of_game[ids_host[0]];
    if (int rc = copy_h2d(e->ev_v_buf + (size_t)r0 * (kP + 1), v_host, (size_t)count * (kP + 1) * 4, s)) return rc;
    if (int rc = copy_h2d(e->ev_pi_buf + (size_t)r0 * kA, pi_host, (size_t)count * kA * 4, s)) return rc;
  } else {
    for (u32 i = 0; i < count; ++i) {
      const u32 r = e->row_of_game[ids_host[i]];
      if (int rc = copy_h2d(e->ev_v_buf + (size_t)r * (kP + 1), v_host + (size_t)i * (kP + 1), (kP + 1) * 4, s)) return rc;
      if (int rc = copy_h2d(e->ev_pi_buf + (size_t)r * kA, pi_host + (size_t)i * kA, kA * 4, s)) return rc;
    }
  }
  if (int rc = stream_sync(s)) return rc;  // the host buffers may be reused by the caller
  e->view.ev_v = e->ev_v_buf;
  e->view.ev_pi = e->ev_pi_buf;
  e->evals_submitted += count;
  return 0;
}

static int drain_history_impl(b2az_engine* e, void* stream, uint32_t max, u32 nsym, float* canon, float* v, float* pi,
                              int dst_is_device, uint32_t* count) {
  if (!e || !count) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  *count = 0;
  if (!e->params.history_enabled || max == 0) return 0;
  unsigned long long wr[2];
  if (int rc = copy_d2h(wr, &e->view.glob->hist_written, sizeof(wr), s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  const unsigned long long avail = wr[0] - wr[1];
  const u32 n = (u32)std::min<unsigned long long>(avail, max);
  if (n == 0) return 0;
  const size_t rows = (size_t)n * nsym;
  float *dc = canon, *dv = v, *dp = pi;
#ifndef B2AZ_HOST_EMU
  if (!dst_is_device) {
    if (e->hist_stage_cap < rows) {
      dev_free(e->hist_canon); dev_free(e->hist_v); dev_free(e->hist_pi);
      e->hist_canon = e->hist_v = e->hist_pi = nullptr;
      const u32 cap = (u32)std::max<size_t>(rows, 4096);
      if (int rc = dev_alloc(&e->hist_canon, (size_t)cap * C4_CANON)) return rc;
      if (int rc = dev_alloc(&e->hist_v, (size_t)cap * 3)) return rc;
      if (int rc = dev_alloc(&e->hist_pi, (size_t)cap * kA)) return rc;
      e->hist_stage_cap = cap;
    }
    dc = e->hist_canon; dv = e->hist_v; dp = e->hist_pi;
  }
  k_hist_expand<<<e->num_sms * 4, 256, 0, s>>>(e->view, wr[1], n, nsym, dc, dv, dp);
  CUDA_TRY(cudaGetLastError());
  if (!dst_is_device) {
    if (int rc = copy_d2h(canon, dc, rows * C4_CANON * 4, s)) return rc;
    if (int rc = copy_d2h(v, dv, rows * 3 * 4, s)) return rc;
    if (int rc = copy_d2h(pi, dp, rows * kA * 4, s)) return rc;
  }
#else
  (void)dst_is_device;
  for (size_t row = 0; row < rows; ++row) {
    const u32 sample = (u32)(row / nsym), mirror = (u32)(row % nsym);
    const HistEntry& h = e->view.hist_out[(wr[1] + sample) % (unsigned long long)e->view.hist_capacity];
    for (u32 el = 0; el < (u32)C4_CANON; ++el)
      dc[row * C4_CANON + el] = c4_canon_elem(h.p0, h.p1, h.player, mirror ? c4_mirror_elem(el) : el);
    for (u32 el = 0; el < 3u; ++el) dv[row * 3 + el] = (h.result == el + 1u) ? 1.0f : 0.0f;
    for (u32 el = 0; el < (u32)kA; ++el) dp[row * kA + el] = h.pi[mirror ? (u32)kA - 1u - el : el];
  }
#endif
  const unsigned long long new_read = wr[1] + n;
  if (int rc = copy_h2d(&e->view.glob->hist_read, &new_read, sizeof(new_read), s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  *count = n;
  return 0;
}

int b2az_drain_history(b2az_engine* e, void* stream, uint32_t max, float* canon, float* v, float* pi,
                       int dst_is_device, uint32_t* count) {
  return drain_history_impl(e, stream, max, 1u, canon, v, pi, dst_is_device, count);
}
int b2az_drain_history_sym(b2az_engine* e, void* stream, uint32_t max, float* canon, float* v, float* pi,
                           int dst_is_device, uint32_t* count) {
  return drain_history_impl(e, stream, max, 2u, canon, v, pi, dst_is_device, count);
}

// ---- overlapped drain: the samples of step k leave the device while step k + 1 runs --------------------------------
// b2az_history_mark(e, s): on the stream the steps run on, AFTER step k has been enqueued: snapshots the sample counters
// (an async 16-byte copy into pinned memory + an event). Returns at once. b2az_drain_history_marked(e, s2, ...): on a
// SECOND stream: waits for the mark, expands the samples up to the mark and copies them out on s2 — step k + 1, already
// running on the first stream, only appends BEYOND the mark (a sample is complete before the counter a later mark reads
// can include it: marks are stream-ordered between two step launches), so nothing it writes is touched.
int b2az_history_mark(b2az_engine* e, void* stream) {
  if (!e) return fail(B2AZ_EINVAL, "null engine");
  if (int rc = bind_device(e)) return rc;
  stream_t s = static_cast<stream_t>(stream);
#ifndef B2AZ_HOST_EMU
  if (!e->hist_mark_host) {
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&e->hist_mark_host), 16));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&e->hist_read_host), 8));
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->hist_event = ev;
  }
  CUDA_TRY(cudaMemcpyAsync(e->hist_mark_host, &e->view.glob->hist_written, 16, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaEventRecord((cudaEvent_t)e->hist_event, s));
#else
  if (!e->hist_mark_host) {
    e->hist_mark_host = static_cast<unsigned long long*>(calloc(2, 8));
    e->hist_read_host = static_cast<unsigned long long*>(calloc(1, 8));
  }
  (void)s;
  memcpy(e->hist_mark_host, &e->view.glob->hist_written, 16);
#endif
  e->hist_marked = true;
  return 0;
}

int b2az_drain_history_marked(b2az_engine* e, void* stream2, uint32_t max, float* canon, float* v, float* pi, int dst_is_device,
                              uint32_t* count) {
  if (!e || !count) return fail(B2AZ_EINVAL, "null argument");
  *count = 0;
  if (!e->hist_marked) return fail(B2AZ_ESTATE, "call b2az_history_mark first");
  if (int rc = bind_device(e)) return rc;
  stream_t s = static_cast<stream_t>(stream2);
  if (!e->params.history_enabled || max == 0) return 0;
#ifndef B2AZ_HOST_EMU
  CUDA_TRY(cudaEventSynchronize((cudaEvent_t)e->hist_event));  // the host needs the counters
  CUDA_TRY(cudaStreamWaitEvent(s, (cudaEvent_t)e->hist_event, 0));
#endif
  const unsigned long long written = e->hist_mark_host[0];
  // the read counter is advanced by this call only (b2az_drain_history and this one must not be mixed within a run)
  const unsigned long long read = std::max(e->hist_mark_host[1], *e->hist_read_host);
  const u32 n = (u32)std::min<unsigned long long>(written - read, max);
  if (n == 0) return 0;
  float *dc = canon, *dv = v, *dp = pi;
#ifndef B2AZ_HOST_EMU
  if (!dst_is_device) {
    if (e->hist_stage_cap < n) {
      // (the staging buffers may still be in use by an earlier drain on another stream: finish it first)
      CUDA_TRY(cudaDeviceSynchronize());
      dev_free(e->hist_canon); dev_free(e->hist_v); dev_free(e->hist_pi);
      e->hist_canon = e->hist_v = e->hist_pi = nullptr;
      const u32 cap = (u32)std::max<size_t>(n, 4096);
      if (int rc = dev_alloc(&e->hist_canon, (size_t)cap * C4_CANON)) return rc;
      if (int rc = dev_alloc(&e->hist_v, (size_t)cap * 3)) return rc;
      if (int rc = dev_alloc(&e->hist_pi, (size_t)cap * kA)) return rc;
      e->hist_stage_cap = cap;
    }
    dc = e->hist_canon; dv = e->hist_v; dp = e->hist_pi;
  }
  k_hist_expand<<<e->num_sms * 4, 256, 0, s>>>(e->view, read, n, 1u, dc, dv, dp);
  CUDA_TRY(cudaGetLastError());
  if (!dst_is_device) {
    if (int rc = copy_d2h(canon, dc, (size_t)n * C4_CANON * 4, s)) return rc;
    if (int rc = copy_d2h(v, dv, (size_t)n * 3 * 4, s)) return rc;
    if (int rc = copy_d2h(pi, dp, (size_t)n * kA * 4, s)) return rc;
  }
#else
  (void)dst_is_device;
  for (size_t row = 0; row < n; ++row) {
    const HistEntry& h = e->view.hist_out[(read + row) % (unsigned long long)e->view.hist_capacity];
    for (u32 el = 0; el < (u32)C4_CANON; ++el) dc[row * C4_CANON + el] = c4_canon_elem(h.p0, h.p1, h.player, el);
    for (u32 el = 0; el < 3u; ++el) dv[row * 3 + el] = (h.result == el + 1u) ? 1.0f : 0.0f;
    for (u32 el = 0; el < (u32)kA; ++el) dp[row * kA + el] = h.pi[el];
  }
#endif
  *e->hist_read_host = read + n;
  // the device copy of the read counter only feeds the ring-overflow check of later game ends: stream-ordered, no wait
  if (int rc = copy_h2d(&e->view.glob->hist_read, e->hist_read_host, 8, s)) return rc;
  if (int rc = stream_sync(s)) return rc;  // the caller's host buffers are filled when this returns
  *count = n;
  return 0;
}

int b2az_perm_scores(b2az_engine* e, void* stream, b2az_perm_stats* out8, uint32_t* n_perms_out) {
  if (!e || !out8) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (int rc = bind_device(e)) return rc;
  Globals G;
  if (int rc = copy_d2h(&G, e->view.glob, sizeof(G), s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  const u32 P = e->n_perms;
  memset(out8, 0, P * sizeof(b2az_perm_stats));
  for (u32 pm = 0; pm < P; ++pm)
    for (int i = 0; i < 3; ++i) {  // perm_scores_ holds integer win counts (play_manager.cc:466-467)
      out8[pm].scores[i] = (float)G.perm_wins[pm][i];
      out8[pm].games_completed += (uint32_t)G.perm_wins[pm][i];
    }
  if (n_perms_out) *n_perms_out = P;
  return 0;
}

int b2az_get_stats(b2az_engine* e, void* stream, b2az_stats* out) {
  if (!e || !out) return fail(B2AZ_EINVAL, "null argument");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  memset(out, 0, sizeof(*out));
  Globals G;
  StatsOut so{0, 0};
  unsigned long long free_pages = 0;
#ifndef B2AZ_HOST_EMU
  if (int rc = dev_zero(e->stats_buf, sizeof(StatsOut), s)) return rc;
  if (int rc = dev_zero(e->freepages_buf, sizeof(unsigned long long), s)) return rc;
  k_stats<<<e->num_sms * 2, 256, 0, s>>>(e->view, e->stats_buf);
  k_count_free_pages<<<e->num_sms * 4, 256, 0, s>>>(e->view, e->freepages_buf);
  CUDA_TRY(cudaGetLastError());
  if (int rc = copy_d2h(&so, e->stats_buf, sizeof(so), s)) return rc;
  if (int rc = copy_d2h(&free_pages, e->freepages_buf, sizeof(free_pages), s)) return rc;
#else
  for (u32 g = 0; g < e->view.G; ++g) { so.sims += e->view.cold[g].sims; so.moves += e->view.cold[g].nmoves; }
  for (u32 i = 0; i < e->view.num_pages; ++i) free_pages += (e->view.ring[i] != kNil);
#endif
  if (int rc = copy_d2h(&G, e->view.glob, sizeof(G), s)) return rc;
  if (int rc = stream_sync(s)) return rc;
  out->simulations = so.sims;
  out->moves = so.moves;
  out->games_completed = G.games_completed;
  out->games_started = std::min(G.games_started, e->view.games_to_play);
  out->active_games = G.active_games;
  out->leaf_count = e->leaves_pending ? G.leaf_count : 0;
  out->hist_count = (u32)(G.hist_written - G.hist_read);
  for (int i = 0; i < 3; ++i) { out->scores[i] = (float)G.wins[i]; out->resign_scores[i] = (float)G.resign_wins[i]; }
  // getters of play_manager.h:288-315 (same float/double conversions)
  out->avg_game_length = (float)G.game_length / (float)G.games_completed;
  out->avg_leaf_depth = G.full_move_count ? (float)(G.total_avg_leaf_depth / (double)G.full_move_count) : 0.0f;
  out->avg_search_entropy = G.full_move_count ? (float)(G.total_search_entropy / (double)G.full_move_count) : 0.0f;
  out->fast_avg_leaf_depth = G.fast_move_count ? (float)(G.fast_total_avg_leaf_depth / (double)G.fast_move_count) : 0.0f;
  out->fast_avg_search_entropy = G.fast_move_count ? (float)(G.fast_total_search_entropy / (double)G.fast_move_count) : 0.0f;
  out->avg_moves_per_turn = G.game_length ? (float)G.total_move_count / (float)G.game_length : 0.0f;
  out->avg_valid_moves = G.total_move_count ? (float)(G.total_valid_moves / (double)G.total_move_count) : 0.0f;
  out->pool_pages_total = e->view.num_pages;
  out->pool_pages_free = free_pages;
  out->device_error = G.error;
  out->compactions = G.compactions;
  out->cache_hits = G.cache_hits; out->cache_misses = G.cache_misses; out->cache_evictions = G.cache_evictions;
  out->cache_reinserts = G.cache_reinserts; out->cache_size = G.cache_size;
  out->cache_max_size = (unsigned long long)e->view.cache_buckets * kCacheWays;
  out->sum_game_length = G.game_length;
  out->total_move_count = G.total_move_count; out->full_move_count = G.full_move_count; out->fast_move_count = G.fast_move_count;
  out->sum_leaf_depth = G.total_avg_leaf_depth; out->sum_search_entropy = G.total_search_entropy;
  out->fast_sum_leaf_depth = G.fast_total_avg_leaf_depth; out->fast_sum_search_entropy = G.fast_total_search_entropy;
  out->sum_valid_moves = G.total_valid_moves;
  return 0;
}

int b2az_set_games_to_play(b2az_engine* e, uint32_t games_to_play) {
  if (!e) return fail(B2AZ_EINVAL, "null engine");
  e->view.games_to_play = games_to_play;  // EngineView travels by value with every launch
  e->params.games_to_play = games_to_play;
  return 0;
}

int b2az_peek(b2az_engine* e, void* stream, uint32_t game, uint32_t seat, uint8_t* state89, uint32_t* counts7,
              float* root_q7, float* root_value3, uint32_t* depth, uint32_t* root_n, float* root_policy7) {
  if (!e) return fail(B2AZ_EINVAL, "null engine");
  if (game >= e->view.G || seat >= (u32)kP) return fail(B2AZ_EINVAL, "bad game/seat");
  stream_t s = static_cast<stream_t>(stream);
  if (e) { if (int rc = bind_device(e)) return rc; }
  PeekOut po;
#ifndef B2AZ_HOST_EMU
  k_peek<<<1, 1, 0, s>>>(e->view, game, seat, e->peek_buf);
  CUDA_TRY(cudaGetLastError());
  if (int rc = copy_d2h(&po, e->peek_buf, sizeof(po), s)) return rc;
  if (int rc = stream_sync(s)) return rc;
#else
  (void)s;
  peek_impl(e->view, game, seat, &po);
#endif
  if (state89) memcpy(state89, po.state, 89);
  if (counts7) memcpy(counts7, po.counts, sizeof(po.counts));
  if (root_q7) memcpy(root_q7, po.q, sizeof(po.q));
  if (root_policy7) memcpy(root_policy7, po.policy, sizeof(po.policy));
  if (root_value3) memcpy(root_value3, po.root_value, sizeof(po.root_value));
  if (depth) *depth = po.depth;
  if (root_n) *root_n = po.root_n;
  return 0;
}

int b2az_c4_batch(int device, uint32_t n, const int8_t* boards_host, const uint8_t* players, const uint32_t* turns,
                  const uint32_t* moves, int8_t* boards_out, uint8_t* players_out, uint8_t* valid, float* scores,
                  uint8_t* terminal, float* canonical, int32_t* status) {
  if (n == 0) return 0;
  if (!boards_host || !players || !turns) return fail(B2AZ_EINVAL, "null argument");
#ifdef B2AZ_HOST_EMU
  (void)device;
  C4BatchArgs a{n, reinterpret_cast<const signed char*>(boards_host), players, turns, moves,
                reinterpret_cast<signed char*>(boards_out), players_out, valid, scores, terminal, canonical, status};
  for (u32 i = 0; i < n; ++i) c4_batch_one(a, i);
  return 0;
#else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2AZ_ECUDA, "no CUDA device: libb2az has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  C4BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n;
  std::vector<void*> owned;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    owned.push_back(d);
    if (src) cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    return d;
  };
  a.boards = static_cast<const signed char*>(up(boards_host, (size_t)n * 84));
  a.players = static_cast<const u8*>(up(players, n));
  a.turns = static_cast<const u32*>(up(turns, (size_t)n * 4));
  a.moves = moves ? static_cast<const u32*>(up(moves, (size_t)n * 4)) : nullptr;
  a.boards_out = boards_out ? static_cast<signed char*>(up(nullptr, (size_t)n * 84)) : nullptr;
  a.players_out = players_out ? static_cast<u8*>(up(nullptr, n)) : nullptr;
  a.valid = valid ? static_cast<u8*>(up(nullptr, (size_t)n * kA)) : nullptr;
  a.scores = scores ? static_cast<float*>(up(nullptr, (size_t)n * 12)) : nullptr;
  a.terminal = terminal ? static_cast<u8*>(up(nullptr, n)) : nullptr;
  a.canonical = canonical ? static_cast<float*>(up(nullptr, (size_t)n * C4_CANON * 4)) : nullptr;
  a.status = status ? static_cast<i32*>(up(nullptr, (size_t)n * 4)) : nullptr;
  int rc = 0;
  k_c4_batch<<<std::max(1u, std::min((n + 255u) / 256u, 148u * 8u)), 256>>>(a);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) rc = fail(B2AZ_ECUDA, std::string("k_c4_batch: ") + cudaGetErrorString(err));
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (dst && src && !rc) cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
  };
  down(boards_out, a.boards_out, (size_t)n * 84);
  down(players_out, a.players_out, n);
  down(valid, a.valid, (size_t)n * kA);
  down(scores, a.scores, (size_t)n * 12);
  down(terminal, a.terminal, n);
  down(canonical, a.canonical, (size_t)n * C4_CANON * 4);
  down(status, a.status, (size_t)n * 4);
  for (void* d : owned) cudaFree(d);
  return rc;
#endif
}

}  // extern "C"

#include "az_tafl_kernels.h"
#include "az_stargambit_kernels.h"
#include "az_forest.h"
#include "az_selfplay.h"
