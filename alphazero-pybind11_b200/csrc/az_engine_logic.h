// az_engine_logic.h — the self-play hot path as device functions, ONE THREAD PER GAME SLOT.
//
// Per step a thread runs exactly one iteration of PlayManager::play()'s loop body
// (play_manager.cc:272-599) for its game:
//     process_result -> [play a move] -> find_leaf
// Why a thread and not a warp per tree: Connect4 has <= 7 children per node, and the per-simulation
// work is a short scalar chain whose float operation ORDER must equal the reference's sequential
// loops (SURVEY.md Appendix A "Float order") for visit counts to be bit-exact. Measured on the B200
// (profiles/r1_*): 8 or 32 cooperating lanes replicate that scalar chain in every lane (697 warp
// instructions per simulation, 29 % issue utilisation at 25 % occupancy) and lose to one thread per game
// by 2x. So the parallelism is ACROSS games (65,536 independent threads), memory-level parallelism
// comes from loading a node's whole child block (five 32 B sectors) in one burst, and the hot loop
// keeps every small array in registers (no indexed local arrays).
//
// Hot (inlined): find_leaf, process_result for interior leaves. Cold (AZ_COLD, out of line): playing a
// move, root priors (temperature, Dirichlet), re-rooting, Cheney compaction, the page ring.
//
// Everything here is __host__ __device__: the host instantiation exists only for the
// -DB2AZ_HOST_EMU test build (tests/cpp/libb2az_hostemu.so), which lets the CPU test-suite compare this
// logic with the oracle without a GPU. The product (az_engine.cu) only runs the device side and
// fails loudly without CUDA.
#pragma once

#include <string.h>

#include "az_engine_types.h"

namespace b2az {

// experiment knobs (profiles/): how many top levels of a tree are loaded with an L2 evict_last policy, and
// whether deeper levels / fresh blocks use evict_first
#ifndef B2AZ_L2_KEEP_LEVELS
#define B2AZ_L2_KEEP_LEVELS 0
#endif
#ifndef B2AZ_L2_STREAM_DEEP
#define B2AZ_L2_STREAM_DEEP 0
#endif
// lanes of a warp that must be ready before the warp runs a step boundary (run_flat); 0 = no gating
#ifndef B2AZ_GATE
#define B2AZ_GATE 0
#endif

#define B2AZ_DEVERR_POOL 1u
#define B2AZ_DEVERR_HIST 2u
#define B2AZ_DEVERR_MOVE 4u
#define B2AZ_DEVERR_DEPTH 8u
#define B2AZ_DEVERR_QUEUE 16u

// ------------------------------------------------------------------------------------ atomics
AZ_HD u32 at_add(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  u32 o = *p; *p += v; return o;
#endif
}
AZ_HD u32 at_sub(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
  return atomicSub(p, v);
#else
  u32 o = *p; *p -= v; return o;
#endif
}
AZ_HD void at_or(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
  atomicOr(p, v);
#else
  *p |= v;
#endif
}
AZ_HD u32 at_exch(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
  return atomicExch(p, v);
#else
  u32 o = *p; *p = v; return o;
#endif
}
AZ_HD u32 at_cas(u32* p, u32 cmp, u32 v) {
#if defined(__CUDA_ARCH__)
  return atomicCAS(p, cmp, v);
#else
  u32 o = *p; if (o == cmp) *p = v; return o;
#endif
}
AZ_HD unsigned long long at_add64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  unsigned long long o = *p; *p += v; return o;
#endif
}
AZ_HD void at_addd(double* p, double v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
AZ_HD void mem_fence() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#endif
}
// the seat permutation slot g plays (EngineView::n_perms)
AZ_HD u32 slot_perm(const EngineView& E, u32 g) { return E.perms ? g % E.perms->n_perms : 0u; }
// the search budget of `seat` in slot g (seat_visits_ / seat_cap_visits_[perm][seat], play_manager.cc:284-285)
template <bool PX = true>
AZ_HD u32 seat_budget(const EngineView& E, u32 g, u32 seat, bool capped) {
  if (PX && E.perms) {
    const u32 pm = g % E.perms->n_perms;
    return capped ? E.perms->cap_visits[pm][seat] : E.perms->visits[pm][seat];
  }
  return capped ? E.cap_visits[seat] : E.visits[seat];
}
// the model group that searches for `seat` in slot g (seat_perms_[perm][seat], play_manager.cc:577)
template <bool PX = true>
AZ_HD u32 seat_group_of(const EngineView& E, u32 g, u32 seat) {
  if (PX && E.perms) return E.perms->seat_group[g % E.perms->n_perms][seat];
  return E.seat_group[seat];
}
template <typename T>
AZ_HD T ld_volatile(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <typename T>
AZ_HD void st_volatile(T* p, T v) { *reinterpret_cast<volatile T*>(p) = v; }

// std::min / std::max argument-order semantics (they differ from fminf/fmaxf on NaN)
AZ_HD float std_min(float a, float b) { return (b < a) ? b : a; }
AZ_HD float std_max(float a, float b) { return (a < b) ? b : a; }

AZ_HD int popc32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// a / b for a finite positive b. Experiment (profiles/r2t_sync_ab.jsonl): giving a zero numerator (a terminal node's 0
// value) back without dividing keeps it off div.rn's slow path (4 % of the step kernel's instructions) but the extra
// compare + select on every division cost more than that: 2.41 G vs 2.50 G simulations/s. Kept as a knob.
AZ_HD float fdiv_pos(float a, float b) {
#if defined(B2AZ_FDIV_POS)
  return a == 0.0f ? a : fdiv(a, b);
#else
  return fdiv(a, b);
#endif
}

// ------------------------------------------------------------------------------------ 16 B vectors
struct V4 {
  u32 x, y, z, w;
};
AZ_HD V4 ld_v4(const void* p) {
  V4 r;
#if defined(__CUDA_ARCH__)
#if defined(B2AZ_LDCG)
  const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p));  // experiment: bypass L1
#else
  const uint4 a = *reinterpret_cast<const uint4*>(p);
#endif
  r.x = a.x; r.y = a.y; r.z = a.z; r.w = a.w;
#else
  memcpy(&r, p, 16);
#endif
  return r;
}
AZ_HD void st_v4(void* p, const V4& v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint4*>(p) = make_uint4(v.x, v.y, v.z, v.w);
#else
  memcpy(p, &v, 16);
#endif
}
// L2 residency hints (sm_80+: createpolicy + ld/st .L2::cache_hint). The top of every tree (root block and
// its children's blocks: 8 x 160 B x 65,536 games = 84 MB) is touched by every simulation and fits the
// 126 MB L2; deeper blocks are touched rarely and only pollute it.
AZ_HD u64 l2_policy_keep() {
#if defined(__CUDA_ARCH__)
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#else
  return 0;
#endif
}
AZ_HD u64 l2_policy_stream() {
#if defined(__CUDA_ARCH__)
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#else
  return 0;
#endif
}
AZ_HD V4 ld_v4_hint(const void* p, u64 pol) {
#if defined(__CUDA_ARCH__)
  V4 r;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol)
               : "memory");
  return r;
#else
  (void)pol;
  return ld_v4(p);
#endif
}
AZ_HD void st_v4_hint(void* p, const V4& v, u64 pol) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
               "l"(pol)
               : "memory");
#else
  (void)pol;
  st_v4(p, v);
#endif
}
AZ_HD V4 mk_v4(u32 x, u32 y, u32 z, u32 w) {
  V4 r;
  r.x = x; r.y = y; r.z = z; r.w = w;
  return r;
}
// scalar accessors (cold paths)
AZ_HD u32 blk_n(const Block* B, u32 j) { return B->rec[j][0]; }
AZ_HD float blk_q(const Block* B, u32 j) { return u2f(B->rec[j][1]); }
AZ_HD float blk_pol(const Block* B, u32 j) { return u2f(B->rec[j][2]); }
AZ_HD float blk_d(const Block* B, u32 j) { return u2f(B->rec[j][3]); }
AZ_HD u32 blk_mv(const Block* B, u32 j) { return (B->mv >> (4u * j)) & 15u; }
AZ_HD u32 blk_term(const Block* B, u32 j) { return (B->term >> (2u * j)) & 3u; }
AZ_HD float blk_v(const Block* B) { return u2f(B->v); }
AZ_HD u32 blk_k(const Block* B) { return B->kp & 0xFFu; }
AZ_HD u32 blk_player(const Block* B) { return (B->kp >> 8) & 0xFFu; }
AZ_HD void blk_copy(Block* D, const Block* S) {
  V4 t[10];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 10; ++i) t[i] = ld_v4(reinterpret_cast<const u32*>(S) + 4 * i);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 10; ++i) st_v4(reinterpret_cast<u32*>(D) + 4 * i, t[i]);
}

// ------------------------------------------------------------------------------------ page ring
// Free pages sit in one ring of page ids. pop and push each take a ticket with one atomicAdd and then
// own ring[ticket % num_pages]; the slot itself is handed over with atomicExch / atomicCAS, so there
// is no retry storm under contention (the first version's tagged-CAS stacks cost 12-28 ms in a
// generation where every game re-rooted). A slot carries everything (the page id), so no memory fence
// is needed: a gpu-scope __threadfence compiles to MEMBAR + CCTL.IVALL, which throws away the whole
// SM's L1 on every page pop (profiles/r3: L1 hit rate 43 %).
AZ_COLD void pool_push_page(const EngineView E, u32 page) {
  const u32 region = page / E.region_pages;  // a page always goes back to the region it came from
  const unsigned long long t = at_add64(&E.ring_tickets[2u * region + 1u], 1ULL);
  u32* slot = &E.ring[(size_t)region * E.region_pages + (size_t)(t % (unsigned long long)E.region_pages)];
#if defined(__CUDA_ARCH__)
  for (u32 spin = 0; spin < (1u << 22); ++spin)
    if (at_cas(slot, kNil, page) == kNil) return;
  at_or(&E.glob->error, B2AZ_DEVERR_POOL);  // cannot happen: there are never more free pages than pages
#else
  *slot = page;
#endif
}
AZ_COLD u32 pool_pop_page(const EngineView E, u32 region) {
  Globals* G = E.glob;
  if (ld_volatile(&G->error) & B2AZ_DEVERR_POOL) return kNil;  // already fatal: do not spin again
  const unsigned long long t = at_add64(&E.ring_tickets[2u * region], 1ULL);
  u32* slot = &E.ring[(size_t)region * E.region_pages + (size_t)(t % (unsigned long long)E.region_pages)];
  u32 page = kNil;
#if defined(__CUDA_ARCH__)
  for (u32 spin = 0; spin < (1u << 16); ++spin) {
    page = at_exch(slot, kNil);
    if (page != kNil) break;
  }
#else
  page = at_exch(slot, kNil);
#endif
  return page;  // kNil: ring empty = the region is exhausted (fatal, reported by the caller)
}
// Give every page of a tree's chain back (the chain links are this thread's own writes).
AZ_COLD void pool_push_chain(const EngineView E, u32 head) {
  u32 p = head;
  for (u32 guard = 0; p != kNil && guard < 0x10000u; ++guard) {
    const u32 nx = E.page_next[p];
    pool_push_page(E, p);
    p = nx;
  }
}
// Bump-allocate one block in the tree's arena.
AZ_HD u32 tree_alloc_block(const EngineView& E, TreeHdr& T) {
  if (T.cur_page == kNil || T.bump >= kPageBlocks) {
    const u32 p = pool_pop_page(E, T.region);
    if (p == kNil) {
      at_or(&E.glob->error, B2AZ_DEVERR_POOL);
      return kNil;
    }
    E.page_next[p] = kNil;
    if (T.cur_page != kNil) E.page_next[T.cur_page] = p;
    else T.first_page = p;
    T.cur_page = p;
    T.bump = 0;
    ++T.pages_used;
  }
  return (T.cur_page << kPageLog2) + T.bump++;
}
AZ_HD void arena_clear(TreeHdr& T) {
  T.first_page = T.cur_page = kNil;
  T.bump = 0;
  T.pages_used = 0;
}
AZ_HD void tree_free_pages(const EngineView& E, TreeHdr& T) {
  if (T.first_page != kNil) pool_push_chain(E, T.first_page);
  arena_clear(T);
}
AZ_HD void tree_reset(TreeHdr& T) {  // a freshly constructed MCTS (mcts.h:52-73): root_ = Node{}
  T.q = T.d = T.v = T.policy = 0.0f;
  T.n = 0;
  T.fc = kNil;
  T.move = 0;
  T.k = 0;
  T.player = 0;
  T.term = 0;
  T.leaf_term = T.leaf_k = T.leaf_player = 0;
  T.leaf_blk = kNil;
  T.path_len = 0;
  T.depth = 0;
  T.total_leaf_depth = 0;
  arena_clear(T);  // T.region stays: it belongs to the slot, not to the search
}

// ------------------------------------------------------------------------------------ path cache
// The first kPathRegs edges of the selection path also stay in registers between find_leaf and the
// process_result of the next fused step (the HBM copy in E.path / E.pslot is what a later launch reads).
constexpr int kPathRegs = 8;
struct PathRegs {
  u32 blk[kPathRegs];
  u32 slots_lo, slots_hi;  // 8 bits per level: child slot | parent's player << 4
  u32 valid;               // the registers describe the pending leaf's path
};
// The queue kernel (az_engine_queue.h) keeps the first kPathSm edges in SHARED memory instead, where an indexed
// access is one LDS / STS and not a chain of selects.
constexpr int kPathSm = 12;
struct PathSm {
  u32 blk[kPathSm];
  u8 slot[kPathSm];
  u32 valid;
};
AZ_HD void path_begin(PathRegs& pr) {
  pr.slots_lo = pr.slots_hi = 0;
  pr.valid = 1;
}
AZ_HD void path_begin(PathSm& pr) { pr.valid = 1; }
// record edge `plen` of the selection path: (block, child slot | parent's player << 4)
AZ_HD void path_put(const EngineView& E, u32 g, PathRegs& pr, u32 plen, u32 blk, u32 slot_byte) {
  if (plen >= (u32)kPathRegs) {  // the first kPathRegs levels reach HBM in ctx_store only
    E.path[(size_t)g * kMaxPath + plen] = blk;
    E.pslot[(size_t)g * kMaxPath + plen] = (u8)slot_byte;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < kPathRegs; ++i)
    if (plen == (u32)i) pr.blk[i] = blk;
  if (plen < 4u) pr.slots_lo |= slot_byte << (8u * plen);
  else if (plen < 8u) pr.slots_hi |= slot_byte << (8u * (plen - 4u));
}
AZ_HD void path_put(const EngineView& E, u32 g, PathSm& pr, u32 plen, u32 blk, u32 slot_byte) {
  if (plen >= (u32)kPathSm) {
    E.path[(size_t)g * kMaxPath + plen] = blk;
    E.pslot[(size_t)g * kMaxPath + plen] = (u8)slot_byte;
  } else {
    pr.blk[plen] = blk;
    pr.slot[plen] = (u8)slot_byte;
  }
}
// edge i of the pending leaf's path; `cached` = the on-chip copy is valid (else everything is in HBM)
AZ_HD void path_get(const EngineView& E, u32 g, const PathRegs& pr, bool cached, u32 i, u32& b, u32& sb) {
#if defined(B2AZ_NO_PATHREGS)
  cached = false;
#endif
  if (cached && i < (u32)kPathRegs) {
    u32 x = pr.blk[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 1; r < kPathRegs; ++r)
      if (i == (u32)r) x = pr.blk[r];
    b = x;
    sb = ((i < 4u ? pr.slots_lo : pr.slots_hi) >> (8u * (i & 3u))) & 0xFFu;
  } else {
    b = E.path[(size_t)g * kMaxPath + i];
    sb = E.pslot[(size_t)g * kMaxPath + i];
  }
}
AZ_HD void path_get(const EngineView& E, u32 g, const PathSm& pr, bool cached, u32 i, u32& b, u32& sb) {
  if (cached && i < (u32)kPathSm) {
    b = pr.blk[i];
    sb = pr.slot[i];
  } else {
    b = E.path[(size_t)g * kMaxPath + i];
    sb = E.pslot[(size_t)g * kMaxPath + i];
  }
}
// the pending leaf's path goes to HBM for a later launch (or for the move code)
AZ_HD void path_flush(const EngineView& E, u32 g, const PathRegs& pr, u32 path_len) {
  if (!pr.valid) return;
  u32* path = E.path + (size_t)g * kMaxPath;
  u8* pslot = E.pslot + (size_t)g * kMaxPath;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < kPathRegs; ++i)
    if ((u32)i < path_len) {
      path[i] = pr.blk[i];
      pslot[i] = (u8)(((i < 4 ? pr.slots_lo : pr.slots_hi) >> (8 * (i & 3))) & 0xFFu);
    }
}
AZ_HD void path_flush(const EngineView& E, u32 g, const PathSm& pr, u32 path_len) {
  if (!pr.valid) return;
  u32* path = E.path + (size_t)g * kMaxPath;
  u8* pslot = E.pslot + (size_t)g * kMaxPath;
  for (u32 i = 0; i < (u32)kPathSm && i < path_len; ++i) {
    path[i] = pr.blk[i];
    pslot[i] = pr.slot[i];
  }
}
// The thread-per-game kernels keep the first kPathSm edges in shared memory too, one COLUMN per thread (entry i of
// thread t at [i * kPathColStride + t]: the 32 lanes of a warp hit 32 different banks).
constexpr int kPathColStride = 64;
struct PathCol {
  u32* blk;   // &s_blk[0][thread]
  u8* slot;   // &s_slot[0][thread]
  u32 valid;
};
AZ_HD void path_begin(PathCol& pr) { pr.valid = 1; }
AZ_HD void path_put(const EngineView& E, u32 g, PathCol& pr, u32 plen, u32 blk, u32 slot_byte) {
  if (plen >= (u32)kPathSm) {
    E.path[(size_t)g * kMaxPath + plen] = blk;
    E.pslot[(size_t)g * kMaxPath + plen] = (u8)slot_byte;
  } else {
    pr.blk[plen * (u32)kPathColStride] = blk;
    pr.slot[plen * (u32)kPathColStride] = (u8)slot_byte;
  }
}
AZ_HD void path_get(const EngineView& E, u32 g, const PathCol& pr, bool cached, u32 i, u32& b, u32& sb) {
  if (cached && i < (u32)kPathSm) {
    b = pr.blk[i * (u32)kPathColStride];
    sb = pr.slot[i * (u32)kPathColStride];
  } else {
    b = E.path[(size_t)g * kMaxPath + i];
    sb = E.pslot[(size_t)g * kMaxPath + i];
  }
}
AZ_HD void path_flush(const EngineView& E, u32 g, const PathCol& pr, u32 path_len) {
  if (!pr.valid) return;
  u32* path = E.path + (size_t)g * kMaxPath;
  u8* pslot = E.pslot + (size_t)g * kMaxPath;
  for (u32 i = 0; i < (u32)kPathSm && i < path_len; ++i) {
    path[i] = pr.blk[i * (u32)kPathColStride];
    pslot[i] = pr.slot[i * (u32)kPathColStride];
  }
}
// ------------------------------------------------------------------------------------ Gumbel root search
// MCTS::set_gumbel_num_sims / reset / init / advance_phase / next_root_child / interior_select /
// improved_policy / final_action (mcts.cc:28-89, 175-401). All cold: they run once per simulation at the
// root (one 80 B state record per tree) or once per move.
#define AZ_GUMBEL_LOG_FLOOR 1e-20f
AZ_HD void gumbel_reset(GumbelState& S) {  // reset_gumbel_state (mcts.cc:180-188)
  S.initialized = 0;
  S.effective_m = 0;
  S.n_surv = 0;
  S.survivors = 0;
  S.n_phases = 0;
  S.phase_idx = 0;
  S.sims_in_phase = 0;
  for (int j = 0; j < kKMax; ++j) S.g[j] = 0.0f;
}
AZ_HD void gumbel_set_num_sims(GumbelState& S, u32 n) {  // mcts.cc:175-178
  S.num_sims_target = n;
  gumbel_reset(S);
}
// seq_halving_phase_plan (mcts.cc:28-66)
AZ_HD void gumbel_phase_plan(GumbelState& S, u32 m, u32 n) {
  S.n_phases = 0;
  if (m <= 1) {
    S.phase_numc[0] = 1; S.phase_vper[0] = n; S.n_phases = 1;
    return;
  }
  u32 log2m = 0;
  for (u32 v = m - 1; v > 0; v >>= 1) ++log2m;
  if (log2m == 0) log2m = 1;
  u32 base_v = n / (log2m * m);
  if (base_v < 1u) base_v = 1u;
  u32 sims_used = 0, num_c = m;
  for (u32 phase_idx = 0; phase_idx < log2m && S.n_phases < (u8)kMaxPhases; ++phase_idx) {
    if (sims_used >= n) break;
    const u32 remaining = n - sims_used;
    const bool is_final = (phase_idx == log2m - 1);
    u32 v_per = is_final ? (remaining / num_c < 1u ? 1u : remaining / num_c) : base_v * (1u << phase_idx);
    if (num_c * v_per > remaining) {
      v_per = remaining / num_c;
      if (v_per == 0) { num_c = remaining; v_per = 1; }
    }
    S.phase_numc[S.n_phases] = num_c;
    S.phase_vper[S.n_phases] = v_per;
    ++S.n_phases;
    sims_used += num_c * v_per;
    num_c = num_c / 2 < 1u ? 1u : num_c / 2;
  }
}
// Top-`take` of k scores, descending. std::partial_sort in the reference; with continuous Gumbel noise in
// every score ties have probability zero, so a plain selection gives the same ranking.
AZ_HD u32 rank_top(const float* score, const u32* idx, u32 k, u32 take) {
  u32 used = 0, out = 0;
  for (u32 r = 0; r < take; ++r) {
    int best = -1;
    for (u32 i = 0; i < k; ++i) {
      if ((used >> i) & 1u) continue;
      if (best < 0 || score[i] > score[best]) best = (int)i;
    }
    used |= 1u << best;
    out |= idx[best] << (4u * r);
  }
  return out;
}
// init_gumbel_state (mcts.cc:190-227)
AZ_COLD void gumbel_init(const EngineView E, GumbelState& S, const TreeHdr& T, Pcg32& rng) {
  const u32 num_legal = T.k;
  if (num_legal == 0 || T.fc == kNil) return;
  const u32 remaining = T.depth < S.num_sims_target ? S.num_sims_target - T.depth : 0u;
  if (remaining == 0) return;
  u32 m = E.gumbel_m < num_legal ? E.gumbel_m : num_legal;
  if (remaining < m) m = remaining;
  if (m < 1u) m = 1u;
  S.effective_m = (u8)m;
  const Block* B = E.blocks + T.fc;
  float score[kKMax];
  u32 idx[kKMax];
  for (u32 i = 0; i < num_legal; ++i) S.g[i] = rng_gumbel(rng);
  for (u32 i = 0; i < num_legal; ++i) {
    score[i] = fadd(S.g[i], az_logf(fadd(blk_pol(B, i), AZ_GUMBEL_LOG_FLOOR)));
    idx[i] = i;
  }
  S.survivors = rank_top(score, idx, num_legal, m);
  S.n_surv = (u8)m;
  gumbel_phase_plan(S, m, remaining);
  S.phase_idx = 0;
  S.sims_in_phase = 0;
  S.initialized = 1;
}
AZ_HD float gumbel_sigma_scale(const EngineView& E, u32 max_visit) {
  return fmul(fadd(E.gumbel_c_visit, (float)max_visit), E.gumbel_c_scale);
}
// gumbel_advance_phase (mcts.cc:229-264)
AZ_HD void gumbel_advance_phase(const EngineView& E, GumbelState& S, const TreeHdr& T) {
  if ((u32)S.phase_idx + 1u >= (u32)S.n_phases) return;
  const u32 next_num_c = S.phase_numc[S.phase_idx + 1];
  if (next_num_c >= (u32)S.n_surv) {
    ++S.phase_idx;
    S.sims_in_phase = 0;
    return;
  }
  const Block* B = E.blocks + T.fc;
  u32 max_visit = 0;
  for (u32 r = 0; r < (u32)S.n_surv; ++r) {
    const u32 c = (S.survivors >> (4u * r)) & 15u;
    if (blk_n(B, c) > max_visit) max_visit = blk_n(B, c);
  }
  const float sigma_scale = gumbel_sigma_scale(E, max_visit);
  float score[kKMax];
  u32 idx[kKMax];
  for (u32 r = 0; r < (u32)S.n_surv; ++r) {
    const u32 c = (S.survivors >> (4u * r)) & 15u;
    const float logit = az_logf(fadd(blk_pol(B, c), AZ_GUMBEL_LOG_FLOOR));
    const float q_hat = blk_n(B, c) > 0 ? blk_q(B, c) : 0.0f;
    score[r] = fadd(fadd(S.g[c], logit), fmul(sigma_scale, q_hat));
    idx[r] = c;
  }
  S.survivors = rank_top(score, idx, S.n_surv, next_num_c);
  S.n_surv = (u8)next_num_c;
  ++S.phase_idx;
  S.sims_in_phase = 0;
}
// gumbel_next_root_child (mcts.cc:266-283)
AZ_COLD u32 gumbel_next_root_child(const EngineView E, GumbelState& S, const TreeHdr& T) {
  if ((u32)S.phase_idx < (u32)S.n_phases) {
    if (S.sims_in_phase >= S.phase_numc[S.phase_idx] * S.phase_vper[S.phase_idx]) gumbel_advance_phase(E, S, T);
  }
  if (S.n_surv == 0) return 0;
  const u32 pick = S.sims_in_phase % (u32)S.n_surv;
  ++S.sims_in_phase;
  return (S.survivors >> (4u * pick)) & 15u;
}
// compute_v_mix_from_children (mcts.cc:71-89) + the softmax of logits + sigma * completedQ shared by
// gumbel_interior_select and gumbel_improved_policy. z_out[i] = exp(z_i - z_max); returns z_sum.
AZ_HD float gumbel_pi_prime(const EngineView& E, const Block* B, u32 k, float raw_v, float* z, u32* ns, u32* sum_visits_out) {
  u32 max_visit = 0, sum_n = 0;
  float sum_visits = 0.0f, sum_priors_visited = 0.0f, weighted_num = 0.0f;
  for (u32 i = 0; i < k; ++i) {
    ns[i] = blk_n(B, i);
    if (ns[i] > max_visit) max_visit = ns[i];
    sum_n += ns[i];
    sum_visits = fadd(sum_visits, (float)ns[i]);
    if (ns[i] > 0) {
      sum_priors_visited = fadd(sum_priors_visited, blk_pol(B, i));
      weighted_num = fadd(weighted_num, fmul(blk_pol(B, i), blk_q(B, i)));
    }
  }
  float v_mix = raw_v;
  if (sum_priors_visited > 0.0f) {
    const float weighted_q = fdiv(weighted_num, sum_priors_visited);
    v_mix = fdiv(fadd(raw_v, fmul(sum_visits, weighted_q)), fadd(sum_visits, 1.0f));
  }
  const float sigma_scale = gumbel_sigma_scale(E, max_visit);
  float z_max = -INFINITY;
  for (u32 i = 0; i < k; ++i) {
    const float completed_q = ns[i] > 0 ? blk_q(B, i) : v_mix;
    z[i] = fadd(az_logf(fadd(blk_pol(B, i), AZ_GUMBEL_LOG_FLOOR)), fmul(sigma_scale, completed_q));
    if (z[i] > z_max) z_max = z[i];
  }
  float z_sum = 0.0f;
  for (u32 i = 0; i < k; ++i) {
    z[i] = az_expf(fsub(z[i], z_max));
    z_sum = fadd(z_sum, z[i]);
  }
  *sum_visits_out = sum_n;
  return z_sum;
}
// gumbel_interior_select (mcts.cc:285-334)
AZ_COLD u32 gumbel_interior_select(const EngineView E, u32 blk, u32 k, float node_v) {
  const Block* B = E.blocks + blk;
  float z[kKMax];
  u32 ns[kKMax], sum_visits = 0;
  const float z_sum = gumbel_pi_prime(E, B, k, node_v, z, ns, &sum_visits);
  const float inv = z_sum > 0.0f ? fdiv(1.0f, z_sum) : 0.0f;
  const float denom = fadd(1.0f, (float)sum_visits);
  u32 best = 0;
  float best_score = -INFINITY;
  for (u32 i = 0; i < k; ++i) {
    const float score = fsub(fmul(z[i], inv), fdiv((float)ns[i], denom));
    if (score > best_score) { best_score = score; best = i; }
  }
  return best;
}
// gumbel_improved_policy (mcts.cc:336-373): pi' over all moves (zeros for illegal ones)
AZ_COLD void gumbel_improved_policy(const EngineView E, const TreeHdr& T, float* out) {
  for (int m = 0; m < kA; ++m) out[m] = 0.0f;
  const u32 k = T.k;
  if (k == 0 || T.fc == kNil) return;
  const Block* B = E.blocks + T.fc;
  float z[kKMax];
  u32 ns[kKMax], sum_visits = 0;
  const float z_sum = gumbel_pi_prime(E, B, k, T.v, z, ns, &sum_visits);
  if (z_sum <= 0.0f) return;
  for (u32 i = 0; i < k; ++i) out[blk_mv(B, i)] = fdiv(z[i], z_sum);
}

// ------------------------------------------------------------------------------------ find_leaf
// MCTS::find_leaf (mcts.cc:462-498), PUCT branch; Node::best_child (mcts.cc:130-149) and Node::uct
// (mcts.cc:123-128) inlined. n_in_flight is always 0 on this path (the WU-UCT variant is not used by
// PlayManager). Split in three pieces so that the fused kernel can run the descent ONE
// LEVEL PER LOOP ITERATION for every game of a warp (run_flat below): begin / level / finish.
struct __attribute__((aligned(16))) Descent {
  C4State s;          // the position replayed along the path
  u32 blk;            // child block of the current node
  u32 cur_n, cur_term, cur_player, cur_k;
  float cur_v;
  u32 par_blk, par_slot;
  u32 plen;
  bool at_root;
  bool gumbel;        // Gumbel selection is active for this descent (MCTS::gumbel_initialized_)
};
static_assert(sizeof(Descent) == 64, "Descent must stay four 16 B vectors (shared-memory record of the queue kernel)");
// Lazy Gumbel init (mcts.cc:468-472): after the root has been expanded, when a sims target is set.
AZ_COLD bool gumbel_begin(const EngineView E, u32 tree, const TreeHdr T, Pcg32& rng) {
  GumbelState S = E.gum[tree];
  if (!S.initialized && S.num_sims_target > 0 && T.n > 0 && T.k > 0) {
    gumbel_init(E, S, T, rng);
    E.gum[tree] = S;
  }
  return S.initialized != 0;
}
// GB = false compiles the Gumbel hooks out of the descent (the step kernel picks the instantiation from
// params.gumbel_enabled): no out-of-line call sites, hence no caller-saved registers, inside the hot loop
// (spill stores 178 B -> 32 B; throughput unchanged).
template <bool GB = true, class PR = PathRegs>
AZ_HD void descent_begin(const EngineView& E, u32 g, const TreeHdr& T, const GameSlot& gs, Descent& D, PR& pr,
                         Pcg32& rng) {
  D.gumbel = false;
  if (GB && E.gumbel_enabled) {
    Pcg32 r = rng;  // a copy keeps the address handed to the out-of-line code away from the hot state
    D.gumbel = gumbel_begin(E, g * (u32)kP + gs.player, T, r);
    rng = r;
  }
  D.s.p[0] = gs.p0; D.s.p[1] = gs.p1; D.s.turn = gs.turn; D.s.player = gs.player;
  D.blk = T.fc;
  D.cur_n = T.n; D.cur_term = T.term; D.cur_player = T.player; D.cur_k = T.k;
  D.cur_v = T.v;
  D.at_root = true;
  D.par_blk = kNil; D.par_slot = 0;
  D.plen = 0;
  path_begin(pr);
}
AZ_HD bool descent_more(const Descent& D) { return D.cur_n > 0 && D.cur_term == 0; }
// One level: load the node's child block, pick best_child, replay the move. Returns false on the
// (impossible for Connect4) structural error, which ends the descent.
template <bool GB = true, class PR = PathRegs>
AZ_HD bool descent_level(const EngineView& E, u32 g, Descent& D, PR& pr) {
  if (D.blk == kNil || D.plen >= (u32)kMaxPath) {  // cannot happen for a non-terminal Connect4 node
    at_or(&E.glob->error, B2AZ_DEVERR_DEPTH);
    return false;
  }
  const Block* B = E.blocks + D.blk;
  // one burst: everything this level (and its backprop) needs — ten independent 16 B loads
  V4 r[kKMax], f0, f1, hd;
#if B2AZ_L2_KEEP_LEVELS > 0
  if (D.plen < (u32)B2AZ_L2_KEEP_LEVELS) {
    const u64 keep = l2_policy_keep();
#pragma unroll
    for (int j = 0; j < kKMax; ++j) r[j] = ld_v4_hint(B->rec[j], keep);
    f0 = ld_v4_hint(&B->fc[0], keep); f1 = ld_v4_hint(&B->fc[4], keep); hd = ld_v4_hint(&B->mv, keep);
  } else
#endif
  {
#pragma unroll
    for (int j = 0; j < kKMax; ++j) r[j] = ld_v4(B->rec[j]);
    f0 = ld_v4(&B->fc[0]); f1 = ld_v4(&B->fc[4]); hd = ld_v4(&B->mv);
  }
  if (!D.at_root) {
    D.cur_v = u2f(hd.z);
    D.cur_k = hd.w & 0xFFu;
    D.cur_player = (hd.w >> 8) & 0xFFu;
  }
  u32 forced = kNil;  // Gumbel: the root child comes from the halving schedule, interior nodes (gumbel_full) from pi'
  if (GB && D.gumbel) {
    if (D.at_root) {
      const u32 tree = g * (u32)kP + D.cur_player;  // the searching seat's tree: the root's side to move
      GumbelState S = E.gum[tree];
      TreeHdr R;  // the fields gumbel_next_root_child reads
      R.fc = D.blk; R.k = (u8)D.cur_k;
      forced = gumbel_next_root_child(E, S, R);
      E.gum[tree] = S;
    } else if (E.gumbel_full) {
      forced = gumbel_interior_select(E, D.blk, D.cur_k, D.cur_v);
    }
  }
  const float fpu = (D.at_root && E.root_fpu_zero) ? 0.0f : E.fpu_reduction;
  float seen = 0.0f;
#pragma unroll
  for (int j = 0; j < kKMax; ++j)
    if ((u32)j < D.cur_k && r[j].x > 0) seen = fadd(seen, u2f(r[j].z));
  const float fpu_value = fsub(D.cur_v, fmul(fpu, fsqrt(seen)));
  const float sqrt_n = fsqrt((float)D.cur_n);
  u32 best = 0, best_n = 0, best_fc = kNil;
  float best_u = 0.0f;
#pragma unroll
  for (int j = 0; j < kKMax; ++j) {
    if (j > 0 && (u32)j >= D.cur_k) continue;  // pad slots: 0 / 1 would take the division's slow path
    const u32 nj = r[j].x;
    const float base = (nj == 0) ? fpu_value : u2f(r[j].y);
    const float u = fadd(base, fdiv_pos(fmul(fmul(E.cpuct, u2f(r[j].z)), sqrt_n), (float)(nj + 1u)));
    if ((!GB || forced == kNil) ? (j == 0 || u > best_u) : ((u32)j == forced)) {
      best_u = u;
      best = (u32)j;
      best_n = nj;
      best_fc = j == 0 ? f0.x : j == 1 ? f0.y : j == 2 ? f0.z : j == 3 ? f0.w : j == 4 ? f1.x : j == 5 ? f1.y : f1.z;
    }
  }
  const u32 move = (hd.x >> (4u * best)) & 15u;
  const u32 cterm = (hd.y >> (2u * best)) & 3u;
  const u32 slot_byte = best | (D.cur_player << 4);
  const u32 plen = D.plen;
  path_put(E, g, pr, plen, D.blk, slot_byte);
  D.plen = plen + 1u;
  c4_play(D.s, move);
  D.par_blk = D.blk;
  D.par_slot = best;
  D.cur_n = best_n;
  D.cur_term = cterm;
  D.blk = best_fc;
  D.at_root = false;
  return true;
}
// The leaf: expansion of a new node (or a terminal node visited again), and the net's input row.
AZ_HD void descent_finish(const EngineView& E, u32 g, TreeHdr& T, GameSlot& gs, Pcg32& rng, const Descent& D) {
  const C4State& s = D.s;
  T.total_leaf_depth += D.plen;
  T.path_len = (u16)D.plen;
  if (D.cur_n == 0) {
    // expand: current_->player, scores, add_children(valid_moves) incl. the shuffle (mcts.cc:490-496, 93-101)
    const u32 term = c4_terminal(s);
    const u32 vm = c4_valid_mask(s);
    u32 moves = 0, k = 0;  // ascending legal moves as nibbles
#pragma unroll
    for (u32 w = 0; w < (u32)kA; ++w)
      if ((vm >> w) & 1u) moves |= w << (4u * k++);
    rng_shuffle_nib(rng, moves, k);
    // Children of a terminal node are never visited (selection stops at scores != nullptr,
    // mcts.cc:473) — their RNG draws are consumed above, their storage is skipped.
    const u32 kk = term ? 0u : k;
    u32 nb = kNil;
    if (kk > 0) {
      nb = tree_alloc_block(E, T);
      if (nb != kNil) {
        Block* N = E.blocks + nb;
        const V4 z = mk_v4(0u, 0u, 0u, 0u), f = mk_v4(kNil, kNil, kNil, kNil);
#pragma unroll
        for (int j = 0; j < kKMax; ++j) st_v4(N->rec[j], z);
        st_v4(&N->fc[0], f);
        st_v4(&N->fc[4], f);
        // moves (already nibble-packed in child order), no terminal codes yet, v by the first backprop
        st_v4(&N->mv, mk_v4(moves, 0u, 0u, kk | ((u32)s.player << 8)));
      }
    }
    const u32 k_eff = (nb == kNil) ? 0u : kk;
    if (D.at_root) {
      T.player = s.player; T.term = (u8)term; T.fc = nb; T.k = (u8)k_eff;
    } else {
      Block* Pb = E.blocks + D.par_blk;
      Pb->fc[D.par_slot] = nb;
      if (term) Pb->term |= term << (2u * D.par_slot);  // was 0 (unknown)
    }
    T.leaf_term = (u8)term;
    T.leaf_k = (u8)k_eff;
    T.leaf_player = s.player;
    T.leaf_blk = nb;
  } else {  // a terminal node visited again
    T.leaf_term = (u8)D.cur_term;
    T.leaf_k = 0;
    T.leaf_player = s.player;
    T.leaf_blk = kNil;
  }
}

// ------------------------------------------------------------------------------------ position cache
// Exact key of a (gravity-valid) Connect4 position: stones of player 0 + all stones is injective per column
// (the classic position + mask encoding), the side to move goes in bit 63, and 0 stays "empty slot".
AZ_HD u64 c4_cache_key(const C4State& s) { return ((s.p[0] + (s.p[0] | s.p[1]) + 1ULL) & 0x7FFFFFFFFFFFFFFFULL) | ((u64)s.player << 63); }
AZ_HD u64 mix64(u64 x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDULL;
  x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ULL;
  x ^= x >> 33;
  return x;
}
// S3FIFOCache::find (s3fifo_cache.h:41-60): returns the value index or kNil; counts hits / misses /
// reinserts (a miss whose key sits in the ghost set); a hit bumps the 2-bit frequency.
AZ_HD u32 cache_find(const EngineView& E, u64 key) {
  const u64 h = mix64(key);
  const u32 b = (u32)(h % (u64)E.cache_buckets);
  const u64* K = E.cache_keys + (size_t)b * kCacheWays;
  const V4 k01 = ld_v4(K), k23 = ld_v4(K + 2);
  const u64 k0 = (u64)k01.x | ((u64)k01.y << 32), k1 = (u64)k01.z | ((u64)k01.w << 32);
  const u64 k2 = (u64)k23.x | ((u64)k23.y << 32), k3 = (u64)k23.z | ((u64)k23.w << 32);
  int way = -1;
  if (k0 == key) way = 0;
  else if (k1 == key) way = 1;
  else if (k2 == key) way = 2;
  else if (k3 == key) way = 3;
  if (way < 0) {
    at_add64(&E.glob->cache_misses, 1ULL);
    if (E.cache_ghost_slots) {
      const u32 fp = (u32)(h >> 32) | 1u;
      if (E.cache_ghost[(u32)((h >> 20) % (u64)E.cache_ghost_slots)] == fp) at_add64(&E.glob->cache_reinserts, 1ULL);
    }
    return kNil;
  }
  at_add64(&E.glob->cache_hits, 1ULL);
  const u32 m = E.cache_meta[b];
  if (((m >> (8 * way)) & 3u) < 3u) at_add(&E.cache_meta[b], 1u << (8 * way));  // freq saturates at 3 (:53)
  return b * (u32)kCacheWays + (u32)way;
}
// S3FIFOCache::insert (s3fifo_cache.h:62-110) for one key, under the bucket's lock. Existing key: no-op.
// New keys enter the "small" generation (main flag 0) unless the ghost set remembers them; the victim of a
// full bucket is chosen S3-FIFO style: small entries that were never hit again go first, hit small entries
// are promoted to main, main entries get second chances while their frequency lasts.
AZ_HD void cache_insert(const EngineView& E, u64 key, const float* v, const float* pi) {
  const u64 h = mix64(key);
  const u32 b = (u32)(h % (u64)E.cache_buckets);
  u64* K = E.cache_keys + (size_t)b * kCacheWays;
#if defined(__CUDA_ARCH__)
  bool locked = false;
  for (u32 spin = 0; spin < 4096u; ++spin)
    if (at_cas(&E.cache_lock[b], 0u, 1u) == 0u) { locked = true; break; }
  if (!locked) return;  // best effort: a cache may always forget
  mem_fence();
#endif
  u64 k[kCacheWays];
  for (int w = 0; w < kCacheWays; ++w) k[w] = ld_volatile(&K[w]);
  int way = -1;
  bool present = false;
  for (int w = 0; w < kCacheWays; ++w) {
    if (k[w] == key) present = true;
    if (k[w] == 0 && way < 0) way = w;
  }
  if (!present) {
    u32 m = ld_volatile(&E.cache_meta[b]);
    const u32 fp = (u32)(h >> 32) | 1u;
    u32* ghost = E.cache_ghost_slots ? &E.cache_ghost[(u32)((h >> 20) % (u64)E.cache_ghost_slots)] : nullptr;
    bool to_main = false;
    if (ghost && ld_volatile(ghost) == fp) {  // ghost hit: admitted to Main (:86-104)
      to_main = true;
      st_volatile(ghost, 0u);
    }
    if (way < 0) {  // evict_one (:118-149)
      for (int round = 0; round < 6 && way < 0; ++round) {
        for (int w = 0; w < kCacheWays && way < 0; ++w) {  // the small generation first
          const u32 mw = (m >> (8 * w)) & 0xFFu;
          if (mw & 4u) continue;
          if (mw & 3u) m = (m & ~(0xFFu << (8 * w))) | (4u << (8 * w));  // hit while small: promote, freq 0
          else way = w;
        }
        for (int w = 0; w < kCacheWays && way < 0; ++w) {  // then Main with second chances
          const u32 mw = (m >> (8 * w)) & 0xFFu;
          if (!(mw & 4u)) continue;
          if (mw & 3u) m -= 1u << (8 * w);
          else way = w;
        }
      }
      if (way < 0) way = (int)(h >> 60) & 3;
      if (!(((m >> (8 * way)) & 4u)) && ghost) {  // evicted from Small: remember it in the ghost set
        const u64 hv = mix64(k[way]);
        E.cache_ghost[(u32)((hv >> 20) % (u64)E.cache_ghost_slots)] = (u32)(hv >> 32) | 1u;
      }
      at_add64(&E.glob->cache_evictions, 1ULL);
    } else {
      at_add64(&E.glob->cache_size, 1ULL);
    }
    CacheVal* V = E.cache_vals + (size_t)b * kCacheWays + way;
    for (int i = 0; i < kA; ++i) V->pi[i] = pi[i];
    for (int i = 0; i < kP + 1; ++i) V->v[i] = v[i];
    m = (m & ~(0xFFu << (8 * way))) | ((to_main ? 4u : 0u) << (8 * way));
    st_volatile(&E.cache_meta[b], m);
    st_volatile(&K[way], key);
  }
#if defined(__CUDA_ARCH__)
  mem_fence();
  at_exch(&E.cache_lock[b], 0u);
#endif
}

// eval_types_[group] == RANDOM next to an NN group (play_manager.cc:577-587): the searches of that group run dumb_eval
// inline and never enter the leaf batch
AZ_HD bool slot_random(const EngineView& E, u32 g, const GameSlot& gs) {
  return E.perms && E.perms->random_groups != 0u && ((E.perms->random_groups >> seat_group_of(E, g, gs.player)) & 1u) != 0u;
}
// PX = false: the instantiation for runs with one seat permutation and no RANDOM group next to an NN one (self-play, the
// bench): the per-slot lookups compile away 
template <bool PX>
AZ_HD bool slot_random_t(const EngineView& E, u32 g, const GameSlot& gs) { return PX ? slot_random(E, g, gs) : false; }
// What happens to a fresh leaf with the NN evaluator (play_manager.cc:586-598): look the position up in the
// cache — every leaf, terminal ones included, like the reference — and on a miss put it into the leaf batch.
// Returns true on a cache hit (the caller goes on with the next simulation right away).
template <bool PX = true>
AZ_HD bool leaf_emit(const EngineView& E, u32 g, GameSlot& gs, const C4State& s, bool allow_hit) {
  if (slot_random_t<PX>(E, g, gs)) return allow_hit;  // answered on the spot like a cache hit; a launch's chain stays bounded (hit_cap)
  u64 key = 0;
  if (E.cache_buckets) {
    // one table for every model group (the reference keeps one cache per group, play_manager.cc:195-203): the group of
    // the searching seat is part of the key (bits 56-59 are free in the position encoding)
    key = c4_cache_key(s) | ((u64)seat_group_of<PX>(E, g, gs.player) << 56);
    const u32 hit = cache_find(E, key);
    if (hit != kNil && allow_hit) {
      E.hit_val[g] = hit;
      return true;
    }
  }
  u32 row;
#if defined(__CUDA_ARCH__)
  // leaf-batch compaction: one atomicAdd per warp, rows handed out by ballot rank
  const unsigned active = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(active) - 1;
  u32 base = 0;
  if ((int)lane == leader) base = atomicAdd(&E.glob->leaf_count, (u32)__popc(active));
  base = __shfl_sync(active, base, leader);
  row = base + (u32)__popc(active & ((1u << lane) - 1u));
#else
  row = at_add(&E.glob->leaf_count, 1u);
#endif
  E.leaf_p0[row] = s.p[0];
  E.leaf_p1[row] = s.p[1];
  E.leaf_player[row] = s.player;
  E.leaf_game[row] = g;
  E.leaf_seat[row] = (u8)(gs.player | (seat_group_of<PX>(E, g, gs.player) << 4));
  if (E.cache_buckets) {
    E.leaf_key[row] = key;
    E.hit_val[g] = kNil;
  }
  gs.eval_row = row;
  return false;
}
AZ_HD void find_leaf(const EngineView& E, u32 g, TreeHdr& T, GameSlot& gs, Pcg32& rng, PathRegs& pr) {
  Descent D;
  descent_begin(E, g, T, gs, D, pr, rng);
  while (descent_more(D))
    if (!descent_level(E, g, D, pr)) break;
  descent_finish(E, g, T, gs, rng, D);
  if (E.eval_type == 0) leaf_emit(E, g, gs, D.s, /*allow_hit=*/false);
}

// ------------------------------------------------------------------------------------ root priors (cold)
// MCTS::add_root_noise (mcts.cc:403-446). `pol` = root child priors in child order.
AZ_COLD void add_root_noise(const EngineView& E, Pcg32& rng, float* pol, u32 k) {
  float noise[kKMax];
  double sum = 0.0;
  if (E.shaped_dirichlet && k > 1) {
    const float N = (float)k;
    float log_sum = 0.0f;
    float lp[kKMax];
    for (u32 j = 0; j < k; ++j) {
      lp[j] = az_logf(fadd(std_min(pol[j], 0.01f), 1e-20f));
      log_sum = fadd(log_sum, lp[j]);
    }
    const float log_mean = fdiv(log_sum, N);
    float shaped_sum = 0.0f;
    for (u32 j = 0; j < k; ++j) shaped_sum = fadd(shaped_sum, std_max(0.0f, fsub(lp[j], log_mean)));
    const float uniform = fdiv(1.0f, N);
    for (u32 j = 0; j < k; ++j) {
      const float shaped = std_max(0.0f, fsub(lp[j], log_mean));
      float alpha_prop = (shaped_sum > 0.0f) ? fmul(0.5f, fadd(fdiv(shaped, shaped_sum), uniform)) : uniform;
      alpha_prop = std_max(alpha_prop, 1e-6f);
      GammaDist gd;
      gamma_init(gd, fmul(10.83f, alpha_prop));
      noise[j] = gamma_draw(rng, gd);
      sum = dadd(sum, (double)noise[j]);
    }
  } else {
    GammaDist gd;
    gamma_init(gd, fdiv(10.83f, (float)k));
    for (u32 j = 0; j < k; ++j) {
      noise[j] = gamma_draw(rng, gd);
      sum = dadd(sum, (double)noise[j]);
    }
  }
  const float fsum = (float)sum;
  const float keep = fsub(1.0f, E.epsilon);  // (1 - epsilon_): int 1 -> float
  for (u32 j = 0; j < k; ++j) pol[j] = fadd(fmul(pol[j], keep), fdiv(fmul(E.epsilon, noise[j]), fsum));
}

// set_policy_normalized at the root (mcts.cc:109-121 with the temperature branch) + add_root_noise
// (mcts.cc:511-519): the priors of a root that has just been evaluated.
AZ_COLD void root_leaf_priors(const EngineView E, Pcg32& rng, float* p8, u32 lk, bool noise_enabled) {
  const bool apply_temp = (E.root_temp != 1.0f);
  const float inv_temp = fdiv(1.0f, E.root_temp);
  float sum = 0.0f;
  for (u32 j = 0; j < lk; ++j) {
    if (apply_temp) p8[j] = az_powf(p8[j], inv_temp);
    sum = fadd(sum, p8[j]);
  }
  for (u32 j = 0; j < lk; ++j) p8[j] = fdiv(p8[j], sum);
  if (noise_enabled && lk > 0) add_root_noise(E, rng, p8, lk);
}

// ------------------------------------------------------------------------------------ process_result
// MCTS::process_result (mcts.cc:500-555).
template <class PR, bool PX = true>
AZ_HD void process_result(const EngineView& E, u32 g, TreeHdr& T, const GameSlot& gs, Pcg32& rng, bool noise_enabled,
                          PR& pr) {
  float val0, val1, vald;  // value[0], value[1], value[P] (draw share)
  const u32 lterm = T.leaf_term, lk = T.leaf_k, lplayer = T.leaf_player, lblk = T.leaf_blk;
  const u32 plen = T.path_len;
  if (lterm != 0) {
    val0 = (lterm == 1) ? 1.0f : 0.0f;
    val1 = (lterm == 2) ? 1.0f : 0.0f;
    vald = (lterm == 3) ? 1.0f : 0.0f;
  } else {
    u32 ps[kKMax];  // the leaf's child priors (f32 bits), child order
#pragma unroll
    for (int j = 0; j < kKMax; ++j) ps[j] = 0u;
    if (E.eval_type == 1 || slot_random_t<PX>(E, g, gs)) {  // dumb_eval (game_state.h:160-173): uniform over the legal moves, value 1/3
      const float third = (float)(1.0 / 3.0);
      val0 = val1 = vald = third;
      // every legal move is a child here, so Vector<uint8_t>::sum() == lk
      const float p = fdiv(1.0f, (float)lk);
#pragma unroll
      for (int j = 0; j < kKMax; ++j)
        if ((u32)j < lk) ps[j] = f2u(p);
    } else {
      const float* vrow = E.ev_v + (size_t)gs.eval_row * (kP + 1);
      const float* prow = E.ev_pi + (size_t)gs.eval_row * kA;
      if (E.cache_buckets) {  // a cache hit answers the leaf instead of the evaluation batch
        const u32 hit = E.hit_val[g];
        if (hit != kNil) {
          vrow = E.cache_vals[hit].v;
          prow = E.cache_vals[hit].pi;
        }
      }
      val0 = vrow[0]; val1 = vrow[1]; vald = vrow[2];
      if (lk > 0) {
        const u32 mvs = (E.blocks + lblk)->mv;
#pragma unroll
        for (int j = 0; j < kKMax; ++j)
          if ((u32)j < lk) ps[j] = f2u(prow[(mvs >> (4 * j)) & 15u]);
      }
    }
    if (plen == 0) {  // the leaf is the root: temperature + Dirichlet noise (cold)
      // copies keep the addresses handed to the out-of-line code away from the register-resident state
      float p8[kKMax];
      Pcg32 r = rng;
#pragma unroll
      for (int j = 0; j < kKMax; ++j) p8[j] = u2f(ps[j]);
      root_leaf_priors(E, r, p8, lk, noise_enabled && !E.gumbel_enabled);  // Gumbel replaces Dirichlet (mcts.cc:514-518)
      rng = r;
#pragma unroll
      for (int j = 0; j < kKMax; ++j) ps[j] = ((u32)j < lk) ? f2u(p8[j]) : 0u;
    } else {  // set_policy_normalized (mcts.cc:109-121), interior node: no temperature
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < kKMax; ++j)
        if ((u32)j < lk) sum = fadd(sum, u2f(ps[j]));
#pragma unroll
      for (int j = 0; j < kKMax; ++j)
        if ((u32)j < lk) ps[j] = f2u(fdiv(u2f(ps[j]), sum));
    }
    if (lk > 0) {  // the leaf was never visited: its records are {n 0, q 0, pol, d 0}
      Block* L = E.blocks + lblk;
#pragma unroll
      for (int j = 0; j < kKMax; ++j)
        if ((u32)j < lk) st_v4(L->rec[j], mk_v4(0u, 0u, ps[j], 0u));
    }
  }
  // backprop (mcts.cc:527-545): level i updates the child slot selected at level i. The levels touch
  // different blocks, so their loads are issued together (4 levels at a time) instead of one dependent
  // round trip per level.
  const float dshare = fdiv(vald, (float)kP);
  const bool cached = pr.valid != 0;
  pr.valid = 0;
  for (u32 base = 0; base < plen; base += 4u) {
    u32 bi[4], sb[4];
    V4 rc[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const u32 i = base + (u32)t;
      bi[t] = kNil;
      sb[t] = 0;
      if (i < plen) path_get(E, g, pr, cached, i, bi[t], sb[t]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      rc[t] = mk_v4(0u, 0u, 0u, 0u);
      if (bi[t] != kNil) rc[t] = ld_v4((E.blocks + bi[t])->rec[sb[t] & 15u]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (bi[t] == kNil) continue;
      Block* B = E.blocks + bi[t];
      const u32 sl = sb[t] & 15u, pp = sb[t] >> 4;
      const float v = fadd(pp == 0 ? val0 : val1, dshare);
      const u32 n0 = rc[t].x;
      const float qc = n0 ? u2f(rc[t].y) : 0.0f;
      const float dc = n0 ? u2f(rc[t].w) : 0.0f;
      const float qn = fdiv_pos(fadd(fmul(qc, (float)n0), v), (float)(n0 + 1u));
      const float dn = fdiv_pos(fadd(fmul(dc, (float)n0), vald), (float)(n0 + 1u));
      st_v4(B->rec[sl], mk_v4(n0 + 1u, f2u(qn), rc[t].z, f2u(dn)));
      // first visit of the node (only the leaf can be new): node.v from its own seat (mcts.cc:538-542)
      if (n0 == 0 && lblk != kNil) (E.blocks + lblk)->v = f2u(fadd(lplayer == 0 ? val0 : val1, dshare));
    }
  }
  if (T.n == 0) {
    T.v = fadd(lplayer == 0 ? val0 : val1, dshare);  // root_.player == the leaf's player when the root is the leaf
    T.d = vald;
  }
  ++T.depth;
  ++T.n;
  T.path_len = 0;
}

// ------------------------------------------------------------------------------------ per-move (cold)
// counts()/probs() family works on dense arrays in MOVE order (mcts.cc:557-618).
struct RootView {
  u32 k;
  u32 mv[kKMax];
  u32 n[kKMax];
  float pol[kKMax];
  float q[kKMax];
  float d[kKMax];
};
AZ_HD void root_view(const EngineView& E, const TreeHdr& T, RootView& R) {
  R.k = T.k;
  for (int j = 0; j < kKMax; ++j) { R.mv[j] = 0; R.n[j] = 0; R.pol[j] = R.q[j] = R.d[j] = 0.0f; }
  if (T.k > 0 && T.fc != kNil) {
    const Block* B = E.blocks + T.fc;
    for (u32 j = 0; j < (u32)kKMax; ++j) {
      R.mv[j] = blk_mv(B, j); R.n[j] = blk_n(B, j); R.pol[j] = blk_pol(B, j); R.q[j] = blk_q(B, j); R.d[j] = blk_d(B, j);
    }
  }
}
AZ_HD float sum7(const float* a) {  // Vector::sum(): sequential, starting from 0
  float s = 0.0f;
  for (int m = 0; m < kA; ++m) s = fadd(s, a[m]);
  return s;
}
// MCTS::probs(temp) (mcts.cc:575-618)
AZ_COLD void mcts_probs(const RootView& R, float temp, float* probs) {
  u32 counts[kA];
  for (int m = 0; m < kA; ++m) counts[m] = 0;
  for (u32 j = 0; j < R.k; ++j) counts[R.mv[j]] = R.n[j];
  float count_sum = 0.0f;
  for (int m = 0; m < kA; ++m) count_sum = fadd(count_sum, (float)counts[m]);
  if (count_sum == 0.0f) {
    for (int m = 0; m < kA; ++m) probs[m] = 0.0f;
    for (u32 j = 0; j < R.k; ++j) probs[R.mv[j]] = R.pol[j];
    if (temp != 0.0f) {
      const float e = fdiv(1.0f, temp);
      for (int m = 0; m < kA; ++m) probs[m] = az_powf(probs[m], e);
    }
    const float s = sum7(probs);
    for (int m = 0; m < kA; ++m) probs[m] = fdiv(probs[m], s);
    return;
  }
  if (temp == 0.0f) {
    u32 best = counts[0];
    int nbest = 1;
    for (int m = 1; m < kA; ++m) {
      if (counts[m] > best) { best = counts[m]; nbest = 1; }
      else if (counts[m] == best) ++nbest;
    }
    for (int m = 0; m < kA; ++m) probs[m] = (counts[m] == best) ? (float)(1.0 / (double)nbest) : 0.0f;
    return;
  }
  for (int m = 0; m < kA; ++m) probs[m] = (float)counts[m];
  float s = sum7(probs);
  for (int m = 0; m < kA; ++m) probs[m] = fdiv(probs[m], s);
  const float e = fdiv(1.0f, temp);  // `1 / temp`
  for (int m = 0; m < kA; ++m) probs[m] = az_powf(probs[m], e);
  s = sum7(probs);
  for (int m = 0; m < kA; ++m) probs[m] = fdiv(probs[m], s);
}
// MCTS::probs_pruned(temp) (mcts.cc:620-674)
AZ_COLD void mcts_probs_pruned(const EngineView& E, const TreeHdr& T, const RootView& R, float temp, float* out) {
  if (T.n <= 1) { mcts_probs(R, temp, out); return; }
  const float es = fmul(E.cpuct, fsqrt((float)T.n));
  float best_sel = -1e30f;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.n[j] == 0) continue;
    const float sel = fadd(R.q[j], fdiv(fmul(es, R.pol[j]), (float)(R.n[j] + 1u)));
    if (sel > best_sel) best_sel = sel;
  }
  float pruned[kA];
  for (int m = 0; m < kA; ++m) pruned[m] = 0.0f;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.n[j] == 0) continue;
    const float gap = fsub(best_sel, R.q[j]);
    float desired;
    if (gap <= 0.0f) desired = (float)R.n[j];
    else desired = fsub(fdiv(fmul(es, R.pol[j]), gap), 1.0f);
    pruned[R.mv[j]] = std_min((float)R.n[j], std_max(0.0f, desired));
  }
  const float total = sum7(pruned);
  if (total == 0.0f) { mcts_probs(R, temp, out); return; }
  if (temp == 0.0f) {
    float best = pruned[0];
    for (int m = 1; m < kA; ++m) best = std_max(best, pruned[m]);
    int cnt = 0;
    for (int m = 0; m < kA; ++m) if (pruned[m] == best) ++cnt;
    for (int m = 0; m < kA; ++m) out[m] = (pruned[m] == best) ? fdiv(1.0f, (float)cnt) : 0.0f;
    return;
  }
  for (int m = 0; m < kA; ++m) out[m] = fdiv(pruned[m], total);
  if (temp != 1.0f) {
    const float e = fdiv(1.0f, temp);
    for (int m = 0; m < kA; ++m) out[m] = az_powf(out[m], e);
    const float s = sum7(out);
    for (int m = 0; m < kA; ++m) out[m] = fdiv(out[m], s);
  }
}
// MCTS::pick_move (mcts.cc:717-735); returns kA if no move has positive probability
AZ_HD u32 mcts_pick_move(Pcg32& rng, const float* p) {
  const float choice = rng_uniform01(rng);
  float sum = 0.0f;
  for (u32 m = 0; m < (u32)kA; ++m) {
    sum = fadd(sum, p[m]);
    if (sum > choice) return m;
  }
  for (int m = kA - 1; m >= 0; --m)
    if (p[m] > 0.0f) return (u32)m;
  return (u32)kA;
}
// MCTS::normalized_root_entropy (mcts.cc:737-750)
AZ_COLD float mcts_root_entropy(const TreeHdr& T, const RootView& R) {
  const float k = (float)R.k;
  if (R.k <= 1 || T.n <= 1) return 0.0f;
  const float log_k = az_logf(k);
  float entropy = 0.0f;
  const float total_n = (float)T.n;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.n[j] > 0) {
      const float p = fdiv((float)R.n[j], total_n);
      entropy = fsub(entropy, fmul(p, az_logf(p)));
    }
  }
  return fdiv(entropy, log_k);
}
// MCTS::root_value (mcts.h:78-100) -> (w, l, d)
AZ_HD void mcts_root_value(const TreeHdr& T, const RootView& R, float* wld) {
  float q = 0.0f, d = 0.0f;
  bool found = false;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.n[j] > 0 && R.q[j] > q) {
      q = R.q[j];
      d = R.d[j];
      found = true;
    }
  }
  if (!found && T.n > 0) { q = T.v; d = T.d; }
  const float w = fsub(q, fdiv(d, (float)kP));
  wld[0] = w;
  wld[1] = (float)dsub(dsub(1.0, (double)w), (double)d);  // `1.0 - w - d` is evaluated in double
  wld[2] = d;
}

// Cheney copy of the live tree into fresh pages, breadth first (the device equivalent of the
// reference's `root_ = std::move(child)` + recursive ~Node of the discarded siblings, mcts.cc:163-172,
// done lazily: only when the arena has outgrown E.compact_pages).
AZ_COLD void tree_compact(const EngineView& E, TreeHdr& T) {
  if (T.fc == kNil) return;
  TreeHdr A;
  arena_clear(A);
  A.region = T.region;
  const u32 root_new = tree_alloc_block(E, A);
  if (root_new == kNil) return;
  blk_copy(E.blocks + root_new, E.blocks + T.fc);
  u32 scan_page = A.first_page, scan_off = 0;
  bool failed = false;
  // One scanned block per iteration, its seven child links handled by straight-line predicated code: in a batch of
  // 32 games (one per lane) every lane walks its own tree, and the first version's `for j ... continue` loop ran with
  // 1.8 of 32 lanes active (profiles/r2k: 3.1 ms of a 15.5 ms step went into this function).
  for (u32 guard = 0; guard < (1u << 24); ++guard) {
    if (scan_page == A.cur_page && scan_off >= A.bump) break;
    if (scan_off >= kPageBlocks) {
      scan_page = E.page_next[scan_page];
      scan_off = 0;
      continue;
    }
    Block* B = E.blocks + ((scan_page << kPageLog2) + scan_off);
    const V4 f0 = ld_v4(&B->fc[0]), f1 = ld_v4(&B->fc[4]);
    u32 fc0 = f0.x, fc1 = f0.y, fc2 = f0.z, fc3 = f0.w, fc4 = f1.x, fc5 = f1.y, fc6 = f1.z;
    bool any = false;
#define AZ_COMPACT_CHILD(fcj)                                   \
    if (fcj != kNil && !failed) {                               \
      const u32 dst = tree_alloc_block(E, A);                   \
      if (dst == kNil) failed = true;                           \
      else {                                                    \
        blk_copy(E.blocks + dst, E.blocks + fcj);               \
        fcj = dst;                                              \
        any = true;                                             \
      }                                                         \
    }
    AZ_COMPACT_CHILD(fc0) AZ_COMPACT_CHILD(fc1) AZ_COMPACT_CHILD(fc2) AZ_COMPACT_CHILD(fc3)
    AZ_COMPACT_CHILD(fc4) AZ_COMPACT_CHILD(fc5) AZ_COMPACT_CHILD(fc6)
#undef AZ_COMPACT_CHILD
    if (failed) break;
    if (any) {
      st_v4(&B->fc[0], mk_v4(fc0, fc1, fc2, fc3));
      st_v4(&B->fc[4], mk_v4(fc4, fc5, fc6, f1.w));
    }
    ++scan_off;
  }
  if (failed) {  // pool exhausted mid-copy (already flagged): keep the old, intact tree
    tree_free_pages(E, A);
    return;
  }
  if (T.first_page != kNil) pool_push_chain(E, T.first_page);
  T.first_page = A.first_page;
  T.cur_page = A.cur_page;
  T.bump = A.bump;
  T.pages_used = A.pages_used;
  T.fc = root_new;
  at_add64(&E.glob->compactions, 1ULL);
}

// MCTS::update_root (mcts.cc:151-173). `vm_before` = valid-move mask of the game state BEFORE the
// move (update_root receives the pre-move state, play_manager.cc:436-439). Re-rooting re-points the
// header at the chosen child's block; nothing is copied unless the arena is over budget.
AZ_COLD void update_root(const EngineView& E, TreeHdr& T, u32 move, u32 vm_before, Pcg32& rng) {
  T.depth = 0;
  T.total_leaf_depth = 0;
  T.path_len = 0;
  T.leaf_term = T.leaf_k = T.leaf_player = 0;
  T.leaf_blk = kNil;
  if (T.k == 0 || T.fc == kNil) {
    // root_.children.empty(): add_children(valid_moves()) shuffles children that are discarded by
    // the re-root onto the (unvisited) chosen child two lines later — only the RNG draws survive.
    rng_shuffle_discard(rng, (u32)popc32(vm_before));
    if (((vm_before >> move) & 1u) == 0u) at_or(&E.glob->error, B2AZ_DEVERR_MOVE);
    tree_free_pages(E, T);
    tree_reset(T);
    T.move = (u16)move;
    return;
  }
  const Block* B = E.blocks + T.fc;
  int ci = -1;
  for (u32 j = 0; j < T.k; ++j)
    if (blk_mv(B, j) == move) { ci = (int)j; break; }
  if (ci < 0) {
    at_or(&E.glob->error, B2AZ_DEVERR_MOVE);
    tree_free_pages(E, T);
    tree_reset(T);
    return;
  }
  // the chosen child becomes the root (Node tmp = std::move(*x); root_ = std::move(tmp))
  const u32 cn = blk_n(B, (u32)ci);
  const float cpol = blk_pol(B, (u32)ci);
  if (cn == 0) {  // never visited: a fresh root that only keeps its prior and move; nothing stays live
    tree_free_pages(E, T);
    tree_reset(T);
    T.policy = cpol;
    T.move = (u16)move;
    return;
  }
  const u32 cfc = B->fc[ci];
  T.q = blk_q(B, (u32)ci);
  T.d = blk_d(B, (u32)ci);
  T.policy = cpol;
  T.n = cn;
  T.move = (u16)move;
  T.term = (u8)blk_term(B, (u32)ci);
  if (cfc != kNil) {
    const Block* CB = E.blocks + cfc;
    T.v = blk_v(CB);
    T.k = (u8)blk_k(CB);
    T.player = (u8)blk_player(CB);
    T.fc = cfc;
    if ((u32)T.pages_used > E.compact_pages) tree_compact(E, T);
  } else {  // a terminal child: the game is over and both trees are reset right after (play_manager.cc:440-513)
    T.v = 0.0f;
    T.k = 0;
    T.player ^= 1;
    T.fc = kNil;
    tree_free_pages(E, T);
  }
}

// set_gumbel_num_sims for the tree that searches next (play_manager.cc:531-539, 562-570): the full budget, or for a
// capped search the cap when fast_search_uses_gumbel, else 0 = "PUCT for this search".
AZ_COLD void gumbel_arm(const EngineView E, u32 g, u32 seat, bool capped) {
  const u32 target = capped ? (E.fast_search_uses_gumbel ? seat_budget(E, g, seat, true) : 0u) : seat_budget(E, g, seat, false);
  GumbelState S = E.gum[(size_t)g * kP + seat];
  gumbel_set_num_sims(S, target);
  E.gum[(size_t)g * kP + seat] = S;
}

struct Ctx {       // what a thread keeps in registers across the fused steps of one launch
  GameSlot gs;
  TreeHdr T;       // the tree of the side to move (gs.player); the other seat's stays in HBM
  Pcg32 rng;
  PathRegs pr;
  u32 sims;        // simulations finished since ctx_load
};
AZ_HD void ctx_load(const EngineView& E, u32 g, Ctx& c) {
  c.gs = E.games[g];
  c.T = E.trees[(size_t)g * kP + c.gs.player];
  c.rng = (E.rng_mode == 1) ? E.glob->global_rng : c.gs.rng;
  c.sims = 0;
  c.pr.valid = 0;  // the pending leaf's path (if any) is in HBM
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < kPathRegs; ++i) c.pr.blk[i] = kNil;
  c.pr.slots_lo = c.pr.slots_hi = 0;
}
AZ_HD void ctx_store(const EngineView& E, u32 g, Ctx& c) {
  path_flush(E, g, c.pr, (u32)c.T.path_len);  // the pending leaf's path: a later launch (or the move code) reads it
  if (E.rng_mode == 1) E.glob->global_rng = c.rng; else c.gs.rng = c.rng;
  E.trees[(size_t)g * kP + c.gs.player] = c.T;
  E.games[g] = c.gs;
  if (c.sims) E.cold[g].sims += c.sims;
}

// The move part of PlayManager::play()'s loop body (play_manager.cc:286-555). Works on the slot's state
// in HBM (the caller stores its register copy before and reloads it after), so nothing of the hot
// loop's state has its address taken. Returns true when the slot retired.
AZ_COLD bool play_move(const EngineView E, u32 g) {
  GameSlot gs = E.games[g];
  Pcg32 rng = (E.rng_mode == 1) ? E.glob->global_rng : gs.rng;
  const u32 cp = gs.player;
  TreeHdr T[kP];
  T[0] = E.trees[(size_t)g * kP + 0];
  T[1] = E.trees[(size_t)g * kP + 1];
  GameCold cold = E.cold[g];
  bool retired = false;

  float temp = E.start_temp;
  if (E.half_life != 0.0f) {
    const float lambda = fdiv(0.693f, E.half_life);
    temp = fsub(temp, E.final_temp);
    temp = fmul(temp, az_expf(fmul(-lambda, (float)gs.turn)));
    temp = fadd(temp, E.final_temp);
  }
  RootView R;
  root_view(E, T[cp], R);
  // resign_percent (play_manager.cc:305-337). The playthrough coin comes from the game's own stream (the
  // reference uses an unseedable thread_local engine for it).
  u32 resign_term = 0;
  if (E.resign_percent > 0.0f && !gs.playthrough) {
    float wld[3];
    mcts_root_value(T[cp], R, wld);
    const double resign_val = dsub(1.0, (double)E.resign_percent);
    u32 t = 0;
    if ((double)wld[0] > resign_val) t = cp + 1u;
    else if ((double)wld[1] > resign_val) t = ((cp + 1u) % 2u) + 1u;
    else if ((double)wld[2] > resign_val) t = (u32)kP + 1u;
    if (t != 0) {
      if (rng_uniform01(rng) < E.resign_playthrough_percent) gs.playthrough = 1;
      else resign_term = t;
    }
  }
  float pi[kA];
  u32 chosen;
  GumbelState GS[kP];
  if (E.gumbel_enabled) {
    GS[0] = E.gum[(size_t)g * kP + 0];
    GS[1] = E.gum[(size_t)g * kP + 1];
  }
  if (E.gumbel_enabled && !gs.capped && GS[cp].initialized && GS[cp].n_surv > 0) {
    // gumbel_final_action (mcts.cc:375-401): argmax over the surviving candidates of g + logit + sigma(q_hat)
    u32 max_visit = 0;
    for (u32 j = 0; j < R.k; ++j)
      if (R.n[j] > max_visit) max_visit = R.n[j];
    const float sigma_scale = gumbel_sigma_scale(E, max_visit);
    u32 best = GS[cp].survivors & 15u;
    float best_score = -INFINITY;
    for (u32 r = 0; r < (u32)GS[cp].n_surv; ++r) {
      const u32 c = (GS[cp].survivors >> (4u * r)) & 15u;
      const float logit = az_logf(fadd(R.pol[c], AZ_GUMBEL_LOG_FLOOR));
      const float q_hat = R.n[c] > 0 ? R.q[c] : 0.0f;
      const float score = fadd(fadd(GS[cp].g[c], logit), fmul(sigma_scale, q_hat));
      if (score > best_score) { best_score = score; best = c; }
    }
    chosen = R.mv[best];
  } else {
    // PUCT acting — and Gumbel's fallback when the state never initialised: pick_move(probs(0)) (mcts.cc:379-381)
    mcts_probs(R, (E.gumbel_enabled && !gs.capped) ? 0.0f : temp, pi);
    chosen = mcts_pick_move(rng, pi);
  }
  if (chosen >= (u32)kA) { at_or(&E.glob->error, B2AZ_DEVERR_MOVE); chosen = R.k ? R.mv[0] : 0u; }
  if (E.history_enabled && !gs.capped) {
    float target[kA];
    if (E.gumbel_enabled) gumbel_improved_policy(E, T[cp], target);  // play_manager.cc:411-417
    else if (E.policy_target_pruning && E.epsilon > 0.0f) mcts_probs_pruned(E, T[cp], R, 1.0f, target);
    else mcts_probs(R, 1.0f, target);
    if (gs.hist_n < (u32)kMaxHist) {
      HistEntry h;
      h.p0 = gs.p0; h.p1 = gs.p1; h.player = gs.player; h.result = 0; h.pad_[0] = h.pad_[1] = 0;
      for (int m = 0; m < kA; ++m) h.pi[m] = target[m];
      E.hist_partial[(size_t)g * kMaxHist + gs.hist_n] = h;
    }
    ++gs.hist_n;
  }
  const float ald = (T[cp].depth == 0) ? 0.0f : fdiv((float)T[cp].total_leaf_depth, (float)T[cp].depth);
  const float ent = mcts_root_entropy(T[cp], R);
  if (!gs.capped) {
    cold.total_avg_leaf_depth += (double)ald;
    cold.total_search_entropy += (double)ent;
    ++gs.full_move_count;
  } else {
    cold.fast_total_avg_leaf_depth += (double)ald;
    cold.fast_total_search_entropy += (double)ent;
    ++gs.fast_move_count;
  }
  cold.total_valid_moves += (double)T[cp].k;
  ++gs.move_count;
  C4State s;
  s.p[0] = gs.p0; s.p[1] = gs.p1; s.turn = gs.turn; s.player = gs.player;
  const u32 vm_before = c4_valid_mask(s);
  for (int seat = 0; seat < kP; ++seat) {
    update_root(E, T[seat], chosen, vm_before, rng);
    if (E.gumbel_enabled) gumbel_reset(GS[seat]);  // update_root ends with reset_gumbel_state() (mcts.cc:172)
  }
  if (!c4_play(s, chosen)) at_or(&E.glob->error, B2AZ_DEVERR_MOVE);
  gs.p0 = s.p[0]; gs.p1 = s.p[1]; gs.turn = s.turn; gs.player = s.player;
  ++cold.nmoves;
  u32 term = c4_terminal(s);
  if (term == 0 && resign_term != 0) term = resign_term;  // play_manager.cc:440-444
  else resign_term = 0;
  if (term != 0) {
    // ---- game over: flush history newest-first (:448-460), accumulate (:463-505), restart
    if (E.history_enabled && gs.hist_n > 0) {
      const u32 cnt = gs.hist_n < (u32)kMaxHist ? gs.hist_n : (u32)kMaxHist;
      const unsigned long long at = at_add64(&E.glob->hist_written, (unsigned long long)cnt);
      const unsigned long long rd = ld_volatile(&E.glob->hist_read);
      if (at + cnt - rd > (unsigned long long)E.hist_capacity) at_or(&E.glob->error, B2AZ_DEVERR_HIST);
      for (u32 i = 0; i < cnt; ++i) {
        HistEntry h = E.hist_partial[(size_t)g * kMaxHist + (cnt - 1u - i)];
        h.result = (u8)term;
        E.hist_out[(at + i) % (unsigned long long)E.hist_capacity] = h;
      }
    }
    gs.hist_n = 0;
    Globals* G = E.glob;
    at_add64(&G->wins[term - 1u], 1ULL);
    at_add64(&G->perm_wins[slot_perm(E, g)][term - 1u], 1ULL);
    if (resign_term != 0) at_add64(&G->resign_wins[resign_term - 1u], 1ULL);
    at_add(&G->games_completed, 1u);
    at_add64(&G->game_length, (unsigned long long)gs.turn);
    at_addd(&G->total_avg_leaf_depth, cold.total_avg_leaf_depth);
    at_addd(&G->total_search_entropy, cold.total_search_entropy);
    at_addd(&G->fast_total_avg_leaf_depth, cold.fast_total_avg_leaf_depth);
    at_addd(&G->fast_total_search_entropy, cold.fast_total_search_entropy);
    at_addd(&G->total_valid_moves, cold.total_valid_moves);
    at_add64(&G->total_move_count, (unsigned long long)gs.move_count);
    at_add64(&G->full_move_count, (unsigned long long)gs.full_move_count);
    at_add64(&G->fast_move_count, (unsigned long long)gs.fast_move_count);
    const u32 started = at_add(&G->games_started, 1u);
    cold.total_avg_leaf_depth = cold.total_search_entropy = 0.0;
    cold.fast_total_avg_leaf_depth = cold.fast_total_search_entropy = 0.0;
    cold.total_valid_moves = 0.0;
    gs.move_count = gs.full_move_count = gs.fast_move_count = 0;
    for (int seat = 0; seat < kP; ++seat) {
      tree_free_pages(E, T[seat]);
      tree_reset(T[seat]);
      if (E.gumbel_enabled) gumbel_set_num_sims(GS[seat], 0u);  // make_mcts: a fresh MCTS has no sims target
    }
    ++cold.games_done;
    // play_manager.cc:506-509: a slot retires once games_to_play games have been STARTED by anybody. Which slot
    // that hits depends on completion order; with a per-slot quota every slot plays the same number of games.
    if (E.slot_quota ? cold.games_done >= E.slot_quota : started >= E.games_to_play) {
      retired = true;
      gs.active = 0;
      at_sub(&E.glob->active_games, 1u);
    } else {
      gs.p0 = gs.p1 = 0; gs.turn = 0; gs.player = 0;  // base_gs_->copy(); randomize_start() is a no-op
    }
  }
  if (!retired) {
    // a move has been played: update the playout cap (play_manager.cc:523-524; `&&` short-circuits the draw)
    gs.capped = (E.playout_cap && rng_uniform01(rng) < E.playout_cap_percent) ? 1 : 0;
    if (E.gumbel_enabled) {  // set_gumbel_num_sims for the seat that searches next (play_manager.cc:531-539)
      const u32 ncp = gs.player;
      gumbel_set_num_sims(GS[ncp], gs.capped ? (E.fast_search_uses_gumbel ? seat_budget(E, g, ncp, true) : 0u) : seat_budget(E, g, ncp, false));
    }
    if (!E.tree_reuse) {
      for (int seat = 0; seat < kP; ++seat) {
        tree_free_pages(E, T[seat]);
        tree_reset(T[seat]);
        // make_mcts AFTER set_gumbel_num_sims: the fresh MCTS has target 0, so without tree reuse the reference
        // never starts a Gumbel search after the first move (play_manager.cc:540-545) — kept as is
        if (E.gumbel_enabled) gumbel_set_num_sims(GS[seat], 0u);
      }
    } else {
      // the reused root gets the root temperature again and fresh noise (play_manager.cc:546-553,
      // mcts.cc:448-460)
      TreeHdr& NT = T[gs.player];
      if (NT.n > 0 && NT.k > 0 && NT.fc != kNil) {
        Block* B = E.blocks + NT.fc;
        float p8[kKMax];
        for (int j = 0; j < kKMax; ++j) p8[j] = blk_pol(B, (u32)j);
        bool changed = false;
        if (E.root_temp != 1.0f) {
          const float e = fdiv(1.0f, E.root_temp);
          float sum = 0.0f;
          for (u32 j = 0; j < NT.k; ++j) {
            p8[j] = az_powf(p8[j], e);
            sum = fadd(sum, p8[j]);
          }
          if (sum > 0.0f)
            for (u32 j = 0; j < NT.k; ++j) p8[j] = fdiv(p8[j], sum);
          changed = true;
        }
        if (E.epsilon > 0.0f && !gs.capped) {
          add_root_noise(E, rng, p8, NT.k);
          changed = true;
        }
        if (changed)
          for (u32 j = 0; j < NT.k; ++j) B->rec[j][2] = f2u(p8[j]);
      }
    }
  }
  E.trees[(size_t)g * kP + 0] = T[0];
  E.trees[(size_t)g * kP + 1] = T[1];
  if (E.gumbel_enabled) {
    E.gum[(size_t)g * kP + 0] = GS[0];
    E.gum[(size_t)g * kP + 1] = GS[1];
  }
  E.cold[g] = cold;
  if (E.rng_mode == 1) E.glob->global_rng = rng; else gs.rng = rng;
  E.games[g] = gs;
  return retired;
}

// One iteration of PlayManager::play()'s loop body for slot g (play_manager.cc:272-599).
AZ_HD void game_step(const EngineView& E, u32 g, Ctx& c) {
  if (!c.gs.active) return;
  bool retired = false;
  if (c.gs.initialized) {
    const u32 cp = c.gs.player;
    const bool noise = (E.epsilon > 0.0f) && !c.gs.capped;
    process_result(E, g, c.T, c.gs, c.rng, noise, c.pr);
    ++c.sims;
    const u32 goal = seat_budget(E, g, cp, c.gs.capped != 0);
    if (c.T.depth >= goal) {
      ctx_store(E, g, c);
      retired = play_move(E, g);
      ctx_load(E, g, c);
    }
  } else {
    c.gs.initialized = 1;
    c.gs.capped = (E.playout_cap && rng_uniform01(c.rng) < E.playout_cap_percent) ? 1 : 0;  // play_manager.cc:559-560
    if (E.gumbel_enabled) gumbel_arm(E, g, c.gs.player, c.gs.capped != 0);                       // :562-570
  }
  if (!retired) find_leaf(E, g, c.T, c.gs, c.rng, c.pr);
}

// n_steps iterations of game_step() for ONE slot with the two nested loops (steps x tree levels)
// flattened into a single loop that advances the game by one tree level per iteration. In a warp of 32
// independent games every iteration then holds exactly one block-load wait for EVERY game, instead of
// the warp idling until its deepest descent is done (measured: mean path 3, deepest of 32 about 7).
// The per-game order of operations — hence every result — is the one of game_step().
template <bool GB = true>
AZ_HD void run_flat(const EngineView& E, u32 g, Ctx& c, u32 n_steps) {
  Descent D;
  bool in_descent = false;
  u32 left = n_steps, hits = 0;
  while (left > 0) {
#if defined(__CUDA_ARCH__) && B2AZ_GATE > 0
    // Step boundaries are the long, divergent part of the loop body (backprop, expansion bookkeeping, maybe a
    // move). Gate them: the lanes of a warp that reached their leaf wait until at least B2AZ_GATE of them can
    // cross the boundary together (or nobody is descending any more), so that code runs with fuller warps.
    // Per-game order of operations is untouched, so results do not depend on the gate. Measured (r17): slower,
    // 1.33 G -> 1.25 G sims/s at 12 lanes — the idle lanes cost more than the fuller warps save.
    const unsigned act = __activemask();
    const unsigned rdy = __ballot_sync(act, !in_descent);
    if (!in_descent && __popc(rdy) < B2AZ_GATE && rdy != act) continue;
#endif
    if (!in_descent) {  // step boundary: finish the previous simulation, maybe play a move, start a descent
      if (!c.gs.active) break;
      bool retired = false;
      if (c.gs.initialized) {
        const u32 cp = c.gs.player;
        const bool noise = (E.epsilon > 0.0f) && !c.gs.capped;
        process_result(E, g, c.T, c.gs, c.rng, noise, c.pr);
        ++c.sims;
        const u32 goal = seat_budget(E, g, cp, c.gs.capped != 0);
        if (c.T.depth >= goal) {
          ctx_store(E, g, c);
          retired = play_move(E, g);
          ctx_load(E, g, c);
        }
      } else {
        c.gs.initialized = 1;
        c.gs.capped = (E.playout_cap && rng_uniform01(c.rng) < E.playout_cap_percent) ? 1 : 0;
        if (GB && E.gumbel_enabled) gumbel_arm(E, g, c.gs.player, c.gs.capped != 0);
      }
      if (retired) break;
      descent_begin<GB>(E, g, c.T, c.gs, D, c.pr, c.rng);
      in_descent = true;
    }
    if (descent_more(D) && descent_level<GB>(E, g, D, c.pr)) continue;
    descent_finish(E, g, c.T, c.gs, c.rng, D);
    in_descent = false;
    if (E.eval_type == 0) {
      // a cache hit is an answered leaf: go on with the next simulation in the same launch (the reference
      // re-queues the game for MCTS at once, play_manager.cc:589-594); bounded so a launch stays short
      if (leaf_emit(E, g, c.gs, D.s, hits < E.hit_cap)) {
        ++hits;
        continue;
      }
    }
    --left;
  }
}

// run_flat() with the warp kept in LOCK STEP: all 32 games of a warp cross the step boundary together, then descend
// (a lane whose descent has ended waits for the deepest one), then expand together. Per game the order of operations is
// run_flat()'s, so the results are the same; per warp every piece of code runs ONCE per simulation with (nearly) all
// lanes active — run_flat() executes the ~1,600-instruction boundary code and the ~500-instruction expansion code in
// almost every iteration for the few lanes that happen to be there (9.6 of 32 lanes active per issued instruction,
// profiles/r49_k_step_ncu_summary.json; B2AZ_GATE = 32 measured 1.97 G vs 1.51 G simulations/s, profiles/r2o).
#if defined(__CUDA_ARCH__)
#define AZ_WARP_ANY(p) (__any_sync(0xFFFFFFFFu, (p)) != 0)
#else
#define AZ_WARP_ANY(p) (p)
#endif
template <bool GB = true, class PR = PathRegs, bool PX = true>
AZ_HD void run_sync(const EngineView& E, u32 g, Ctx& c, PR& pr, u32 n_steps, bool alive) {
  Descent D;
  u32 left = n_steps, hits = 0;
  alive = alive && c.gs.active;
  for (;;) {
    // ---- step boundary: finish the previous simulation, maybe play a move, start a descent
    bool go = alive && left > 0;
    if (go) {
      bool retired = false;
      if (c.gs.initialized) {
        const u32 cp = c.gs.player;
        const bool noise = (E.epsilon > 0.0f) && !c.gs.capped;
        process_result<PR, PX>(E, g, c.T, c.gs, c.rng, noise, pr);
        ++c.sims;
        const u32 goal = seat_budget<PX>(E, g, cp, c.gs.capped != 0);
        if (c.T.depth >= goal) {
          ctx_store(E, g, c);  // the path is empty here: process_result has just consumed it
          retired = play_move(E, g);
          ctx_load(E, g, c);
        }
      } else {
        c.gs.initialized = 1;
        c.gs.capped = (E.playout_cap && rng_uniform01(c.rng) < E.playout_cap_percent) ? 1 : 0;
        if (GB && E.gumbel_enabled) gumbel_arm(E, g, c.gs.player, c.gs.capped != 0);
      }
      if (retired) {
        alive = false;
        go = false;
      } else {
        descent_begin<GB>(E, g, c.T, c.gs, D, pr, c.rng);
      }
    }
    if (!AZ_WARP_ANY(go)) break;
    // ---- descent: one tree level per iteration for every game that is still on its way down
    bool desc = go && descent_more(D);
    while (AZ_WARP_ANY(desc)) {
      if (desc) desc = descent_level<GB>(E, g, D, pr) && descent_more(D);
    }
    // ---- leaf: expansion, and the leaf batch / position cache with the NN evaluator
    if (go) {
      descent_finish(E, g, c.T, c.gs, c.rng, D);
      bool hit = false;
      if (E.eval_type == 0) {
        hit = leaf_emit<PX>(E, g, c.gs, D.s, hits < E.hit_cap);
        if (hit) ++hits;
      }
      if (!hit) --left;
    }
  }
}

}  // namespace b2az
