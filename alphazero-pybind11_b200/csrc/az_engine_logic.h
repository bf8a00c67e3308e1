// az_engine_logic.h — the self-play hot path as cooperative device functions.
//
// One GROUP of W lanes (W = 1..32, a power of two; W lanes of one warp) owns one game slot and runs,
// per step, exactly one iteration of PlayManager::play()'s loop body (play_manager.cc:272-599):
//     process_result -> [play a move] -> find_leaf
// All control flow is group-uniform: every lane keeps the same copy of the small state (tree header,
// game slot, RNG), loads of child statistics are spread over the lanes (one 32 B sector per field
// for a Connect4 block) and all-gathered with warp shuffles, and every lane then evaluates the same
// scalar code. That keeps the float operation ORDER identical to the reference's sequential loops
// (SURVEY.md Appendix A "Float order"), which is what makes visit counts bit-exact.
//
// Everything here is __host__ __device__: the host instantiation (W = 1, plain memory instead of
// atomics/shuffles) exists only for tests/cpp/engine_host_capi.cc, which lets the CPU test-suite
// compare this logic with the oracle without a GPU. The product (az_engine.cu) only instantiates
// the device side and fails loudly without CUDA.
#pragma once

#include "az_engine_types.h"

namespace b2az {

#define B2AZ_DEVERR_POOL 1u
#define B2AZ_DEVERR_HIST 2u
#define B2AZ_DEVERR_MOVE 4u
#define B2AZ_DEVERR_DEPTH 8u

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
AZ_HD unsigned long long at_cas64(unsigned long long* p, unsigned long long cmp, unsigned long long val) {
#if defined(__CUDA_ARCH__)
  return atomicCAS(p, cmp, val);
#else
  unsigned long long o = *p; if (o == cmp) *p = val; return o;
#endif
}
AZ_HD void mem_fence() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#endif
}
template <typename T>
AZ_HD T ld_volatile(const T* p) { return *reinterpret_cast<const volatile T*>(p); }

// ------------------------------------------------------------------------------------ lane groups
template <int W>
struct Grp {
  AZ_HD static int lane() {
#if defined(__CUDA_ARCH__)
    return (int)(threadIdx.x & (W - 1));
#else
    return 0;
#endif
  }
  AZ_HD static unsigned mask() {
#if defined(__CUDA_ARCH__)
    if (W == 32) return 0xFFFFFFFFu;
    return ((1u << (W & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(W - 1));
#else
    return 1u;
#endif
  }
  template <typename T>
  AZ_HD static T bcast(T v, int src) {
#if defined(__CUDA_ARCH__)
    if (W == 1) return v;
    return __shfl_sync(mask(), v, src, W);
#else
    (void)src;
    return v;
#endif
  }
  AZ_HD static void sync() {
#if defined(__CUDA_ARCH__)
    if (W > 1) __syncwarp(mask());
#endif
  }
};

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

// ------------------------------------------------------------------------------------ page pool
// Sharded lock-free stacks of free pages; head = (tag << 32) | top, next links in page_next[].
AZ_HD u32 pool_pop_page(const EngineView& E, u32 home) {
  for (u32 s = 0; s < (u32)kNumStacks; ++s) {
    unsigned long long* head = &E.stack_head[(home + s) % (u32)kNumStacks];
    unsigned long long old = ld_volatile(head);
    while ((u32)old != kNil) {
      const u32 top = (u32)old;
      const u32 nxt = ld_volatile(&E.page_next[top]);
      const unsigned long long nw = (((old >> 32) + 1ULL) << 32) | (unsigned long long)nxt;
      const unsigned long long got = at_cas64(head, old, nw);
      if (got == old) return top;
      old = got;
    }
  }
  return kNil;
}
AZ_HD void pool_push_chain(const EngineView& E, u32 home, u32 first, u32 last) {
  unsigned long long* head = &E.stack_head[home % (u32)kNumStacks];
  unsigned long long old = ld_volatile(head);
  for (;;) {
    E.page_next[last] = (u32)old;
    mem_fence();
    const unsigned long long nw = (((old >> 32) + 1ULL) << 32) | (unsigned long long)first;
    const unsigned long long got = at_cas64(head, old, nw);
    if (got == old) return;
    old = got;
  }
}
// Free every page of a tree's chain (lane 0 of the group).
template <int W>
AZ_HD void tree_free_pages(const EngineView& E, u32 home, u32 first_page) {
  if (first_page == kNil) return;
  if (Grp<W>::lane() == 0) {
    u32 last = first_page;
    for (;;) {
      const u32 nx = E.page_next[last];
      if (nx == kNil) break;
      last = nx;
    }
    pool_push_chain(E, home, first_page, last);
  }
}
// Bump-allocate a child block of kk nodes (padded to 8) in the tree's arena. Group-uniform.
template <int W>
AZ_HD u32 tree_alloc_block(const EngineView& E, TreeHdr& T, u32 home, u32 kk) {
  const u32 need = (kk + 7u) & ~7u;
  if (T.cur_page == kNil || T.bump + need > kPageNodes) {
    u32 p = kNil;
    if (Grp<W>::lane() == 0) {
      p = pool_pop_page(E, home);
      if (p != kNil) {
        E.page_next[p] = kNil;
        E.page_fill[p] = 0;
        if (T.cur_page != kNil) {
          E.page_fill[T.cur_page] = T.bump;
          E.page_next[T.cur_page] = p;
        }
      } else {
        at_or(&E.glob->error, B2AZ_DEVERR_POOL);
      }
    }
    Grp<W>::sync();  // page_next / page_fill written by lane 0 are read by every lane (Cheney scan)
    p = Grp<W>::bcast(p, 0);
    if (p == kNil) return kNil;
    if (T.cur_page == kNil) T.first_page = p;
    T.cur_page = p;
    T.bump = 0;
  }
  const u32 base = (T.cur_page << kPageLog2) + T.bump;
  T.bump += need;
  return base;
}
AZ_HD void tree_reset(TreeHdr& T) {  // a freshly constructed MCTS (mcts.h:52-73): root_ = Node{}
  T.q = T.d = T.v = T.policy = 0.0f;
  T.n = 0;
  T.fc = kNil;
  T.k = 0;
  T.player = 0;
  T.term = 0;
  T.move = 0;
  T.path_len = 0;
  T.depth = 0;
  T.total_leaf_depth = 0;
  T.first_page = T.cur_page = kNil;
  T.bump = 0;
  T.leaf = kRootRef;
  T.pad_[0] = T.pad_[1] = 0;
}

// ------------------------------------------------------------------------------------ child blocks
struct Kids {  // statistics of one (<= 8 wide) child block, replicated in every lane
  u32 n[kKMax];
  float pol[kKMax];
  float q[kKMax];
};
// Coalesced load of n/pol/q (one 32 B sector each) + all-gather inside the group.
template <int W, bool WITH_Q>
AZ_HD void kids_load(const EngineView& E, u32 fc, Kids& K) {
  constexpr int R = (kKMax + W - 1) / W;
  u32 mn[R];
  float mp[R], mq[R];
  const int lane = Grp<W>::lane();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * W + lane;
    mn[r] = 0; mp[r] = 0.0f; mq[r] = 0.0f;
    if (j < kKMax) {
      mn[r] = E.n[fc + j];
      mp[r] = E.pol[fc + j];
      if (WITH_Q) mq[r] = E.q[fc + j];
    }
  }
#pragma unroll
  for (int i = 0; i < kKMax; ++i) {
    K.n[i] = Grp<W>::bcast(mn[i / W], i % W);
    K.pol[i] = Grp<W>::bcast(mp[i / W], i % W);
    K.q[i] = WITH_Q ? Grp<W>::bcast(mq[i / W], i % W) : 0.0f;
  }
}
template <int W>
AZ_HD void moves_load(const EngineView& E, u32 fc, u32* mv8) {
  constexpr int R = (kKMax + W - 1) / W;
  u32 mm[R];
  const int lane = Grp<W>::lane();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * W + lane;
    mm[r] = (j < kKMax) ? (u32)E.mv[fc + j] : 0u;
  }
#pragma unroll
  for (int i = 0; i < kKMax; ++i) mv8[i] = Grp<W>::bcast(mm[i / W], i % W);
}
// Store one float per child from a replicated array (lane j%W writes child j).
template <int W>
AZ_HD void kids_store_pol(const EngineView& E, u32 fc, u32 k, const float* p8) {
  const int lane = Grp<W>::lane();
#pragma unroll
  for (int j = 0; j < kKMax; ++j)
    if ((j % W) == lane && (u32)j < k) E.pol[fc + j] = p8[j];
}

// Node::best_child (mcts.cc:130-149) + Node::uct (mcts.cc:123-128); n_in_flight is always 0 on
// this path (the WU-UCT variant is not used by PlayManager).
AZ_HD int best_child(const Kids& K, u32 k, u32 parent_n, float parent_v, float cpuct, float fpu_reduction) {
  float seen = 0.0f;
  for (u32 j = 0; j < k; ++j)
    if (K.n[j] > 0) seen = fadd(seen, K.pol[j]);
  const float fpu_value = fsub(parent_v, fmul(fpu_reduction, fsqrt(seen)));
  const float sqrt_n = fsqrt((float)parent_n);
  int best_i = 0;
  float best_u = 0.0f;
  for (u32 j = 0; j < k; ++j) {
    const float base = (K.n[j] == 0) ? fpu_value : K.q[j];
    const float u = fadd(base, fdiv(fmul(fmul(cpuct, K.pol[j]), sqrt_n), (float)(K.n[j] + 1u)));
    if (j == 0 || u > best_u) {
      best_u = u;
      best_i = (int)j;
    }
  }
  return best_i;
}

struct Leaf {  // what find_leaf hands to the evaluator
  C4State s;
  u32 k;     // legal moves at the leaf
  u32 term;  // terminal code of the leaf node
};

// MCTS::find_leaf (mcts.cc:462-498), PUCT branch.
template <int W>
AZ_HD void find_leaf(const EngineView& E, u32 g, u32 tree, TreeHdr& T, const GameSlot& gs, Pcg32& rng, Leaf& out) {
  C4State s;
  s.p[0] = gs.p0; s.p[1] = gs.p1; s.turn = gs.turn; s.player = gs.player;
  u32* path = E.path + (size_t)g * kMaxPath;
  u32 cur = kRootRef;
  u32 cur_n = T.n, cur_term = T.term, cur_fc = T.fc, cur_k = T.k;
  float cur_v = T.v;
  u32 plen = 0;
  const int lane = Grp<W>::lane();
  while (cur_n > 0 && cur_term == 0) {
    if (cur_k == 0 || plen >= (u32)kMaxPath) {  // cannot happen for a non-terminal Connect4 node
      at_or(&E.glob->error, B2AZ_DEVERR_DEPTH);
      break;
    }
    Kids K;
    kids_load<W, true>(E, cur_fc, K);
    const float fpu = (cur == kRootRef && E.root_fpu_zero) ? 0.0f : E.fpu_reduction;
    const int j = best_child(K, cur_k, cur_n, cur_v, E.cpuct, fpu);
    const u32 c = cur_fc + (u32)j;
    if (lane == 0) path[plen] = c;
    ++plen;
    const u32 move = E.mv[c];
    c4_play(s, move);
    cur = c;
    cur_n = K.n[j];
    if (cur_n > 0) {
      const NodeRec r = E.rec[c];
      cur_term = r.term; cur_fc = r.fc; cur_k = r.k; cur_v = r.v;
    }
  }
  T.total_leaf_depth += plen;
  out.term = cur_term;
  out.k = cur_k;
  if (cur_n == 0) {
    // expand: current_->player, scores, add_children(valid_moves) incl. the shuffle (mcts.cc:490-496, 93-101)
    const u32 term = c4_terminal(s);
    const u32 vm = c4_valid_mask(s);
    u32 moves[kKMax];
    u32 k = 0;
#pragma unroll
    for (u32 w = 0; w < (u32)kA; ++w)
      if ((vm >> w) & 1u) moves[k++] = w;
    rng_shuffle(rng, moves, k);
    // Children of a terminal node are never visited (selection stops at scores != nullptr,
    // mcts.cc:473) — their RNG draws are consumed above, their storage is skipped.
    const u32 kk = term ? 0u : k;
    u32 fc = kNil;
    if (kk > 0) {
      fc = tree_alloc_block<W>(E, T, g, kk);
      if (fc != kNil) {
#pragma unroll
        for (int j = 0; j < kKMax; ++j)
          if ((j % W) == lane) {
            E.n[fc + j] = 0u;  // full-sector write; pads stay n = 0 forever
            E.mv[fc + j] = (u16)((u32)j < kk ? moves[j] : 0u);
          }
      }
    }
    const u32 k_eff = (fc == kNil) ? 0u : kk;
    if (cur == kRootRef) {
      T.player = s.player; T.term = (u8)term; T.fc = fc; T.k = (u16)k_eff;
    } else if (lane == 0) {
      // only the structural half of the record; v/d are written by the first backprop
      NodeRec r = E.rec[cur];
      r.fc = fc; r.k = (u16)k_eff; r.player = s.player; r.term = (u8)term;
      E.rec[cur] = r;
    }
    out.term = term;
    out.k = k;  // dumb_eval counts every legal move, terminal or not
  } else {
    out.k = (u32)popc32(c4_valid_mask(s));
  }
  T.leaf = cur;
  T.path_len = (u16)plen;
  out.s = s;
}

// MCTS::add_root_noise (mcts.cc:403-446). `pol` = root child priors in child order (replicated).
AZ_HD void add_root_noise(const EngineView& E, Pcg32& rng, float* pol, u32 k) {
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

// MCTS::process_result (mcts.cc:500-555).
template <int W>
AZ_HD void process_result(const EngineView& E, u32 g, TreeHdr& T, const GameSlot& gs, Pcg32& rng, bool noise_enabled) {
  float value[kP + 1];
  const u32 leaf = T.leaf;
  u32 lterm, lfc, lk, lplayer;
  if (leaf == kRootRef) {
    lterm = T.term; lfc = T.fc; lk = T.k; lplayer = T.player;
  } else {
    const NodeRec r = E.rec[leaf];
    lterm = r.term; lfc = r.fc; lk = r.k; lplayer = r.player;
  }
  if (lterm != 0) {
    value[0] = (lterm == 1) ? 1.0f : 0.0f;
    value[1] = (lterm == 2) ? 1.0f : 0.0f;
    value[2] = (lterm == 3) ? 1.0f : 0.0f;
  } else {
    float p8[kKMax];
    if (E.eval_type == 1) {  // dumb_eval (game_state.h:160-173)
      const float third = (float)(1.0 / 3.0);
      value[0] = value[1] = value[2] = third;
      const float sum = (float)(gs.leaf_k & 0xFFu);  // Vector<uint8_t>::sum() is uint8-typed
      for (u32 j = 0; j < lk; ++j) p8[j] = fdiv(1.0f, sum);
    } else {
      const float* vrow = E.ev_v + (size_t)gs.eval_row * (kP + 1);
      const float* prow = E.ev_pi + (size_t)gs.eval_row * kA;
      value[0] = vrow[0]; value[1] = vrow[1]; value[2] = vrow[2];
      u32 mv8[kKMax];
      if (lk > 0) moves_load<W>(E, lfc, mv8);
      for (u32 j = 0; j < lk; ++j) p8[j] = prow[mv8[j]];
    }
    // set_policy_normalized (mcts.cc:109-121)
    const bool apply_temp = (leaf == kRootRef) && (E.root_temp != 1.0f);
    const float inv_temp = fdiv(1.0f, E.root_temp);
    float sum = 0.0f;
    for (u32 j = 0; j < lk; ++j) {
      if (apply_temp) p8[j] = az_powf(p8[j], inv_temp);
      sum = fadd(sum, p8[j]);
    }
    for (u32 j = 0; j < lk; ++j) p8[j] = fdiv(p8[j], sum);
    if (leaf == kRootRef && noise_enabled && lk > 0) add_root_noise(E, rng, p8, lk);
    if (lk > 0) kids_store_pol<W>(E, lfc, lk, p8);
  }
  // backprop (mcts.cc:527-545): level i updates node path[i]; its parent is path[i-1] or the root
  const u32* path = E.path + (size_t)g * kMaxPath;
  const float dshare = fdiv(value[kP], (float)kP);
  const u32 plen = T.path_len;
  const int lane = Grp<W>::lane();
  for (u32 i = (u32)lane; i < plen; i += W) {
    const u32 c = path[i];
    const u32 pp = (i == 0) ? (u32)T.player : (u32)E.rec[path[i - 1]].player;
    const float v = fadd(value[pp], dshare);
    const u32 nc = E.n[c];
    NodeRec r = E.rec[c];
    const float qc = nc ? E.q[c] : 0.0f;
    const float dc = nc ? r.d : 0.0f;
    E.q[c] = fdiv(fadd(fmul(qc, (float)nc), v), (float)(nc + 1u));
    r.d = fdiv(fadd(fmul(dc, (float)nc), value[kP]), (float)(nc + 1u));
    if (nc == 0) r.v = fadd(value[r.player], dshare);
    E.rec[c] = r;
    E.n[c] = nc + 1u;
  }
  if (T.n == 0) {
    T.v = fadd(value[lplayer], dshare);  // root_.player == the leaf's player when the root is the leaf
    T.d = value[kP];
  }
  ++T.depth;
  ++T.n;
  T.path_len = 0;
  (void)lplayer;
}

// counts()/probs() family works on dense arrays in MOVE order (mcts.cc:557-618).
struct RootView {
  u32 k;
  u32 mv[kKMax];
  Kids K;
};
template <int W>
AZ_HD void root_view(const EngineView& E, const TreeHdr& T, RootView& R) {
  R.k = T.k;
  if (T.k > 0) {
    kids_load<W, true>(E, T.fc, R.K);
    moves_load<W>(E, T.fc, R.mv);
  }
}
AZ_HD float sum7(const float* a) {  // Vector::sum(): sequential, starting from 0
  float s = 0.0f;
  for (int m = 0; m < kA; ++m) s = fadd(s, a[m]);
  return s;
}
// MCTS::probs(temp) (mcts.cc:575-618)
AZ_HD void mcts_probs(const RootView& R, float temp, float* probs) {
  u32 counts[kA];
  for (int m = 0; m < kA; ++m) counts[m] = 0;
  for (u32 j = 0; j < R.k; ++j) counts[R.mv[j]] = R.K.n[j];
  float count_sum = 0.0f;
  for (int m = 0; m < kA; ++m) count_sum = fadd(count_sum, (float)counts[m]);
  if (count_sum == 0.0f) {
    for (int m = 0; m < kA; ++m) probs[m] = 0.0f;
    for (u32 j = 0; j < R.k; ++j) probs[R.mv[j]] = R.K.pol[j];
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
AZ_HD void mcts_probs_pruned(const EngineView& E, const TreeHdr& T, const RootView& R, float temp, float* out) {
  if (T.n <= 1) { mcts_probs(R, temp, out); return; }
  const float es = fmul(E.cpuct, fsqrt((float)T.n));
  float best_sel = -1e30f;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.K.n[j] == 0) continue;
    const float sel = fadd(R.K.q[j], fdiv(fmul(es, R.K.pol[j]), (float)(R.K.n[j] + 1u)));
    if (sel > best_sel) best_sel = sel;
  }
  float pruned[kA];
  for (int m = 0; m < kA; ++m) pruned[m] = 0.0f;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.K.n[j] == 0) continue;
    const float gap = fsub(best_sel, R.K.q[j]);
    float desired;
    if (gap <= 0.0f) desired = (float)R.K.n[j];
    else desired = fsub(fdiv(fmul(es, R.K.pol[j]), gap), 1.0f);
    pruned[R.mv[j]] = std_min((float)R.K.n[j], std_max(0.0f, desired));
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
AZ_HD float mcts_root_entropy(const TreeHdr& T, const RootView& R) {
  const float k = (float)R.k;
  if (R.k <= 1 || T.n <= 1) return 0.0f;
  const float log_k = az_logf(k);
  float entropy = 0.0f;
  const float total_n = (float)T.n;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.K.n[j] > 0) {
      const float p = fdiv((float)R.K.n[j], total_n);
      entropy = fsub(entropy, fmul(p, az_logf(p)));
    }
  }
  return fdiv(entropy, log_k);
}
// MCTS::root_value (mcts.h:78-100) -> (w, l, d)
AZ_HD void mcts_root_value(const EngineView& E, const TreeHdr& T, const RootView& R, float* wld) {
  float q = 0.0f, d = 0.0f;
  bool found = false;
  for (u32 j = 0; j < R.k; ++j) {
    if (R.K.n[j] > 0 && R.K.q[j] > q) {
      q = R.K.q[j];
      d = E.rec[T.fc + j].d;
      found = true;
    }
  }
  if (!found && T.n > 0) { q = T.v; d = T.d; }
  const float w = fsub(q, fdiv(d, (float)kP));
  wld[0] = w;
  wld[1] = (float)dsub(dsub(1.0, (double)w), (double)d);  // `1.0 - w - d` is evaluated in double
  wld[2] = d;
}

// MCTS::apply_root_policy_temp (mcts.cc:448-460) on the (reused) root's children
template <int W>
AZ_HD void apply_root_policy_temp(const EngineView& E, const TreeHdr& T, float* pol8_out, bool* changed) {
  *changed = false;
  if (E.root_temp == 1.0f || T.k == 0) return;
  Kids K;
  kids_load<W, false>(E, T.fc, K);
  const float e = fdiv(1.0f, E.root_temp);
  float sum = 0.0f;
  for (u32 j = 0; j < T.k; ++j) {
    pol8_out[j] = az_powf(K.pol[j], e);
    sum = fadd(sum, pol8_out[j]);
  }
  if (sum > 0.0f)
    for (u32 j = 0; j < T.k; ++j) pol8_out[j] = fdiv(pol8_out[j], sum);
  *changed = true;
}

// MCTS::update_root (mcts.cc:151-173). `vm_before` = valid-move mask of the game state BEFORE the
// move (update_root receives the pre-move state, play_manager.cc:436-439).
template <int W>
AZ_HD void update_root(const EngineView& E, u32 home, TreeHdr& T, u32 move, u32 vm_before, Pcg32& rng) {
  T.depth = 0;
  T.total_leaf_depth = 0;
  T.path_len = 0;
  T.leaf = kRootRef;
  const int lane = Grp<W>::lane();
  if (T.k == 0) {
    // root_.children.empty(): add_children(valid_moves()) shuffles children that are discarded by
    // the re-root onto the (unvisited) chosen child two lines later — only the RNG draws survive.
    rng_shuffle_discard(rng, (u32)popc32(vm_before));
    if (((vm_before >> move) & 1u) == 0u) at_or(&E.glob->error, B2AZ_DEVERR_MOVE);
    const u32 old_first = T.first_page;
    tree_reset(T);
    T.move = (u16)move;
    tree_free_pages<W>(E, home, old_first);
    return;
  }
  u32 mv8[kKMax];
  moves_load<W>(E, T.fc, mv8);
  int ci = -1;
  for (u32 j = 0; j < T.k; ++j)
    if (mv8[j] == move) { ci = (int)j; break; }
  const u32 old_first = T.first_page;
  if (ci < 0) {
    at_or(&E.glob->error, B2AZ_DEVERR_MOVE);
    tree_reset(T);
    tree_free_pages<W>(E, home, old_first);
    return;
  }
  const u32 c = T.fc + (u32)ci;
  // the chosen child becomes the root (Node tmp = std::move(*x); root_ = std::move(tmp))
  const u32 cn = E.n[c];
  TreeHdr N;
  tree_reset(N);
  N.policy = E.pol[c];
  N.move = (u16)move;
  N.n = cn;
  u32 src_fc = kNil, src_k = 0;
  if (cn > 0) {
    const NodeRec r = E.rec[c];
    N.q = E.q[c]; N.d = r.d; N.v = r.v; N.player = r.player; N.term = r.term;
    src_fc = r.fc; src_k = r.k;
  }
  // Cheney copy of the kept subtree into fresh pages, BFS order.
  if (src_k > 0 && src_fc != kNil) {
    const u32 nb = tree_alloc_block<W>(E, N, home, src_k);
    if (nb != kNil) {
      const u32 need = (src_k + 7u) & ~7u;
      for (u32 j = (u32)lane; j < need; j += W) {
        E.q[nb + j] = E.q[src_fc + j];
        E.pol[nb + j] = E.pol[src_fc + j];
        E.n[nb + j] = E.n[src_fc + j];
        E.mv[nb + j] = E.mv[src_fc + j];
        E.rec[nb + j] = E.rec[src_fc + j];
      }
      N.fc = nb;
      N.k = (u16)src_k;
      Grp<W>::sync();
      // scan the to-space; every expanded node found gets its child block copied behind
      u32 scan_page = N.first_page, scan_off = 0;
      for (;;) {
        const u32 limit = (scan_page == N.cur_page) ? N.bump : E.page_fill[scan_page];
        if (scan_off >= limit) {
          if (scan_page == N.cur_page) break;
          scan_page = E.page_next[scan_page];
          scan_off = 0;
          continue;
        }
        // 8 nodes at a time (blocks are 8-aligned): every lane inspects all 8 (replicated, uniform)
        const u32 base = (scan_page << kPageLog2) + scan_off;
        u32 n8[kKMax];
        {
          constexpr int R = (kKMax + W - 1) / W;
          u32 mn[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int j = r * W + lane;
            mn[r] = (j < kKMax) ? E.n[base + j] : 0u;
          }
#pragma unroll
          for (int i = 0; i < kKMax; ++i) n8[i] = Grp<W>::bcast(mn[i / W], i % W);
        }
        bool failed = false;
#pragma unroll 1
        for (int i = 0; i < kKMax; ++i) {
          if (n8[i] == 0) continue;
          const NodeRec r = E.rec[base + i];
          if (r.k == 0 || r.fc == kNil) continue;
          const u32 dst = tree_alloc_block<W>(E, N, home, r.k);
          if (dst == kNil) { failed = true; break; }
          const u32 need2 = ((u32)r.k + 7u) & ~7u;
          for (u32 j = (u32)lane; j < need2; j += W) {
            E.q[dst + j] = E.q[r.fc + j];
            E.pol[dst + j] = E.pol[r.fc + j];
            E.n[dst + j] = E.n[r.fc + j];
            E.mv[dst + j] = E.mv[r.fc + j];
            E.rec[dst + j] = E.rec[r.fc + j];
          }
          Grp<W>::sync();  // every lane has read rec[base + i] and finished its share of the copy
          if (lane == 0) {
            NodeRec r2 = r;
            r2.fc = dst;
            E.rec[base + i] = r2;
          }
          Grp<W>::sync();
        }
        if (failed) break;
        scan_off += kKMax;
      }
    }
  }
  T = N;
  Grp<W>::sync();
  tree_free_pages<W>(E, home, old_first);
}

// The canonical 4x6x7 planes are rebuilt from the compact position wherever they are needed
// (leaf batch, history drain): Connect4GS::canonicalized (connect4_gs.cc:131-149).

// One iteration of PlayManager::play()'s loop body for slot g (play_manager.cc:272-599).
template <int W>
AZ_HD void game_step(const EngineView& E, u32 g) {
  GameSlot gs = E.games[g];
  if (!gs.active) return;
  const int lane = Grp<W>::lane();
  Pcg32 rng = (E.rng_mode == 1) ? E.glob->global_rng : gs.rng;
  TreeHdr T[kP];
  T[0] = E.trees[(size_t)g * kP + 0];
  T[1] = E.trees[(size_t)g * kP + 1];
  bool retired = false;

  if (gs.initialized) {
    const u32 cp = gs.player;
    const bool noise = (E.epsilon > 0.0f) && !gs.capped;
    process_result<W>(E, g, T[cp], gs, rng, noise);
    Grp<W>::sync();
    ++gs.sims;
    const u32 goal = gs.capped ? E.cap_visits[cp] : E.visits[cp];
    if (T[cp].depth >= goal) {
      // ---------------------------------------------------------------- play a move (:286-555)
      float temp = E.start_temp;
      if (E.half_life != 0.0f) {
        const float lambda = fdiv(0.693f, E.half_life);
        temp = fsub(temp, E.final_temp);
        temp = fmul(temp, az_expf(fmul(-lambda, (float)gs.turn)));
        temp = fadd(temp, E.final_temp);
      }
      RootView R;
      root_view<W>(E, T[cp], R);
      float pi[kA];
      mcts_probs(R, temp, pi);
      u32 chosen = mcts_pick_move(rng, pi);
      if (chosen >= (u32)kA) { at_or(&E.glob->error, B2AZ_DEVERR_MOVE); chosen = R.k ? R.mv[0] : 0u; }
      if (E.history_enabled && !gs.capped) {
        float target[kA];
        if (E.policy_target_pruning && E.epsilon > 0.0f) mcts_probs_pruned(E, T[cp], R, 1.0f, target);
        else mcts_probs(R, 1.0f, target);
        if (lane == 0 && gs.hist_n < (u32)kMaxHist) {
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
        gs.total_avg_leaf_depth += (double)ald;
        gs.total_search_entropy += (double)ent;
        ++gs.full_move_count;
      } else {
        gs.fast_total_avg_leaf_depth += (double)ald;
        gs.fast_total_search_entropy += (double)ent;
        ++gs.fast_move_count;
      }
      gs.total_valid_moves += (double)T[cp].k;
      ++gs.move_count;
      C4State s;
      s.p[0] = gs.p0; s.p[1] = gs.p1; s.turn = gs.turn; s.player = gs.player;
      const u32 vm_before = c4_valid_mask(s);
      for (int seat = 0; seat < kP; ++seat) update_root<W>(E, g, T[seat], chosen, vm_before, rng);
      if (!c4_play(s, chosen)) at_or(&E.glob->error, B2AZ_DEVERR_MOVE);
      gs.p0 = s.p[0]; gs.p1 = s.p[1]; gs.turn = s.turn; gs.player = s.player;
      ++gs.nmoves;
      const u32 term = c4_terminal(s);
      if (term != 0) {
        // ---- game over: flush history newest-first (:448-460), accumulate (:463-505), restart
        if (E.history_enabled && lane == 0 && gs.hist_n > 0) {
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
        u32 started = 0;
        if (lane == 0) {
          Globals* G = E.glob;
          at_add64(&G->wins[term - 1u], 1ULL);
          at_add(&G->games_completed, 1u);
          at_add64(&G->game_length, (unsigned long long)gs.turn);
          at_addd(&G->total_avg_leaf_depth, gs.total_avg_leaf_depth);
          at_addd(&G->total_search_entropy, gs.total_search_entropy);
          at_addd(&G->fast_total_avg_leaf_depth, gs.fast_total_avg_leaf_depth);
          at_addd(&G->fast_total_search_entropy, gs.fast_total_search_entropy);
          at_addd(&G->total_valid_moves, gs.total_valid_moves);
          at_add64(&G->total_move_count, (unsigned long long)gs.move_count);
          at_add64(&G->full_move_count, (unsigned long long)gs.full_move_count);
          at_add64(&G->fast_move_count, (unsigned long long)gs.fast_move_count);
          started = at_add(&G->games_started, 1u);
        }
        started = Grp<W>::bcast(started, 0);
        gs.total_avg_leaf_depth = gs.total_search_entropy = 0.0;
        gs.fast_total_avg_leaf_depth = gs.fast_total_search_entropy = 0.0;
        gs.total_valid_moves = 0.0;
        gs.move_count = gs.full_move_count = gs.fast_move_count = 0;
        for (int seat = 0; seat < kP; ++seat) {
          const u32 old_first = T[seat].first_page;
          tree_reset(T[seat]);
          tree_free_pages<W>(E, g, old_first);
        }
        if (started >= E.games_to_play) {
          retired = true;  // play_manager.cc:506-509
          gs.active = 0;
          if (lane == 0) at_sub(&E.glob->active_games, 1u);
        } else {
          gs.p0 = gs.p1 = 0; gs.turn = 0; gs.player = 0;  // base_gs_->copy(); randomize_start() is a no-op
        }
      }
      if (!retired) {
        gs.capped = 0;  // playout-cap randomisation draws from an unseedable engine; not carried yet
        if (!E.tree_reuse) {
          for (int seat = 0; seat < kP; ++seat) {
            const u32 old_first = T[seat].first_page;
            tree_reset(T[seat]);
            tree_free_pages<W>(E, g, old_first);
          }
        } else {
          TreeHdr& NT = T[gs.player];
          if (NT.n > 0) {
            float p8[kKMax];
            bool changed = false;
            apply_root_policy_temp<W>(E, NT, p8, &changed);
            if (!changed && E.epsilon > 0.0f && NT.k > 0) {
              Kids K;
              kids_load<W, false>(E, NT.fc, K);
              for (u32 j = 0; j < NT.k; ++j) p8[j] = K.pol[j];
            }
            if (E.epsilon > 0.0f && !gs.capped && NT.k > 0) {
              add_root_noise(E, rng, p8, NT.k);
              changed = true;
            }
            if (changed) kids_store_pol<W>(E, NT.fc, NT.k, p8);
            Grp<W>::sync();
          }
        }
      }
    }
  } else {
    gs.initialized = 1;
    gs.capped = 0;
  }

  if (!retired) {
    const u32 cp = gs.player;
    Leaf leaf;
    find_leaf<W>(E, g, (u32)(g * kP + cp), T[cp], gs, rng, leaf);
    gs.leaf_k = leaf.k;
    if (E.eval_type == 0) {
      u32 row = 0;
      if (lane == 0) {
        row = at_add(&E.glob->leaf_count, 1u);
        E.leaf_p0[row] = leaf.s.p[0];
        E.leaf_p1[row] = leaf.s.p[1];
        E.leaf_player[row] = leaf.s.player;
        E.leaf_game[row] = g;
      }
      gs.eval_row = Grp<W>::bcast(row, 0);
    }
    Grp<W>::sync();
  }

  if (lane == 0) {
    E.trees[(size_t)g * kP + 0] = T[0];
    E.trees[(size_t)g * kP + 1] = T[1];
    if (E.rng_mode == 1) E.glob->global_rng = rng; else gs.rng = rng;
    E.games[g] = gs;
  }
  Grp<W>::sync();
}

}  // namespace b2az
