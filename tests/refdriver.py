"""ctypes face of oracle/_ref/libazref.so — the UNMODIFIED reference compiled against oracle/shim.

Test infrastructure only (tests/, bench.py cpu_baseline / --impl reference). Nothing in the product
imports this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libazref.so")


class PlayCfg(C.Structure):
    _fields_ = [
        ("games_to_play", C.c_uint32), ("concurrent_games", C.c_uint32), ("max_batch_size", C.c_uint32),
        ("max_cache_size", C.c_uint32), ("cache_shards", C.c_uint32), ("queue_shards", C.c_uint32),
        ("mcts_visits", C.c_uint32 * 2), ("cpuct", C.c_float), ("start_temp", C.c_float), ("final_temp", C.c_float),
        ("temp_decay_half_life", C.c_float), ("history_enabled", C.c_uint8), ("self_play", C.c_uint8),
        ("tree_reuse", C.c_uint8), ("playout_cap_randomization", C.c_uint8), ("epsilon", C.c_float),
        ("mcts_root_temp", C.c_float), ("playout_cap_depth", C.c_uint32), ("playout_cap_percent", C.c_float),
        ("fpu_reduction", C.c_float), ("root_fpu_zero", C.c_uint8), ("shaped_dirichlet", C.c_uint8),
        ("policy_target_pruning", C.c_uint8), ("gumbel_enabled", C.c_uint8), ("gumbel_m", C.c_uint32),
        ("gumbel_c_visit", C.c_float), ("gumbel_c_scale", C.c_float), ("gumbel_full", C.c_uint8),
        ("fast_search_uses_gumbel", C.c_uint8), ("eval_type", C.c_uint8), ("pad_", C.c_uint8),
        ("resign_percent", C.c_float), ("resign_playthrough_percent", C.c_float),
        ("has_groups", C.c_uint8), ("model_groups", C.c_uint8 * 2), ("n_seat_perms", C.c_uint8),
        ("seat_perms", (C.c_uint8 * 2) * 8), ("group_eval", C.c_uint8 * 2), ("pad2_", C.c_uint8 * 2),
    ]


class MctsCfg(C.Structure):
    _fields_ = [
        ("cpuct", C.c_float), ("num_players", C.c_uint32), ("num_moves", C.c_uint32), ("epsilon", C.c_float),
        ("root_policy_temp", C.c_float), ("fpu_reduction", C.c_float), ("relative_values", C.c_uint8),
        ("root_fpu_zero", C.c_uint8), ("shaped_dirichlet", C.c_uint8), ("gumbel_enabled", C.c_uint8),
        ("gumbel_m", C.c_uint32), ("gumbel_c_visit", C.c_float), ("gumbel_c_scale", C.c_float),
        ("gumbel_full", C.c_uint8),
    ]


_lib = None


def available():
    return os.path.exists(REF_LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB)
        vp, u32 = C.c_void_p, C.c_uint32
        L.azref_last_error.restype = C.c_char_p
        for name in ("azref_c4_new", "azref_c4_from_board", "azref_c4_copy", "azref_mcts_new", "azref_mcts_find_leaf",
                     "azref_pm_new_connect4", "azref_cache_new", "azref_rng_new"):
            getattr(L, name).restype = vp
        L.azref_c4_from_board.argtypes = [vp, C.c_int8, C.c_int32]
        for name in ("azref_c4_copy", "azref_c4_free", "azref_mcts_free", "azref_pm_free", "azref_pm_join",
                     "azref_pm_stop", "azref_cache_free", "azref_rng_free"):
            getattr(L, name).argtypes = [vp]
        L.azref_c4_play.argtypes = [vp, u32]
        L.azref_c4_valid.argtypes = [vp, vp]
        L.azref_c4_scores.argtypes = [vp, vp]
        L.azref_c4_canonical.argtypes = [vp, vp]
        L.azref_c4_player.argtypes = [vp]
        L.azref_c4_turn.argtypes = [vp]
        L.azref_c4_turn.restype = u32
        L.azref_c4_to_bytes.argtypes = [vp, vp]
        L.azref_c4_equal.argtypes = [vp, vp]
        L.azref_c4_hash.argtypes = [vp]
        L.azref_c4_hash.restype = C.c_uint64
        L.azref_c4_mirror.argtypes = [vp] * 7
        L.azref_mcts_new.argtypes = [C.POINTER(MctsCfg)]
        L.azref_mcts_find_leaf.argtypes = [vp, vp]
        L.azref_mcts_process_result.argtypes = [vp, vp, vp, u32, vp, u32, C.c_int]
        L.azref_mcts_update_root.argtypes = [vp, vp, u32]
        for name in ("azref_mcts_counts", "azref_mcts_root_q", "azref_mcts_root_value",
                     "azref_mcts_gumbel_improved_policy"):
            getattr(L, name).argtypes = [vp, vp]
        L.azref_mcts_probs.argtypes = [vp, C.c_float, vp]
        L.azref_mcts_probs_pruned.argtypes = [vp, C.c_float, vp]
        for name in ("azref_mcts_depth", "azref_mcts_root_n", "azref_mcts_gumbel_final_action"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = u32
        for name in ("azref_mcts_avg_leaf_depth", "azref_mcts_entropy"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = C.c_float
        L.azref_mcts_apply_root_policy_temp.argtypes = [vp]
        L.azref_mcts_add_root_noise.argtypes = [vp]
        L.azref_mcts_set_gumbel_num_sims.argtypes = [vp, u32]
        L.azref_pick_move.argtypes = [vp, u32, C.POINTER(u32)]
        L.azref_seed_thread_rng.argtypes = [C.c_uint64]
        L.azref_pm_new_connect4.argtypes = [C.POINTER(PlayCfg)]
        L.azref_pm_play_here.argtypes = [vp, C.c_uint64, C.c_int]
        L.azref_pm_start_workers.argtypes = [vp, u32, C.c_uint64, C.c_int]
        L.azref_pm_wait_quiescent.argtypes = [vp, u32]
        L.azref_pm_build_batch.argtypes = [vp, u32, u32, vp, vp]
        L.azref_pm_build_batch.restype = u32
        L.azref_pm_update_inferences.argtypes = [vp, u32, vp, u32, vp, u32, vp, u32]
        L.azref_pm_drain_history.argtypes = [vp, u32, vp, vp, vp]
        L.azref_pm_drain_history.restype = u32
        for name in ("azref_pm_hist_count", "azref_pm_games_completed", "azref_pm_remaining_games",
                     "azref_pm_awaiting_inference", "azref_pm_awaiting_mcts"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = u32
        for name in ("azref_pm_scores", "azref_pm_resign_scores", "azref_pm_metrics", "azref_pm_cache_stats"):
            getattr(L, name).argtypes = [vp, vp]
        L.azref_pm_perm_scores.argtypes = [vp, u32, vp]
        L.azref_pm_perm_scores.restype = u32
        L.azref_pm_progress_sims.argtypes = [vp, u32]
        L.azref_pm_progress_sims.restype = C.c_double
        L.azref_pm_game_state_bytes.argtypes = [vp, u32, vp]
        L.azref_pm_game_counts.argtypes = [vp, u32, u32, vp]
        L.azref_pm_game_root_q.argtypes = [vp, u32, u32, vp]
        L.azref_pm_game_root_value.argtypes = [vp, u32, u32, vp]
        L.azref_pm_game_depth.argtypes = [vp, u32, u32]
        L.azref_pm_game_depth.restype = u32
        L.azref_pm_game_root_n.argtypes = [vp, u32, u32]
        L.azref_pm_game_root_n.restype = u32
        L.azref_cache_new.argtypes = [u32, u32, u32, u32]
        L.azref_cache_find.argtypes = [vp, C.c_uint64, vp, vp]
        L.azref_cache_insert.argtypes = [vp, C.c_uint64, vp, vp]
        L.azref_cache_stats.argtypes = [vp, vp]
        L.azref_rng_new.argtypes = [C.c_uint64, C.c_int, C.c_uint64]
        L.azref_rng_u32.argtypes = [vp]
        L.azref_rng_u32.restype = u32
        L.azref_rng_shuffle.argtypes = [vp, u32, vp]
        L.azref_rng_uniform01.argtypes = [vp]
        L.azref_rng_uniform01.restype = C.c_float
        L.azref_rng_gamma.argtypes = [vp, C.c_float, u32, vp]
        L.azref_rng_gumbel.argtypes = [vp]
        L.azref_rng_gumbel.restype = C.c_float
        _lib = L
    return _lib


def play_cfg(**kw):
    c = PlayCfg(games_to_play=1, concurrent_games=1, max_batch_size=1, cache_shards=1, queue_shards=1, cpuct=2.0,
                start_temp=1.0, final_temp=1.0, tree_reuse=1, mcts_root_temp=1.0, playout_cap_depth=25,
                playout_cap_percent=0.75, gumbel_m=16, gumbel_c_visit=50.0, gumbel_c_scale=1.0)
    c.mcts_visits[0] = c.mcts_visits[1] = 100
    for k, v in kw.items():
        if k == "mcts_visits":
            c.mcts_visits[0], c.mcts_visits[1] = v
        else:
            if not hasattr(c, k):
                raise AttributeError(k)
            setattr(c, k, v)
    return c


def P(a):
    return a.ctypes.data_as(C.c_void_p)


class RefPlayManager:
    """The reference PlayManager (Connect4) behind the lock-step harness of SURVEY.md Appendix A."""

    def __init__(self, cfg):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.azref_pm_new_connect4(C.byref(cfg))
        if not self.h:
            raise RuntimeError(self.L.azref_last_error().decode())
        self.G = cfg.concurrent_games

    def close(self):
        if self.h:
            self.L.azref_pm_free(self.h)
            self.h = None

    def play_here(self, seed):
        if self.L.azref_pm_play_here(self.h, seed, 1) != 0:
            raise RuntimeError(self.L.azref_last_error().decode())

    def start_workers(self, n=1, seed=0, do_seed=True):
        self.L.azref_pm_start_workers(self.h, n, seed, 1 if do_seed else 0)

    def join(self):
        self.L.azref_pm_join(self.h)

    def wait_quiescent(self, timeout_ms=20000):
        return self.L.azref_pm_wait_quiescent(self.h, timeout_ms)

    def build_batch(self, group=0, max_rows=None):
        m = max_rows or self.G
        ids = np.empty(m, np.uint32)
        canon = np.empty((m, 4, 6, 7), np.float32)
        n = self.L.azref_pm_build_batch(self.h, group, m, P(ids), P(canon))
        return ids[:n], canon[:n]

    def update_inferences(self, ids, v, pi, group=0):
        ids = np.ascontiguousarray(ids, np.uint32)
        v = np.ascontiguousarray(v, np.float32)
        pi = np.ascontiguousarray(pi, np.float32)
        self.L.azref_pm_update_inferences(self.h, group, P(ids), len(ids), P(v), v.shape[1], P(pi), pi.shape[1])

    def drain_history(self, max_rows):
        canon = np.empty((max_rows, 4, 6, 7), np.float32)
        v = np.empty((max_rows, 3), np.float32)
        pi = np.empty((max_rows, 7), np.float32)
        n = self.L.azref_pm_drain_history(self.h, max_rows, P(canon), P(v), P(pi))
        return canon[:n], v[:n], pi[:n]

    def perm_scores(self, perm):
        """(perm_scores(perm) float32[3], perm_games_completed(perm))"""
        s = np.zeros(3, np.float32)
        n = self.L.azref_pm_perm_scores(self.h, perm, P(s))
        assert n != 0xFFFFFFFF, "perm out of range"
        return s, n

    def scores(self):
        s = np.zeros(3, np.float32)
        self.L.azref_pm_scores(self.h, P(s))
        return s

    def metrics(self):
        m = np.zeros(7, np.float32)
        self.L.azref_pm_metrics(self.h, P(m))
        return dict(zip(["avg_game_length", "avg_leaf_depth", "avg_search_entropy", "fast_avg_leaf_depth",
                         "fast_avg_search_entropy", "avg_moves_per_turn", "avg_valid_moves"], m.tolist()))

    def games_completed(self):
        return self.L.azref_pm_games_completed(self.h)

    def remaining_games(self):
        return self.L.azref_pm_remaining_games(self.h)

    def peek(self, game, seat):
        state = np.zeros(89, np.uint8)
        counts = np.zeros(7, np.uint32)
        q = np.zeros(7, np.float32)
        rv = np.zeros(3, np.float32)
        self.L.azref_pm_game_state_bytes(self.h, game, P(state))
        self.L.azref_pm_game_counts(self.h, game, seat, P(counts))
        self.L.azref_pm_game_root_q(self.h, game, seat, P(q))
        self.L.azref_pm_game_root_value(self.h, game, seat, P(rv))
        return dict(state=state, counts=counts, q=q, root_value=rv,
                    depth=self.L.azref_pm_game_depth(self.h, game, seat),
                    root_n=self.L.azref_pm_game_root_n(self.h, game, seat))
