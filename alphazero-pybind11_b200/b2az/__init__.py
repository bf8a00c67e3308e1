"""b2az — ctypes binding of the C ABI in include/b2az.h (libb2az.so, CUDA sm_100a).

This is the thin Python face used by tests/, bench.py and __graft_entry__.py. It mirrors the
reference's pybind11 surface for the self-play hot path (src/py_wrapper.cc:352-504:
PlayManager / PlayParams / build_batch / update_inferences / build_history_batch) on top of the
C ABI; see INTEGRATION.md for the pybind11 stub a maintainer of the reference would add instead.

There is no CPU fallback: `load()` raises if the CUDA library is missing and every call raises
B2azError when no CUDA device is usable. (`load(path)` with an explicit path exists so the CPU
test-suite can point the same binding at the host-emulation build of the engine logic.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(os.path.dirname(_HERE), "libb2az.so")

EVAL_NN, EVAL_RANDOM = 0, 1
RNG_PER_GAME, RNG_GLOBAL = 0, 1
STEP_DEFAULT, STEP_FLAT, STEP_WAVES, STEP_SYNC, STEP_QUEUE = 0, 1, 2, 3, 4
CANON_SHAPE = (4, 6, 7)
NUM_MOVES = 7
NUM_PLAYERS = 2


class B2azError(RuntimeError):
    """Mirrors the reference's std::runtime_error -> RuntimeError convention."""

    def __init__(self, code, msg):
        super().__init__(f"b2az error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("game", C.c_uint32), ("games_to_play", C.c_uint32), ("concurrent_games", C.c_uint32),
        ("max_batch_size", C.c_uint32), ("max_cache_size", C.c_uint32), ("mcts_visits", C.c_uint32 * 2),
        ("cpuct", C.c_float), ("start_temp", C.c_float), ("final_temp", C.c_float),
        ("temp_decay_half_life", C.c_float), ("history_enabled", C.c_uint8), ("self_play", C.c_uint8),
        ("tree_reuse", C.c_uint8), ("playout_cap_randomization", C.c_uint8), ("epsilon", C.c_float),
        ("mcts_root_temp", C.c_float), ("playout_cap_depth", C.c_uint32), ("playout_cap_percent", C.c_float),
        ("fpu_reduction", C.c_float), ("root_fpu_zero", C.c_uint8), ("shaped_dirichlet", C.c_uint8),
        ("policy_target_pruning", C.c_uint8), ("gumbel_enabled", C.c_uint8), ("resign_percent", C.c_float),
        ("resign_playthrough_percent", C.c_float), ("eval_type", C.c_uint8), ("rng_mode", C.c_uint8),
        ("per_slot_quota", C.c_uint8), ("pad1_", C.c_uint8), ("seed", C.c_uint64), ("pool_nodes", C.c_uint64),
        ("history_capacity", C.c_uint32), ("lanes_per_game", C.c_uint32), ("compact_pages", C.c_uint32),
        ("gumbel_m", C.c_uint32), ("gumbel_c_visit", C.c_float), ("gumbel_c_scale", C.c_float),
        ("gumbel_full", C.c_uint8), ("fast_search_uses_gumbel", C.c_uint8), ("model_groups", C.c_uint8 * 2),
        ("step_kernel", C.c_uint32), ("seat_cap_visits", C.c_uint32 * 2),
        ("n_seat_perms", C.c_uint32), ("seat_perms", (C.c_uint8 * 2) * 8), ("perm_seat_visits", (C.c_uint32 * 2) * 8),
        ("perm_seat_cap_visits", (C.c_uint32 * 2) * 8), ("group_random", C.c_uint8 * 2), ("pad5_", C.c_uint8 * 2),
    ]


class PermStats(C.Structure):  # b2az_perm_stats (include/b2az.h)
    _fields_ = [("scores", C.c_float * 3), ("games_completed", C.c_uint32), ("variant_scores", (C.c_float * 3) * 4),
                ("variant_games_completed", C.c_uint32 * 4)]


class Stats(C.Structure):
    _fields_ = [
        ("simulations", C.c_uint64), ("moves", C.c_uint64), ("games_completed", C.c_uint32),
        ("games_started", C.c_uint32), ("active_games", C.c_uint32), ("leaf_count", C.c_uint32),
        ("hist_count", C.c_uint32), ("scores", C.c_float * 3), ("resign_scores", C.c_float * 3),
        ("avg_game_length", C.c_float), ("avg_leaf_depth", C.c_float), ("avg_search_entropy", C.c_float),
        ("fast_avg_leaf_depth", C.c_float), ("fast_avg_search_entropy", C.c_float),
        ("avg_moves_per_turn", C.c_float), ("avg_valid_moves", C.c_float),
        ("cache_hits", C.c_uint64), ("cache_misses", C.c_uint64), ("cache_evictions", C.c_uint64),
        ("cache_reinserts", C.c_uint64), ("cache_size", C.c_uint64), ("cache_max_size", C.c_uint64),
        ("pool_pages_total", C.c_uint64), ("pool_pages_free", C.c_uint64), ("device_error", C.c_uint32),
        ("pad_", C.c_uint32), ("compactions", C.c_uint64),
        ("sum_game_length", C.c_uint64), ("total_move_count", C.c_uint64), ("full_move_count", C.c_uint64),
        ("fast_move_count", C.c_uint64), ("sum_leaf_depth", C.c_double), ("sum_search_entropy", C.c_double),
        ("fast_sum_leaf_depth", C.c_double), ("fast_sum_search_entropy", C.c_double), ("sum_valid_moves", C.c_double),
    ]


class ForestParams(C.Structure):  # b2az_forest_params (include/b2az.h)
    _fields_ = [("game", C.c_uint32), ("n_trees", C.c_uint32), ("max_turns", C.c_uint32), ("words_per_tree", C.c_uint32),
                ("cpuct", C.c_float), ("fpu_reduction", C.c_float), ("epsilon", C.c_float), ("root_policy_temp", C.c_float),
                ("root_fpu_zero", C.c_uint8), ("relative_values", C.c_uint8), ("gumbel_enabled", C.c_uint8),
                ("gumbel_full", C.c_uint8), ("gumbel_m", C.c_uint32), ("seed", C.c_uint64), ("gumbel_c_visit", C.c_float),
                ("gumbel_c_scale", C.c_float), ("shaped_dirichlet", C.c_uint8), ("debug_serial_shuffle", C.c_uint8),
                ("pad2_", C.c_uint8 * 2), ("max_in_flight", C.c_uint32)]


class TaflSelfplayParams(C.Structure):  # b2az_tafl_selfplay_params (include/b2az.h)
    _fields_ = [("forest", ForestParams), ("n_games", C.c_uint32), ("games_per_slot", C.c_uint32), ("visits", C.c_uint32),
                ("start_temp", C.c_float), ("final_temp", C.c_float), ("temp_decay_half_life", C.c_float),
                ("history_enabled", C.c_uint8), ("policy_target_pruning", C.c_uint8), ("tree_reuse", C.c_uint8),
                ("pad_", C.c_uint8), ("hist_capacity", C.c_uint32), ("seat_visits", C.c_uint32 * 2),
                ("seat_cap_visits", C.c_uint32 * 2), ("playout_cap_depth", C.c_uint32), ("playout_cap_percent", C.c_float),
                ("resign_percent", C.c_float), ("resign_playthrough_percent", C.c_float),
                ("playout_cap_randomization", C.c_uint8), ("fast_search_uses_gumbel", C.c_uint8), ("pad2_", C.c_uint8 * 2),
                ("n_variant_half_life", C.c_uint32), ("variant_half_life", C.c_float * 4), ("variant_probs", C.c_float * 4),
                ("cache_entries", C.c_uint32),
                ("n_seat_perms", C.c_uint32), ("seat_perms", (C.c_uint8 * 2) * 8), ("perm_seat_visits", (C.c_uint32 * 2) * 8),
                ("perm_seat_cap_visits", (C.c_uint32 * 2) * 8), ("group_random", C.c_uint8 * 2), ("has_seat_search", C.c_uint8),
                ("pad4_", C.c_uint8), ("seat_epsilon", (C.c_float * 2) * 8), ("seat_root_temp", (C.c_float * 2) * 8),
                ("seat_root_fpu_zero", (C.c_uint8 * 2) * 8), ("seat_gumbel_enabled", (C.c_uint8 * 2) * 8),
                ("seat_gumbel_full", (C.c_uint8 * 2) * 8), ("seat_gumbel_m", (C.c_uint32 * 2) * 8),
                ("seat_gumbel_c_visit", (C.c_float * 2) * 8), ("seat_gumbel_c_scale", (C.c_float * 2) * 8),
                ("seat_resign_threshold", (C.c_float * 2) * 8), ("seat_resign_consecutive", (C.c_uint32 * 2) * 8)]


SLOT_DTYPE = np.dtype([("active", "u4"), ("games_started", "u4"), ("games_completed", "u4"), ("pending", "u4"),
                       ("move_count", "u4"), ("full_move_count", "u4"), ("total_move_count", "u4"),
                       ("total_full_move_count", "u4"), ("game_length", "u4"), ("picked", "u4"), ("error", "u4"),
                       ("capped", "u4"), ("g_leaf_depth", "f8"), ("g_entropy", "f8"), ("g_valid_moves", "f8"),
                       ("leaf_depth", "f8"), ("entropy", "f8"), ("valid_moves", "f8"), ("simulations", "u8"),
                       ("scores", "f4", 3), ("playthrough", "u4"), ("fast_move_count", "u4"), ("total_fast_move_count", "u4"),
                       ("g_fast_leaf_depth", "f8"), ("g_fast_entropy", "f8"), ("fast_leaf_depth", "f8"), ("fast_entropy", "f8"),
                       ("resign_scores", "f4", 3), ("resign_streak", "u2", 2), ("coin_state", "u8"), ("coin_inc", "u8")])  # b2az_tafl_selfplay_slot
assert SLOT_DTYPE.itemsize == 192

_libs = {}


def load(path=None):
    """dlopen the engine. Without `path` this is the product library; it must exist."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise B2azError(-2, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                            " (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u32, u64p = C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)
    L.b2az_last_error.restype = C.c_char_p
    L.b2az_params_default.argtypes = [C.POINTER(Params)]
    L.b2az_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.b2az_destroy.argtypes = [vp]
    L.b2az_step.argtypes = [vp, u32, vp]
    L.b2az_leaf_batch.argtypes = [vp, vp, C.POINTER(u32), C.POINTER(vp), C.POINTER(vp)]
    L.b2az_leaf_batch_host.argtypes = [vp, vp, u32, vp, vp, C.POINTER(u32)]
    L.b2az_submit_eval.argtypes = [vp, vp, vp, u32]
    L.b2az_submit_eval_host.argtypes = [vp, vp, vp, vp, vp, u32]
    L.b2az_leaf_batch_device.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.b2az_submit_eval_all.argtypes = [vp, vp, vp]
    L.b2az_leaf_valid_device.argtypes = [vp, vp, C.POINTER(vp)]
    L.b2az_drain_history.argtypes = [vp, vp, u32, vp, vp, vp, C.c_int, C.POINTER(u32)]
    L.b2az_drain_history_sym.argtypes = [vp, vp, u32, vp, vp, vp, C.c_int, C.POINTER(u32)]
    L.b2az_get_stats.argtypes = [vp, vp, C.POINTER(Stats)]
    L.b2az_peek.argtypes = [vp, vp, u32, u32, vp, vp, vp, vp, C.POINTER(u32), C.POINTER(u32), vp]
    L.b2az_c4_batch.argtypes = [C.c_int, u32] + [vp] * 11
    L.b2az_leaf_seats_host.argtypes = [vp, vp, vp, u32]
    L.b2az_set_games_to_play.argtypes = [vp, u32]
    L.b2az_history_mark.argtypes = [vp, vp]
    L.b2az_drain_history_marked.argtypes = [vp, vp, u32, vp, vp, vp, C.c_int, C.POINTER(u32)]
    L.b2az_cache_insert_host.argtypes = [vp, vp, vp, vp, vp, u32]
    L.b2az_cache_find_host.argtypes = [vp, vp, vp, u32, vp, vp, vp]
    L.b2az_tafl_replay.argtypes = [C.c_int, u32, u32, u32, u32] + [vp] * 11
    L.b2az_tafl_replay_device.argtypes = [u32, u32, u32, u32] + [vp] * 10
    L.b2az_sg_replay.argtypes = [C.c_int, u32, u32, u32] + [vp] * 8
    L.b2az_sg_symmetries.argtypes = [C.c_int, u32, u32] + [vp] * 6 + [C.c_int, C.c_int, vp]
    L.b2az_sg_replay_device.argtypes = [u32, u32, u32, vp, vp, vp, u32] + [vp] * 7
    L.b2az_forest_create.argtypes = [C.POINTER(ForestParams), C.c_int, C.POINTER(vp)]
    L.b2az_forest_destroy.argtypes = [vp]
    L.b2az_forest_find_leaf.argtypes = [vp, vp, C.POINTER(vp)]
    L.b2az_forest_leaf_canon_host.argtypes = [vp, vp, vp]
    L.b2az_forest_process_result.argtypes = [vp, vp, vp, vp, C.c_int]
    L.b2az_forest_process_result_host.argtypes = [vp, vp, vp, vp, C.c_int]
    L.b2az_forest_simulate.argtypes = [vp, vp, u32, C.c_int]
    L.b2az_forest_root_noise.argtypes = [vp, vp, C.c_int]
    L.b2az_forest_find_leaf_batched.argtypes = [vp, vp, C.POINTER(vp)]
    L.b2az_forest_process_result_batched.argtypes = [vp, vp, u32, vp, vp, C.c_int, C.c_int]
    L.b2az_forest_simulate_batched.argtypes = [vp, vp, u32, u32]
    L.b2az_forest_reset_batch.argtypes = [vp, vp]
    L.b2az_forest_probs.argtypes = [vp, vp, C.c_float, C.c_int, C.c_int, vp, vp]
    L.b2az_forest_advance.argtypes = [vp, vp]
    L.b2az_forest_set_gumbel_num_sims.argtypes = [vp, vp, u32]
    L.b2az_forest_gumbel_result.argtypes = [vp, vp, vp, vp]
    L.b2az_forest_update_root.argtypes = [vp, vp, vp]
    L.b2az_forest_counts.argtypes = [vp, vp, vp, vp, vp]
    L.b2az_tafl_selfplay_variant_stats.argtypes = [vp, vp, vp]
    L.b2az_forest_principal_variation.argtypes = [vp, vp, u32, vp, vp]
    L.b2az_forest_leaf_path.argtypes = [vp, vp, C.c_int, vp, vp]
    L.b2az_forest_root_ops.argtypes = [vp, vp, C.c_int, C.c_int]
    L.b2az_forest_set_root.argtypes = [vp, u32, vp, u32, vp, u32]
    L.b2az_tafl_symmetries.argtypes = [C.c_int, u32, u32] + [vp] * 6
    L.b2az_tafl_positions.argtypes = [C.c_int, u32, u32, u32] + [vp] * 12
    L.b2az_tafl_selfplay_create.argtypes = [C.POINTER(TaflSelfplayParams), C.c_int, C.POINTER(vp)]
    L.b2az_tafl_selfplay_destroy.argtypes = [vp]
    L.b2az_tafl_selfplay_play.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.b2az_tafl_selfplay_find_leaf.argtypes = [vp, vp, C.POINTER(vp)]
    L.b2az_tafl_selfplay_process_result.argtypes = [vp, vp, vp, vp, C.c_int, C.POINTER(u32)]
    L.b2az_tafl_selfplay_drain_history.argtypes = [vp, vp, u32, vp, vp, vp, vp, C.POINTER(u32)]
    L.b2az_tafl_selfplay_slots.argtypes = [vp, vp, vp, vp]
    L.b2az_tafl_selfplay_get_stats.argtypes = [vp, vp, C.POINTER(Stats)]
    L.b2az_tafl_selfplay_leaf_batch_host.argtypes = [vp, vp, u32, vp, vp, C.POINTER(u32)]
    L.b2az_tafl_selfplay_submit_eval_host.argtypes = [vp, vp, vp, vp, vp, u32]
    L.b2az_tafl_selfplay_leaf_groups_host.argtypes = [vp, vp, u32]
    L.b2az_tafl_selfplay_perm_stats.argtypes = [vp, vp, vp, C.POINTER(u32)]
    L.b2az_leaf_groups_host.argtypes = [vp, vp, vp, u32]
    L.b2az_perm_scores.argtypes = [vp, vp, vp, C.POINTER(u32)]
    _libs[path] = L
    return L


def default_params(lib=None, **kw):
    L = lib or load()
    p = Params()
    L.b2az_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "mcts_visits":
            p.mcts_visits[0], p.mcts_visits[1] = v
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fill_perms(p, seat_perms=None, perm_seat_visits=None, perm_seat_cap_visits=None, group_random=None):
    """PlayParams::seat_perms and the per-permutation budgets into a Params / TaflSelfplayParams structure."""
    for i, perm in enumerate(seat_perms or []):
        p.seat_perms[i][0], p.seat_perms[i][1] = perm
        p.n_seat_perms = i + 1
    for i, row in enumerate(perm_seat_visits or []):
        p.perm_seat_visits[i][0], p.perm_seat_visits[i][1] = row
    for i, row in enumerate(perm_seat_cap_visits or []):
        p.perm_seat_cap_visits[i][0], p.perm_seat_cap_visits[i][1] = row
    for i, r in enumerate(group_random or []):
        p.group_random[i] = int(r)


def perm_stats_array(n=8):
    return (PermStats * n)()


class Engine:
    """One device-resident self-play pool (the reference's PlayManager for one model group)."""

    def __init__(self, params, device=0, lib=None):
        self.L = lib or load()
        self.params = params
        self.h = C.c_void_p()
        self._check(self.L.b2az_create(C.byref(params), device, C.byref(self.h)))
        self.G = params.concurrent_games

    def _check(self, rc):
        if rc != 0:
            raise B2azError(rc, self.L.b2az_last_error().decode())

    def close(self):
        if self.h:
            self.L.b2az_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- PlayManager::play() loop body over every slot
    def step(self, n_steps=1, stream=None):
        self._check(self.L.b2az_step(self.h, n_steps, stream))

    # -- build_batch (device pointers): returns (count, canon_ptr, ids_ptr)
    def leaf_batch(self, stream=None):
        n, cp, ip = C.c_uint32(), C.c_void_p(), C.c_void_p()
        self._check(self.L.b2az_leaf_batch(self.h, stream, C.byref(n), C.byref(cp), C.byref(ip)))
        return n.value, cp.value, ip.value

    # -- build_batch (host buffers): returns (ids uint32[n], canonical float32[n,4,6,7])
    def leaf_batch_host(self, max_rows=None, stream=None):
        m = max_rows or self.G
        canon = np.empty((m,) + CANON_SHAPE, np.float32)
        ids = np.empty(m, np.uint32)
        n = C.c_uint32()
        self._check(self.L.b2az_leaf_batch_host(self.h, stream, m, _ptr(canon), _ptr(ids), C.byref(n)))
        return ids[: n.value], canon[: n.value]

    # -- pointer flavours (caller-owned, e.g. pinned, host buffers): no allocation per call
    def leaf_batch_host_into(self, canon_ptr, ids_ptr, max_rows, stream=None):
        n = C.c_uint32()
        self._check(self.L.b2az_leaf_batch_host(self.h, stream, max_rows, canon_ptr, ids_ptr, C.byref(n)))
        return n.value

    def submit_eval_host_from(self, ids_ptr, v_ptr, pi_ptr, count, stream=None):
        self._check(self.L.b2az_submit_eval_host(self.h, stream, ids_ptr, v_ptr, pi_ptr, count))

    def drain_history_into(self, canon_ptr, v_ptr, pi_ptr, max_rows, stream=None):
        n = C.c_uint32()
        self._check(self.L.b2az_drain_history(self.h, stream, max_rows, canon_ptr, v_ptr, pi_ptr, 0, C.byref(n)))
        return n.value

    # -- the same without any host synchronisation: (canon_ptr, ids_ptr, count_ptr), all DEVICE pointers
    def leaf_batch_device(self, stream=None):
        cp, ip, np_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.L.b2az_leaf_batch_device(self.h, stream, C.byref(cp), C.byref(ip), C.byref(np_)))
        return cp.value, ip.value, np_.value

    def submit_eval_all(self, v_dev_ptr, pi_dev_ptr):
        self._check(self.L.b2az_submit_eval_all(self.h, v_dev_ptr, pi_dev_ptr))

    def submit_eval(self, v_dev_ptr, pi_dev_ptr, count):
        self._check(self.L.b2az_submit_eval(self.h, v_dev_ptr, pi_dev_ptr, count))

    # -- update_inferences(group, indices, v, pi)
    def submit_eval_host(self, ids, v, pi, stream=None):
        ids = np.ascontiguousarray(ids, np.uint32)
        v = np.ascontiguousarray(v, np.float32)
        pi = np.ascontiguousarray(pi, np.float32)
        assert v.shape == (len(ids), NUM_PLAYERS + 1) and pi.shape == (len(ids), NUM_MOVES)
        self._check(self.L.b2az_submit_eval_host(self.h, stream, _ptr(ids), _ptr(v), _ptr(pi), len(ids)))

    # -- build_history_batch
    def drain_history(self, max_rows, stream=None):
        canon = np.empty((max_rows,) + CANON_SHAPE, np.float32)
        v = np.empty((max_rows, NUM_PLAYERS + 1), np.float32)
        pi = np.empty((max_rows, NUM_MOVES), np.float32)
        n = C.c_uint32()
        self._check(self.L.b2az_drain_history(self.h, stream, max_rows, _ptr(canon), _ptr(v), _ptr(pi), 0, C.byref(n)))
        return canon[: n.value], v[: n.value], pi[: n.value]

    def drain_history_sym(self, max_samples, stream=None):
        """build_history_batch + exploit_symmetries: every sample followed by its mirror image (2 rows per sample)."""
        canon = np.empty((2 * max_samples,) + CANON_SHAPE, np.float32)
        v = np.empty((2 * max_samples, NUM_PLAYERS + 1), np.float32)
        pi = np.empty((2 * max_samples, NUM_MOVES), np.float32)
        n = C.c_uint32()
        self._check(self.L.b2az_drain_history_sym(self.h, stream, max_samples, _ptr(canon), _ptr(v), _ptr(pi), 0, C.byref(n)))
        return canon[: 2 * n.value], v[: 2 * n.value], pi[: 2 * n.value]

    def drain_history_device(self, max_rows, canon_ptr, v_ptr, pi_ptr, stream=None):
        n = C.c_uint32()
        self._check(self.L.b2az_drain_history(self.h, stream, max_rows, canon_ptr, v_ptr, pi_ptr, 1, C.byref(n)))
        return n.value

    def stats(self, stream=None):
        s = Stats()
        self._check(self.L.b2az_get_stats(self.h, stream, C.byref(s)))
        return s

    # -- overlapped drain: mark on the step stream, drain on a second stream while the next step runs
    def history_mark(self, stream=None):
        self._check(self.L.b2az_history_mark(self.h, stream))

    def drain_history_marked_into(self, canon_ptr, v_ptr, pi_ptr, max_rows, stream2=None, dst_is_device=0):
        n = C.c_uint32()
        self._check(self.L.b2az_drain_history_marked(self.h, stream2, max_rows, canon_ptr, v_ptr, pi_ptr, dst_is_device, C.byref(n)))
        return n.value

    def set_games_to_play(self, n):
        self._check(self.L.b2az_set_games_to_play(self.h, int(n)))
        self.params.games_to_play = int(n)

    # -- the searching seat of every row of the current leaf batch (several model groups)
    def leaf_seats_host(self, count, stream=None):
        seats = np.zeros(count, np.uint8)
        self._check(self.L.b2az_leaf_seats_host(self.h, stream, _ptr(seats), count))
        return seats

    # -- the position cache key by key, in order (S3FIFOCache::insert / ::find)

    def leaf_groups_host(self, count, stream=None):
        """Model group of the first `count` rows of the leaf batch (seat_perms[slot's permutation][searching seat])."""
        groups = np.empty(count, np.uint8)
        self._check(self.L.b2az_leaf_groups_host(self.h, stream, _ptr(groups), count))
        return groups

    def perm_scores(self, stream=None):
        """[(scores[3], games_completed)] per seat permutation (perm_scores / perm_games_completed)."""
        out, n = perm_stats_array(), C.c_uint32(0)
        self._check(self.L.b2az_perm_scores(self.h, stream, out, C.byref(n)))
        return [(np.array(out[i].scores[:], np.float32), int(out[i].games_completed)) for i in range(n.value)]

    def cache_insert(self, keys, v, pi, stream=None):
        keys = np.ascontiguousarray(keys, np.uint64)
        v = np.ascontiguousarray(v, np.float32).reshape(len(keys), NUM_PLAYERS + 1)
        pi = np.ascontiguousarray(pi, np.float32).reshape(len(keys), NUM_MOVES)
        self._check(self.L.b2az_cache_insert_host(self.h, stream, _ptr(keys), _ptr(v), _ptr(pi), len(keys)))

    def cache_find(self, keys, stream=None):
        keys = np.ascontiguousarray(keys, np.uint64)
        found = np.zeros(len(keys), np.uint8)
        v = np.zeros((len(keys), NUM_PLAYERS + 1), np.float32)
        pi = np.zeros((len(keys), NUM_MOVES), np.float32)
        self._check(self.L.b2az_cache_find_host(self.h, stream, _ptr(keys), len(keys), _ptr(found), _ptr(v), _ptr(pi)))
        return found.astype(bool), v, pi

    def peek(self, game, seat, stream=None):
        state = np.zeros(89, np.uint8)
        counts = np.zeros(7, np.uint32)
        q = np.zeros(7, np.float32)
        pol = np.zeros(7, np.float32)
        rv = np.zeros(3, np.float32)
        depth, root_n = C.c_uint32(), C.c_uint32()
        self._check(self.L.b2az_peek(self.h, stream, game, seat, _ptr(state), _ptr(counts), _ptr(q), _ptr(rv),
                                     C.byref(depth), C.byref(root_n), _ptr(pol)))
        return dict(state=state, counts=counts, q=q, policy=pol, root_value=rv, depth=depth.value,
                    root_n=root_n.value)


def c4_batch(boards, players, turns, moves=None, device=0, lib=None):
    """Batched Connect4 game kernels (valid_moves / play_move / scores / canonicalized)."""
    L = lib or load()
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 84)
    n = boards.shape[0]
    players = np.ascontiguousarray(players, np.uint8)
    turns = np.ascontiguousarray(turns, np.uint32)
    mv = None if moves is None else np.ascontiguousarray(moves, np.uint32)
    out = dict(boards=np.zeros((n, 2, 6, 7), np.int8), players=np.zeros(n, np.uint8),
               valid=np.zeros((n, 7), np.uint8), scores=np.zeros((n, 3), np.float32),
               terminal=np.zeros(n, np.uint8), canonical=np.zeros((n, 4, 6, 7), np.float32),
               status=np.zeros(n, np.int32))
    rc = L.b2az_c4_batch(device, n, _ptr(boards), _ptr(players), _ptr(turns), _ptr(mv), _ptr(out["boards"]),
                         _ptr(out["players"]), _ptr(out["valid"]), _ptr(out["scores"]), _ptr(out["terminal"]),
                         _ptr(out["canonical"]), _ptr(out["status"]))
    if rc != 0:
        raise B2azError(rc, L.b2az_last_error().decode())
    return out


TAFL_BRANDUBH, TAFL_OPENTAFL, TAFL_TAWLBWRDD = 0, 1, 2
TAFL_DIMS = {0: (7, 7), 1: (11, 8), 2: (11, 7)}  # game -> (board side S, canonical planes)


def game_dims(game):
    """(grid side, canonical planes, actions) of a game of the wide-tree search: B2AZ_TAFL_* or a Star Gambit id
    (10 + variant: the variant's own class, 20 + variant: the Unified 13x13 view)."""
    if game in TAFL_DIMS:
        S, P = TAFL_DIMS[game]
        return S, P, 2 * S ** 3
    if 10 <= game <= 13 or 20 <= game <= 24:  # 24: the Unified view with the variant mix (self-play engine)
        D = 13 if game >= 20 or game == 13 else 11
        return D, 36 if game >= 20 else 32, D * D * 10 + 19
    if game == 30:  # Connect4 under the wide-tree search API
        return 7, 4, 7
    raise B2azError(-1, f"unknown game {game}")


def tafl_replay(game, moves, lens, max_turns, want_valid=True, want_canonical=True, device=0, lib=None):
    """Tafl game kernels on a batch of transcripts: the position after every move of every game.
    moves uint16[n][max_len], lens[n]. Returns arrays shaped [n][max_len + 1][...] (rows beyond a game's length
    stay zero)."""
    L = lib or load()
    if game not in TAFL_DIMS:
        raise B2azError(-1, f"unknown tafl game {game}")
    S, P = TAFL_DIMS[game]
    moves = np.ascontiguousarray(moves, np.uint16)
    n, max_len = moves.shape
    lens = np.ascontiguousarray(lens, np.uint32)
    R = (n, max_len + 1)
    out = dict(boards=np.zeros(R + (3, S, S), np.int8), players=np.zeros(R, np.uint8), turns=np.zeros(R, np.uint32),
               reps=np.zeros(R, np.uint8), terminal=np.zeros(R, np.uint8), n_valid=np.zeros(R, np.uint32),
               valid=np.zeros(R + (2 * S ** 3,), np.uint8) if want_valid else None,
               canonical=np.zeros(R + (P, S, S), np.float32) if want_canonical else None,
               status=np.zeros(n, np.int32))
    rc = L.b2az_tafl_replay(device, game, n, max_len, max_turns, _ptr(moves), _ptr(lens), _ptr(out["boards"]),
                            _ptr(out["players"]), _ptr(out["turns"]), _ptr(out["reps"]), _ptr(out["terminal"]),
                            _ptr(out["n_valid"]), _ptr(out["valid"]), _ptr(out["canonical"]), _ptr(out["status"]))
    if rc != 0:
        raise B2azError(rc, L.b2az_last_error().decode())
    return out


SG_STATE_BYTES = 200


def sg_dims(game):
    """Star Gambit game id (10 + variant: the variant's own class; 20 + variant: the Unified 13x13 view) ->
    (grid side D, actions, canonical planes)."""
    if not (10 <= game <= 13 or 20 <= game <= 23):
        raise B2azError(-1, f"unknown Star Gambit game {game}")
    D = 13 if game >= 20 or game == 13 else 11
    return D, D * D * 10 + 19, 36 if game >= 20 else 32


def sg_replay(game, moves, lens, want_valid=True, want_canonical=True, device=0, lib=None):
    """Star Gambit game kernels on a batch of transcripts: the position after every move of every game.
    moves uint16[n][max_len], lens[n]. Returns arrays shaped [n][max_len + 1][...]; `states` rows are the device's
    SGState records (units 9 B each in the reference's field order, then n_units, reserves, player, acted, over,
    winner, variant, pad, turn)."""
    L = lib or load()
    D, A, P = sg_dims(game)
    moves = np.ascontiguousarray(moves, np.uint16)
    n, max_len = moves.shape
    lens = np.ascontiguousarray(lens, np.uint32)
    R = (n, max_len + 1)
    out = dict(states=np.zeros(R + (SG_STATE_BYTES,), np.uint8), terminal=np.zeros(R, np.uint8),
               n_valid=np.zeros(R, np.uint32), valid=np.zeros(R + (A,), np.uint8) if want_valid else None,
               canonical=np.zeros(R + (P, D, D), np.float32) if want_canonical else None, status=np.zeros(n, np.int32))
    rc = L.b2az_sg_replay(device, game, n, max_len, _ptr(moves), _ptr(lens), _ptr(out["states"]), _ptr(out["terminal"]),
                          _ptr(out["n_valid"]), _ptr(out["valid"]), _ptr(out["canonical"]), _ptr(out["status"]))
    if rc != 0:
        raise B2azError(rc, L.b2az_last_error().decode())
    return out


def sg_symmetries(game, canon, v, pi, fp16=False, device=0, lib=None):
    """Every Star Gambit sample followed by its NW-axis mirror image: ([n,2,P,D,D], [n,2,3], [n,2,A]), float32 or float16."""
    L = lib or load()
    D, A, P = sg_dims(game if game != 24 else 23)
    canon = np.ascontiguousarray(canon, np.float32).reshape(-1, P, D, D)
    n = canon.shape[0]
    v = np.ascontiguousarray(v, np.float32).reshape(n, 3)
    pi = np.ascontiguousarray(pi, np.float32).reshape(n, A)
    dt = np.float16 if fp16 else np.float32
    co, vo, po = np.zeros((n, 2, P, D, D), dt), np.zeros((n, 2, 3), dt), np.zeros((n, 2, A), dt)
    rc = L.b2az_sg_symmetries(device, game, n, _ptr(canon), _ptr(v), _ptr(pi), _ptr(co), _ptr(vo), _ptr(po), 0, int(fp16), None)
    if rc != 0:
        raise B2azError(rc, L.b2az_last_error().decode())
    return co, vo, po


def tafl_positions(game, boards, players, turns, reps, max_turns, moves=None, device=0, lib=None):
    """Tafl game kernels on arbitrary positions: scores / valid_moves / canonicalized of each position and,
    with `moves`, the board after play_move (no repetition bookkeeping) + whether anything was captured."""
    L = lib or load()
    S, P = TAFL_DIMS[game]
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 3, S, S)
    n = boards.shape[0]
    players = np.ascontiguousarray(players, np.uint8)
    turns = np.ascontiguousarray(turns, np.uint32)
    reps = np.ascontiguousarray(reps, np.uint8)
    mv = None if moves is None else np.ascontiguousarray(moves, np.uint32)
    out = dict(terminal=np.zeros(n, np.uint8), n_valid=np.zeros(n, np.uint32), valid=np.zeros((n, 2 * S ** 3), np.uint8),
               canonical=np.zeros((n, P, S, S), np.float32), boards_out=np.zeros((n, 3, S, S), np.int8),
               captured_any=np.zeros(n, np.uint8), status=np.zeros(n, np.int32))
    rc = L.b2az_tafl_positions(device, game, n, max_turns, _ptr(boards), _ptr(players), _ptr(turns), _ptr(reps), _ptr(mv),
                               _ptr(out["terminal"]), _ptr(out["n_valid"]), _ptr(out["valid"]), _ptr(out["canonical"]),
                               _ptr(out["boards_out"]), _ptr(out["captured_any"]), _ptr(out["status"]))
    if rc != 0:
        raise B2azError(rc, L.b2az_last_error().decode())
    return out


class Forest:
    """n_trees device-resident single-tree searches over a tafl game: the reference's `MCTS` class, batched
    (find_leaf / process_result / update_root / counts)."""
    INFO = ("depth", "root_n", "root_k", "root_term", "root_player", "turn", "rep", "error", "words_used", "root_v_bits",
            "total_leaf_depth", "player", "root_value_w_bits", "root_value_l_bits", "root_value_d_bits", "in_flight")

    def __init__(self, game, n_trees, max_turns, cpuct=1.25, fpu_reduction=0.25, root_fpu_zero=False, seed=0,
                 words_per_tree=0, epsilon=0.0, root_policy_temp=1.0, gumbel_m=0, gumbel_c_visit=50.0, gumbel_c_scale=1.0,
                 gumbel_full=False, shaped_dirichlet=False, serial_shuffle=False, max_in_flight=0, device=0, lib=None,
                 relative_values=None):
        self.L = lib or load()
        self.game, self.n = game, n_trees
        self.S, self.P, self.A = game_dims(game)
        if relative_values is None:  # GameState::relative_values(): true for the Star Gambit classes
            relative_values = game >= 10
        p = ForestParams(game=game, relative_values=int(relative_values), n_trees=n_trees, max_turns=max_turns, words_per_tree=words_per_tree, cpuct=cpuct,
                         fpu_reduction=fpu_reduction, epsilon=epsilon, root_policy_temp=root_policy_temp,
                         root_fpu_zero=int(root_fpu_zero), seed=seed, gumbel_enabled=int(gumbel_m > 0), gumbel_m=gumbel_m,
                         gumbel_c_visit=gumbel_c_visit, gumbel_c_scale=gumbel_c_scale, gumbel_full=int(gumbel_full),
                         shaped_dirichlet=int(shaped_dirichlet), debug_serial_shuffle=int(serial_shuffle),
                         max_in_flight=max_in_flight)
        self.slots = max(1, max_in_flight)
        self.h = C.c_void_p()
        self._check(self.L.b2az_forest_create(C.byref(p), device, C.byref(self.h)))

    def _check(self, rc):
        if rc != 0:
            raise B2azError(rc, self.L.b2az_last_error().decode())

    def close(self):
        if self.h:
            self.L.b2az_forest_destroy(self.h)
            self.h = None

    def find_leaf(self, stream=None):
        ptr = C.c_void_p()
        self._check(self.L.b2az_forest_find_leaf(self.h, stream, C.byref(ptr)))
        return ptr.value

    def leaf_canon(self, stream=None):
        """Leaf canonical planes: [n_trees][P][S][S], or [max_in_flight][n_trees][P][S][S] for a WU-UCT forest."""
        out = np.zeros((self.slots, self.n, self.P, self.S, self.S), np.float32)
        self._check(self.L.b2az_forest_leaf_canon_host(self.h, stream, _ptr(out)))
        return out if self.slots > 1 else out[0]

    # -- WU-UCT (find_leaf_batched / process_result_batched / reset_batch)
    def find_leaf_batched(self, stream=None):
        ptr = C.c_void_p()
        self._check(self.L.b2az_forest_find_leaf_batched(self.h, stream, C.byref(ptr)))
        return ptr.value

    def process_result_batched(self, leaf_index, v, pi, root_noise=False, stream=None):
        v = np.ascontiguousarray(v, np.float32)
        pi = np.ascontiguousarray(pi, np.float32)
        assert v.shape == (self.n, 3) and pi.shape == (self.n, self.A)
        self._check(self.L.b2az_forest_process_result_batched(self.h, stream, leaf_index, _ptr(v), _ptr(pi), int(root_noise), 1))

    def simulate_batched(self, n_rounds, width, stream=None):
        self._check(self.L.b2az_forest_simulate_batched(self.h, stream, n_rounds, width))

    def reset_batch(self, stream=None):
        self._check(self.L.b2az_forest_reset_batch(self.h, stream))

    def process_result(self, v, pi, root_noise=False, stream=None):
        v = np.ascontiguousarray(v, np.float32)
        pi = np.ascontiguousarray(pi, np.float32)
        assert v.shape == (self.n, 3) and pi.shape == (self.n, self.A)
        self._check(self.L.b2az_forest_process_result_host(self.h, stream, _ptr(v), _ptr(pi), int(root_noise)))

    def process_result_device(self, v_ptr, pi_ptr, root_noise=False, stream=None):
        self._check(self.L.b2az_forest_process_result(self.h, stream, v_ptr, pi_ptr, int(root_noise)))

    def simulate(self, n_sims, stream=None, root_noise=False):
        self._check(self.L.b2az_forest_simulate(self.h, stream, n_sims, int(root_noise)))

    def root_noise(self, add_noise=True, stream=None):
        """apply_root_policy_temp + add_root_noise on the reused roots (PlayManager after a move)."""
        self._check(self.L.b2az_forest_root_noise(self.h, stream, int(add_noise)))

    def set_gumbel_num_sims(self, n, stream=None):
        self._check(self.L.b2az_forest_set_gumbel_num_sims(self.h, stream, n))

    def gumbel_result(self, stream=None):
        """(gumbel_final_action per tree, gumbel_improved_policy [n_trees][A])"""
        action = np.zeros(self.n, np.uint32)
        policy = np.zeros((self.n, self.A), np.float32)
        self._check(self.L.b2az_forest_gumbel_result(self.h, stream, _ptr(action), _ptr(policy)))
        return action, policy

    def probs(self, temp, pick=False, pruned=False, stream=None):
        """MCTS::probs(temp) (or probs_pruned) per tree; with pick=True also pick_move(probs) (one RNG draw per tree)."""
        probs = np.zeros((self.n, self.A), np.float32)
        moves = np.zeros(self.n, np.uint32)
        self._check(self.L.b2az_forest_probs(self.h, stream, temp, int(pruned), int(pick), _ptr(probs), _ptr(moves)))
        return (probs, moves) if pick else probs

    def advance(self, stream=None):
        self._check(self.L.b2az_forest_advance(self.h, stream))

    def update_root(self, moves, stream=None):
        moves = np.ascontiguousarray(moves, np.uint32)
        assert moves.shape == (self.n,)
        self._check(self.L.b2az_forest_update_root(self.h, stream, _ptr(moves)))

    def counts(self, stream=None, want_q=True):
        counts = np.zeros((self.n, self.A), np.uint32)
        q = np.zeros((self.n, self.A), np.float32) if want_q else None
        info = np.zeros((self.n, 16), np.uint32)
        self._check(self.L.b2az_forest_counts(self.h, stream, _ptr(counts), _ptr(q), _ptr(info)))
        return counts, q, {k: info[:, i] for i, k in enumerate(self.INFO)}


def tafl_symmetries(game, canon, v, pi, device=0, lib=None):
    """The eight symmetric images of each sample (tafl_helper::eightSym order): ([n,8,P,S,S], [n,8,3], [n,8,A])."""
    L = lib or load()
    S, P = TAFL_DIMS[game]
    A = 2 * S ** 3
    canon = np.ascontiguousarray(canon, np.float32).reshape(-1, P, S, S)
    n = canon.shape[0]
    v = np.ascontiguousarray(v, np.float32).reshape(n, 3)
    pi = np.ascontiguousarray(pi, np.float32).reshape(n, A)
    co, vo, po = np.zeros((n, 8, P, S, S), np.float32), np.zeros((n, 8, 3), np.float32), np.zeros((n, 8, A), np.float32)
    rc = L.b2az_tafl_symmetries(device, game, n, _ptr(canon), _ptr(v), _ptr(pi), _ptr(co), _ptr(vo), _ptr(po))
    if rc != 0:
        raise B2azError(rc, L.b2az_last_error().decode())
    return co, vo, po


class TaflSelfplay:
    """PlayManager::play over a tafl game on the device (b2az_tafl_selfplay_*): n_games slots, each playing
    games_per_slot games with the two seats' trees and one generator — slot g == the reference PlayManager with
    concurrent_games = 1 run after MCTS::seed_thread_rng(seed + g)."""

    def __init__(self, game, n_games, max_turns, visits, games_per_slot=1, cpuct=1.25, fpu_reduction=0.25, root_fpu_zero=False,
                 seed=0, words_per_tree=0, epsilon=0.0, root_policy_temp=1.0, shaped_dirichlet=False, gumbel_m=0,
                 gumbel_c_visit=50.0, gumbel_c_scale=1.0, start_temp=1.0, final_temp=1.0, temp_decay_half_life=0.0,
                 history_enabled=True, policy_target_pruning=False, tree_reuse=True, hist_capacity=0, device=0, lib=None,
                 seat_visits=None, seat_cap_visits=None, playout_cap_randomization=False, playout_cap_depth=25,
                 playout_cap_percent=0.75, fast_search_uses_gumbel=False, resign_percent=0.0, resign_playthrough_percent=0.0,
                 temp_decay_half_life_by_variant=None, variant_probs=None, cache_entries=0, gumbel_full=False,
                 seat_perms=None, perm_seat_visits=None, perm_seat_cap_visits=None, group_random=None, seat_search=None):
        self.L = lib or load()
        self.game, self.n = game, n_games
        self.S, self.P, self.A = game_dims(game)
        fp = ForestParams(game=game, relative_values=int(10 <= game <= 24), gumbel_full=int(gumbel_full), n_trees=2 * n_games, max_turns=max_turns, words_per_tree=words_per_tree, cpuct=cpuct,
                          fpu_reduction=fpu_reduction, epsilon=epsilon, root_policy_temp=root_policy_temp,
                          root_fpu_zero=int(root_fpu_zero), seed=seed, gumbel_enabled=int(gumbel_m > 0), gumbel_m=gumbel_m,
                          gumbel_c_visit=gumbel_c_visit, gumbel_c_scale=gumbel_c_scale, shaped_dirichlet=int(shaped_dirichlet))
        p = TaflSelfplayParams(forest=fp, n_games=n_games, games_per_slot=games_per_slot, visits=visits, start_temp=start_temp,
                               final_temp=final_temp, temp_decay_half_life=temp_decay_half_life,
                               history_enabled=int(history_enabled), policy_target_pruning=int(policy_target_pruning),
                               tree_reuse=int(tree_reuse), hist_capacity=hist_capacity, playout_cap_depth=playout_cap_depth,
                               playout_cap_percent=playout_cap_percent, resign_percent=resign_percent,
                               resign_playthrough_percent=resign_playthrough_percent,
                               playout_cap_randomization=int(playout_cap_randomization),
                               fast_search_uses_gumbel=int(fast_search_uses_gumbel))
        p.cache_entries = cache_entries
        for i, w in enumerate((variant_probs or [])[:4]):
            p.variant_probs[i] = w
        for i, hl in enumerate((temp_decay_half_life_by_variant or [])[:4]):
            p.variant_half_life[i] = hl
            p.n_variant_half_life = i + 1
        for seat in range(2):
            p.seat_visits[seat] = (seat_visits or (0, 0))[seat]
            p.seat_cap_visits[seat] = (seat_cap_visits or (0, 0))[seat]
        fill_perms(p, seat_perms, perm_seat_visits, perm_seat_cap_visits, group_random)
        if seat_search:  # {field: [[seat 0, seat 1] per permutation]} for every seat_* field of the structure
            p.has_seat_search = 1
            defaults = dict(seat_epsilon=epsilon, seat_root_temp=root_policy_temp, seat_root_fpu_zero=int(root_fpu_zero),
                            seat_gumbel_enabled=int(gumbel_m > 0), seat_gumbel_full=int(gumbel_full), seat_gumbel_m=gumbel_m or 16,
                            seat_gumbel_c_visit=gumbel_c_visit, seat_gumbel_c_scale=gumbel_c_scale, seat_resign_threshold=-2.0,
                            seat_resign_consecutive=1)
            for name, dflt in defaults.items():
                rows = seat_search.get(name)
                for pm in range(max(1, len(seat_perms or []))):
                    for seat in range(2):
                        getattr(p, name)[pm][seat] = rows[pm][seat] if rows is not None else dflt
            assert not set(seat_search) - set(defaults), set(seat_search) - set(defaults)
        self.hist_capacity = hist_capacity or n_games * max_turns
        self.h = C.c_void_p()
        self._check(self.L.b2az_tafl_selfplay_create(C.byref(p), device, C.byref(self.h)))

    def _check(self, rc):
        if rc != 0:
            raise B2azError(rc, self.L.b2az_last_error().decode())

    def close(self):
        if self.h:
            self.L.b2az_tafl_selfplay_destroy(self.h)
            self.h = None

    def play(self, n_moves, stream=None, want_active=True):
        """n_moves x (search + move) for every active slot; returns the number of slots still cycling."""
        act = C.c_uint32(0)
        self._check(self.L.b2az_tafl_selfplay_play(self.h, stream, n_moves, C.byref(act) if want_active else None))
        return act.value if want_active else None

    def find_leaf(self, stream=None):
        ptr = C.c_void_p()
        self._check(self.L.b2az_tafl_selfplay_find_leaf(self.h, stream, C.byref(ptr)))
        return ptr.value

    def process_result(self, v, pi, host=True, stream=None, want_active=False):
        act = C.c_uint32(0)
        if host:
            v, pi = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(pi, np.float32)
            assert v.shape == (self.n, 3) and pi.shape == (self.n, self.A)
            v, pi = _ptr(v), _ptr(pi)
        self._check(self.L.b2az_tafl_selfplay_process_result(self.h, stream, v, pi, int(host),
                                                             C.byref(act) if want_active else None))
        return act.value if want_active else None

    def drain_history(self, stream=None, out=None):
        """Waiting samples, oldest first: (canonical, v, pi, slot). `out` = preallocated (canonical[cap,P,S,S], v[cap,3],
        pi[cap,A], slot[cap]) arrays to fill (e.g. views of pinned memory); else arrays sized for what is waiting."""
        if out is None:
            cap = max(1, int(self.stats(stream).hist_count))
            out = (np.empty((cap, self.P, self.S, self.S), np.float32), np.empty((cap, 3), np.float32),
                   np.empty((cap, self.A), np.float32), np.empty(cap, np.uint32))
        canon, v, pi, slot = out
        n = C.c_uint32(0)
        self._check(self.L.b2az_tafl_selfplay_drain_history(self.h, stream, len(slot), _ptr(canon), _ptr(v), _ptr(pi), _ptr(slot),
                                                            C.byref(n)))
        return canon[:n.value], v[:n.value], pi[:n.value], slot[:n.value]

    def leaf_batch_host(self, stream=None):
        """build_batch with host buffers: (slot ids uint32[n], canonical float32[n,P,S,S]) of the slots whose leaf waits."""
        canon = np.empty((self.n, self.P, self.S, self.S), np.float32)
        ids = np.empty(self.n, np.uint32)
        n = C.c_uint32(0)
        self._check(self.L.b2az_tafl_selfplay_leaf_batch_host(self.h, stream, self.n, _ptr(canon), _ptr(ids), C.byref(n)))
        return ids[:n.value], canon[:n.value]

    def submit_eval_host(self, ids, v, pi, stream=None):
        """update_inferences with host buffers (row i for slot ids[i]); plays the move of every complete search."""
        ids = np.ascontiguousarray(ids, np.uint32)
        v, pi = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(pi, np.float32)
        assert v.shape == (len(ids), 3) and pi.shape == (len(ids), self.A)
        self._check(self.L.b2az_tafl_selfplay_submit_eval_host(self.h, stream, _ptr(ids), _ptr(v), _ptr(pi), len(ids)))

    def leaf_groups_host(self, n):
        """Model group of every row of the last leaf_batch_host."""
        groups = np.empty(n, np.uint8)
        self._check(self.L.b2az_tafl_selfplay_leaf_groups_host(self.h, _ptr(groups), n))
        return groups

    def perm_stats(self, stream=None):
        """Per seat permutation: dict(scores[3], games_completed, variant_scores[4][3], variant_games_completed[4])."""
        out, n = perm_stats_array(), C.c_uint32(0)
        self._check(self.L.b2az_tafl_selfplay_perm_stats(self.h, stream, out, C.byref(n)))
        return [dict(scores=np.array(out[i].scores[:], np.float32), games_completed=int(out[i].games_completed),
                     variant_scores=np.array([list(r) for r in out[i].variant_scores], np.float32),
                     variant_games_completed=list(out[i].variant_games_completed)) for i in range(n.value)]

    def stats(self, stream=None):
        st = Stats()
        self._check(self.L.b2az_tafl_selfplay_get_stats(self.h, stream, C.byref(st)))
        return st

    VARIANT_DTYPE = np.dtype([("scores", "f4", 3), ("games_completed", "u4"), ("game_length", "u4"), ("total_move_count", "u4"),
                              ("full_move_count", "u4"), ("fast_move_count", "u4"), ("leaf_depth", "f8"), ("entropy", "f8"),
                              ("valid_moves", "f8"), ("fast_leaf_depth", "f8"), ("fast_entropy", "f8")])  # b2az_variant_stats

    def variant_stats(self, stream=None):
        """PlayManager's per-variant tables (StarGambitUnifiedGS): a record array of 4, sums over completed games."""
        out = np.zeros(4, self.VARIANT_DTYPE)
        self._check(self.L.b2az_tafl_selfplay_variant_stats(self.h, stream, _ptr(out)))
        return out

    def slots(self, stream=None):
        out = np.zeros(self.n, SLOT_DTYPE)
        err = np.zeros(2 * self.n, np.uint32)
        self._check(self.L.b2az_tafl_selfplay_slots(self.h, stream, _ptr(out), _ptr(err)))
        return out, err
