"""ctypes face of oracle/_ref/libazref_tafl.so — the UNMODIFIED reference tafl games (brandubh / opentafl /
tawlbwrdd) compiled against oracle/shim (oracle/ref_tafl_driver.cc). Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libazref_tafl.so")
BRANDUBH, OPENTAFL, TAWLBWRDD = 0, 1, 2
_lib = None


def available():
    return os.path.exists(REF_LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB)
        vp, u32 = C.c_void_p, C.c_uint32
        L.azref_tafl_last_error.restype = C.c_char_p
        L.azref_tafl_dims.argtypes = [C.c_int, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.azref_tafl_random_game.argtypes = [C.c_int, C.c_uint16, C.c_uint64, u32, vp]
        L.azref_tafl_random_game.restype = u32
        L.azref_tafl_replay.argtypes = [C.c_int, C.c_uint16, vp, u32] + [vp] * 9
        L.azref_tafl_search.argtypes = [C.c_int, C.c_uint16, C.c_uint64, C.c_float, C.c_float, C.c_int, u32, u32, C.c_int, vp, vp, vp, vp, vp, vp, u32, C.c_float, C.c_float, vp, C.c_float, C.c_float, C.c_int, u32, C.c_float, vp, C.c_float]
        L.azref_tafl_symmetries.argtypes = [C.c_int] + [vp] * 6
        L.azref_tafl_selfplay.argtypes = [C.c_int, C.c_uint16, C.c_uint64, C.POINTER(SelfplayCfg), u32] + [vp] * 7
        L.azref_tafl_position.argtypes = [C.c_int, vp, C.c_int8, C.c_uint16, C.c_uint16, C.c_uint8, u32, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


class SelfplayCfg(C.Structure):  # AzRefTaflSpCfg (oracle/ref_tafl_driver.cc)
    _fields_ = [("games_to_play", C.c_uint32), ("visits", C.c_uint32), ("cpuct", C.c_float), ("fpu_reduction", C.c_float),
                ("epsilon", C.c_float), ("mcts_root_temp", C.c_float), ("start_temp", C.c_float), ("final_temp", C.c_float),
                ("temp_decay_half_life", C.c_float), ("gumbel_m", C.c_uint32), ("gumbel_c_visit", C.c_float),
                ("gumbel_c_scale", C.c_float), ("root_fpu_zero", C.c_uint8), ("shaped_dirichlet", C.c_uint8),
                ("policy_target_pruning", C.c_uint8), ("gumbel_enabled", C.c_uint8), ("tree_reuse", C.c_uint8),
                ("history_enabled", C.c_uint8), ("pad_", C.c_uint8 * 2), ("seat_visits", C.c_uint32 * 2),
                ("seat_cap_visits", C.c_uint32 * 2), ("playout_cap_depth", C.c_uint32), ("playout_cap_percent", C.c_float),
                ("resign_percent", C.c_float), ("resign_playthrough_percent", C.c_float),
                ("playout_cap_randomization", C.c_uint8), ("fast_search_uses_gumbel", C.c_uint8), ("pad2_", C.c_uint8 * 2),
                ("has_perm", C.c_uint8), ("seat_perm", C.c_uint8 * 2), ("has_seat", C.c_uint8), ("group_visits", C.c_uint32 * 2),
                ("seat_epsilon", C.c_float * 2), ("seat_root_temp", C.c_float * 2), ("seat_root_fpu_zero", C.c_uint8 * 2),
                ("seat_gumbel_enabled", C.c_uint8 * 2), ("seat_gumbel_m", C.c_uint32 * 2), ("seat_gumbel_c_visit", C.c_float * 2),
                ("seat_gumbel_c_scale", C.c_float * 2), ("seat_resign_threshold", C.c_float * 2),
                ("seat_resign_consecutive", C.c_uint32 * 2)]


def selfplay(game, seed, max_turns, games_to_play, visits, cpuct=1.25, fpu_reduction=0.25, root_fpu_zero=False, epsilon=0.0,
             root_policy_temp=1.0, shaped_dirichlet=False, policy_target_pruning=False, gumbel_m=0, gumbel_c_visit=50.0,
             gumbel_c_scale=1.0, start_temp=1.0, final_temp=1.0, temp_decay_half_life=0.0, tree_reuse=True,
             seat_visits=None, seat_cap_visits=None, playout_cap_randomization=False, playout_cap_depth=25,
             playout_cap_percent=0.75, fast_search_uses_gumbel=False, resign_percent=0.0, resign_playthrough_percent=0.0,
             seat_perm=None, group_visits=None, seat_cfg=None):
    """The unmodified PlayManager, one slot, games_to_play games one after the other (EvalType::RANDOM) after
    MCTS::seed_thread_rng(seed). Returns dict(canonical, v, pi in history_ order, scores, games_completed,
    avg_game_length, avg_leaf_depth, avg_valid_moves, avg_search_entropy)."""
    S, A, P = dims(game)
    cap = games_to_play * max_turns
    cfg = SelfplayCfg(games_to_play=games_to_play, visits=visits, cpuct=cpuct, fpu_reduction=fpu_reduction, epsilon=epsilon,
                      mcts_root_temp=root_policy_temp, start_temp=start_temp, final_temp=final_temp,
                      temp_decay_half_life=temp_decay_half_life, gumbel_m=gumbel_m or 16, gumbel_c_visit=gumbel_c_visit,
                      gumbel_c_scale=gumbel_c_scale, root_fpu_zero=int(root_fpu_zero), shaped_dirichlet=int(shaped_dirichlet),
                      policy_target_pruning=int(policy_target_pruning), gumbel_enabled=int(gumbel_m > 0),
                      tree_reuse=int(tree_reuse), history_enabled=1, playout_cap_depth=playout_cap_depth,
                      playout_cap_percent=playout_cap_percent, resign_percent=resign_percent,
                      resign_playthrough_percent=resign_playthrough_percent,
                      playout_cap_randomization=int(playout_cap_randomization),
                      fast_search_uses_gumbel=int(fast_search_uses_gumbel))
    for seat in range(2):
        cfg.seat_visits[seat] = (seat_visits or (0, 0))[seat]
        cfg.seat_cap_visits[seat] = (seat_cap_visits or (0, 0))[seat]
    if seat_perm is not None:  # model_groups = [0, 1], seat_perms = [seat_perm], mcts_visits = group_visits (per group)
        cfg.has_perm = 1
        cfg.seat_perm[0], cfg.seat_perm[1] = seat_perm
        cfg.group_visits[0], cfg.group_visits[1] = group_visits
    if seat_cfg is not None:  # per-seat settings {field: (seat 0, seat 1)}; missing fields take the globals
        cfg.has_seat = 1
        dflt = dict(seat_epsilon=epsilon, seat_root_temp=root_policy_temp, seat_root_fpu_zero=int(root_fpu_zero),
                    seat_gumbel_enabled=int(gumbel_m > 0), seat_gumbel_m=gumbel_m or 16, seat_gumbel_c_visit=gumbel_c_visit,
                    seat_gumbel_c_scale=gumbel_c_scale, seat_resign_threshold=-2.0, seat_resign_consecutive=1)
        assert not set(seat_cfg) - set(dflt)
        for name, d in dflt.items():
            for i in range(2):
                getattr(cfg, name)[i] = seat_cfg[name][i] if name in seat_cfg else d
    canon = np.zeros((cap, P, S, S), np.float32)
    v, pi = np.zeros((cap, 3), np.float32), np.zeros((cap, A), np.float32)
    n, done = C.c_uint32(0), C.c_uint32(0)
    scores, metrics = np.zeros(3, np.float32), np.zeros(10, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().azref_tafl_selfplay(game, max_turns, seed, C.byref(cfg), cap, p(canon), p(v), p(pi),
                                   C.cast(C.byref(n), C.c_void_p), p(scores), C.cast(C.byref(done), C.c_void_p), p(metrics))
    if rc != 0:
        raise RuntimeError(lib().azref_tafl_last_error().decode())
    k = n.value
    return dict(canonical=canon[:k], v=v[:k], pi=pi[:k], scores=scores, games_completed=done.value,
                avg_game_length=metrics[0], avg_leaf_depth=metrics[1], avg_valid_moves=metrics[2],
                avg_search_entropy=metrics[3], fast_avg_leaf_depth=metrics[4], fast_avg_search_entropy=metrics[5],
                resign_scores=metrics[6:9].copy(), avg_moves_per_turn=metrics[9])


def dims(game):
    s, a, p = C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert lib().azref_tafl_dims(game, C.byref(s), C.byref(a), C.byref(p)) == 0
    return s.value, a.value, p.value


def random_game(game, seed, max_turns=150, max_len=512):
    buf = np.zeros(max_len, np.uint32)
    n = lib().azref_tafl_random_game(game, max_turns, seed, max_len, buf.ctypes.data_as(C.c_void_p))
    return buf[:n].copy()


def replay(game, moves, max_turns=150, want_valid=True, want_canonical=True):
    """The reference's observable state after k = 0..len moves."""
    S, A, P = dims(game)
    moves = np.ascontiguousarray(moves, np.uint32)
    n = len(moves) + 1
    out = dict(boards=np.zeros((n, 3, S, S), np.int8), players=np.zeros(n, np.uint8), turns=np.zeros(n, np.uint32),
               reps=np.zeros(n, np.uint8), terminal=np.zeros(n, np.uint8), scores=np.zeros((n, 3), np.float32),
               n_valid=np.zeros(n, np.uint32), valid=np.zeros((n, A), np.uint8) if want_valid else None,
               canonical=np.zeros((n, P, S, S), np.float32) if want_canonical else None)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib().azref_tafl_replay(game, max_turns, p(moves), len(moves), p(out["boards"]), p(out["players"]),
                                 p(out["turns"]), p(out["reps"]), p(out["terminal"]), p(out["scores"]),
                                 p(out["n_valid"]), p(out["valid"]), p(out["canonical"]))
    if rc != 0:
        raise RuntimeError(lib().azref_tafl_last_error().decode())
    return out


def position(game, board, player, turn, max_turns, rep, move=None):
    """One arbitrary position through the reference: terminal code, legal-move mask, canonical planes, and the
    board after play_move(move). Returns None for board_out when the reference threw."""
    S, A, P = dims(game)
    board = np.ascontiguousarray(board, np.int8)
    term, nv = C.c_uint8(0), C.c_uint32(0)
    valid = np.zeros(A, np.uint8)
    canon = np.zeros((P, S, S), np.float32)
    bout = np.zeros((3, S, S), np.int8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().azref_tafl_position(game, p(board), int(player), int(turn), int(max_turns), int(rep),
                                   0xFFFFFFFF if move is None else int(move), C.cast(C.byref(term), C.c_void_p),
                                   C.cast(C.byref(nv), C.c_void_p), p(valid), p(canon), p(bout))
    return dict(terminal=term.value, n_valid=nv.value, valid=valid, canonical=canon, board_out=bout if rc == 0 else None,
                threw=rc != 0)


EVAL_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def search(game, seed, n_moves, sims, max_turns, cpuct=1.25, fpu_reduction=0.25, root_fpu_zero=False, evaluator=None,
           gumbel_m=0, gumbel_c_visit=50.0, gumbel_c_scale=1.0, epsilon=0.0, root_policy_temp=1.0, shaped_dirichlet=False,
           batch_width=0, act_temp=None, pruned_temp=1.0):
    """One single-tree MCTS run through the reference's MCTS class. evaluator(canonical[P,S,S]) -> (v[3], pi[A]), or
    None for dumb_eval. Returns counts[m][A], q[m][A], moves[m], total leaf depth per move for the moves searched."""
    S, A, P = dims(game)
    counts = np.zeros((n_moves, A), np.uint32)
    q = np.zeros((n_moves, A), np.float32)
    moves = np.zeros(n_moves, np.uint32)
    depth = np.zeros(n_moves, np.uint32)
    policy = np.zeros((n_moves, A), np.float32)
    probs = np.zeros((n_moves, A), np.float32)

    def cb(canon_p, v_p, pi_p, _user):
        canon = np.ctypeslib.as_array(canon_p, shape=(P, S, S))
        v, pi = evaluator(canon)
        np.ctypeslib.as_array(v_p, shape=(3,))[:] = v
        np.ctypeslib.as_array(pi_p, shape=(A,))[:] = pi

    fn = EVAL_FN(cb) if evaluator is not None else None
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib().azref_tafl_search(game, max_turns, seed, cpuct, fpu_reduction, int(root_fpu_zero), n_moves, sims,
                                0 if evaluator is not None else 1, C.cast(fn, C.c_void_p) if fn else None, None,
                                p(counts), p(q), p(moves), p(depth), gumbel_m, gumbel_c_visit, gumbel_c_scale, p(policy),
                                epsilon, root_policy_temp, int(shaped_dirichlet), batch_width,
                                -1.0 if act_temp is None else act_temp, p(probs), pruned_temp)
    if n < 0:
        raise RuntimeError(lib().azref_tafl_last_error().decode())
    if act_temp is not None:
        return counts[:n], q[:n], moves[:n], depth[:n], probs[:n], policy[:n]
    if gumbel_m:
        return counts[:n], q[:n], moves[:n], depth[:n], policy[:n]
    return counts[:n], q[:n], moves[:n], depth[:n]


def symmetries(game, canon, v, pi):
    """GameState::symmetries of one sample through the reference: ([8,P,S,S], [8,3], [8,A])."""
    S, A, P = dims(game)
    canon = np.ascontiguousarray(canon, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    pi = np.ascontiguousarray(pi, np.float32)
    co, vo, po = np.zeros((8, P, S, S), np.float32), np.zeros((8, 3), np.float32), np.zeros((8, A), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib().azref_tafl_symmetries(game, p(canon), p(v), p(pi), p(co), p(vo), p(po))
    assert n in (2, 8)  # tafl: eightSym; Star Gambit: identity + NW-axis mirror
    return co[:n], vo[:n], po[:n]


# ---- Star Gambit (oracle/ref_tafl_driver.cc: games 10-13 the variants' own classes, 20-23 StarGambitUnifiedGS pinned)
SG_SKIRMISH, SG_SHOWDOWN, SG_CLASH, SG_BATTLE = 10, 11, 12, 13
SG_UNIFIED = 20  # + variant
SG_BYTES_STRIDE = 256


def sg_lib():
    L = lib()
    if not hasattr(L, "_sg_ready"):
        vp, u32 = C.c_void_p, C.c_uint32
        L.azref_sg_replay.argtypes = [C.c_int, vp, u32, u32] + [vp] * 9
        L.azref_sg_units.argtypes = [C.c_int, vp, u32, u32, vp, u32, vp]
        L._sg_ready = True
    return L


def sg_dims(game):
    """(board dim of the action / observation grid, actions, planes)"""
    unified = game >= 20
    variant = game % 10
    D = 13 if unified or variant == 3 else 11
    return D, D * D * 10 + 19, 36 if unified else 32


def sg_random_game(game, seed, max_len=4096):
    sg_lib()
    buf = np.zeros(max_len, np.uint32)
    n = lib().azref_tafl_random_game(game, 0, seed, max_len, buf.ctypes.data_as(C.c_void_p))
    return buf[:n].copy()


def sg_replay(game, moves, want_valid=True, want_canonical=True):
    """The reference's observable state after k = 0..len moves of a Star Gambit transcript."""
    D, A, P = sg_dims(game)
    moves = np.ascontiguousarray(moves, np.uint32)
    n = len(moves) + 1
    out = dict(bytes=np.zeros((n, SG_BYTES_STRIDE), np.uint8), bytes_len=np.zeros(n, np.uint32),
               players=np.zeros(n, np.uint8), turns=np.zeros(n, np.uint32), terminal=np.zeros(n, np.uint8),
               scores=np.zeros((n, 3), np.float32), n_valid=np.zeros(n, np.uint32),
               valid=np.zeros((n, A), np.uint8) if want_valid else None,
               canonical=np.zeros((n, P, D, D), np.float32) if want_canonical else None)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = sg_lib().azref_sg_replay(game, p(moves), len(moves), SG_BYTES_STRIDE, p(out["bytes"]), p(out["bytes_len"]),
                                  p(out["players"]), p(out["turns"]), p(out["terminal"]), p(out["scores"]),
                                  p(out["n_valid"]), p(out["valid"]), p(out["canonical"]))
    if rc != 0:
        raise RuntimeError(lib().azref_tafl_last_error().decode())
    return out


def sg_units(game, moves, fire_move=0):
    moves = np.ascontiguousarray(moves, np.uint32)
    units = np.zeros((32, 8), np.int32)
    fire = np.zeros(5, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = sg_lib().azref_sg_units(game, p(moves), len(moves), int(fire_move), p(units), 32, p(fire))
    assert n >= 0, lib().azref_tafl_last_error().decode()
    return units[:n].copy(), fire
