"""`PlayManager(BrandubhGS | OpenTaflGS | TawlbwrddGS, params)` of the drop-in `alphazero` module: the tafl self-play
engine (b2az_tafl_selfplay_*) behind the reference's PlayManager surface — play() with EvalType.RANDOM, the
build_batch / update_inferences hand-off with EvalType.NN, build_history_batch, scores and the metric getters — against
the UNMODIFIED reference PlayManager (oracle/_ref/libazref_tafl.so) slot by slot. (Named zz: it runs after the engine's
own parity tests.)"""
import threading

import numpy as np
import pytest

import parity_harness as ph
import tafl_ref
from conftest import has_cuda
from test_pybind_module import module

needs_tafl_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")
gpu = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]
CLS = {0: "BrandubhGS", 1: "OpenTaflGS", 2: "TawlbwrddGS"}


def _params(az, G, per_slot, visits, seed, random_eval, **kw):
    p = az.PlayParams()
    p.games_to_play, p.concurrent_games, p.max_batch_size = G * per_slot, G, G
    p.mcts_visits = [visits, visits]
    p.model_groups = [0, 0]  # game_runner.set_model_groups for self-play
    p.history_enabled = p.self_play = p.tree_reuse = True
    p.cpuct, p.fpu_reduction = 1.25, 0.25
    for k, v in kw.items():
        setattr(p, k, v)
    p.seed = seed
    if random_eval:
        p.eval_type = [az.EvalType.RANDOM, az.EvalType.RANDOM]
    return p


def _history(az, pm, game, cap):
    S, A, P = tafl_ref.dims(game) if tafl_ref.available() else {0: (7, 686, 7), 1: (11, 2662, 8), 2: (11, 2662, 7)}[game]
    canon, v, pi = np.zeros((cap, P, S, S), np.float32), np.zeros((cap, 3), np.float32), np.zeros((cap, A), np.float32)
    n = pm.build_history_batch(canon, v, pi)
    return canon[:n], v[:n], pi[:n]


def _rows(canon, v, pi):
    """Order-free view of a sample set: one byte string per sample, sorted."""
    return sorted(canon[i].tobytes() + v[i].tobytes() + pi[i].tobytes() for i in range(len(v)))


REF_KW = {"mcts_root_temp": "root_policy_temp"}


@pytest.mark.parametrize("game,G,per_slot,max_turns,visits,kw", [
    pytest.param(0, 5, 2, 40, 40, dict(epsilon=0.25, mcts_root_temp=1.25, shaped_dirichlet=True, policy_target_pruning=True,
                                       start_temp=1.0, final_temp=0.2, temp_decay_half_life=10.0), marks=gpu, id="brandubh-puct"),
    pytest.param(0, 5, 1, 40, 48, dict(gumbel_enabled=True, gumbel_m=16), marks=gpu, id="brandubh-gumbel"),
    pytest.param(1, 3, 1, 20, 40, dict(gumbel_enabled=True, gumbel_m=16), marks=gpu, id="opentafl-gumbel"),
])
@needs_tafl_ref
def test_random_eval_play_equals_the_reference(game, G, per_slot, max_turns, visits, kw):
    az = module("cuda")
    seed = 4000 + game
    pm = az.PlayManager(getattr(az, CLS[game])(max_turns), _params(az, G, per_slot, visits, seed, True, **kw))
    pm.play()
    assert pm.games_completed() == G * per_slot and pm.remaining_games() == 0
    canon, v, pi = _history(az, pm, game, G * per_slot * max_turns)
    rkw = {REF_KW.get(k, k): x for k, x in kw.items() if k != "gumbel_enabled"}
    refs = [tafl_ref.selfplay(game, seed + g, max_turns, per_slot, visits, **rkw) for g in range(G)]
    want = _rows(np.concatenate([r["canonical"] for r in refs]), np.concatenate([r["v"] for r in refs]),
                 np.concatenate([r["pi"] for r in refs]))
    assert _rows(canon, v, pi) == want
    assert np.array_equal(pm.scores(), np.sum([r["scores"] for r in refs], axis=0))
    lengths = sum(int(round(float(r["avg_game_length"]) * r["games_completed"])) for r in refs)
    assert pm.avg_game_length() == np.float32(np.float32(lengths) / np.float32(G * per_slot))
    assert pm.hist_count() == 0 and pm.simulations() == visits * len(v)


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
@needs_tafl_ref
def test_nn_hand_off_equals_the_random_eval_run(kind):
    """EvalType.NN through build_batch / update_inferences (the reference's batcher and result worker): with the
    evaluator answering dumb_eval's numbers the games equal the EvalType.RANDOM run, sample for sample."""
    az = module(kind)
    game, G, max_turns, visits, seed = 0, 4, 30, 24, 4100
    kw = dict(epsilon=0.25, mcts_root_temp=1.25, policy_target_pruning=True)
    pm0 = az.PlayManager(az.BrandubhGS(max_turns), _params(az, G, 1, visits, seed, True, **kw))
    pm0.play()
    want = _rows(*_history(az, pm0, game, G * max_turns))
    pm = az.PlayManager(az.BrandubhGS(max_turns), _params(az, G, 1, visits, seed, False, **kw))
    t = threading.Thread(target=pm.play)
    t.start()
    S, A, P = tafl_ref.dims(game)
    batch = np.zeros((G, P, S, S), np.float32)
    first = True
    while pm.remaining_games() > 0:
        ids = pm.build_batch(0, batch)
        if not ids:
            continue
        if first:  # game_data(i).gs() (py_wrapper.cc:265-288): the first leaf of a fresh search is the root position itself
            first = False
            for r, i in enumerate(ids):
                gs = pm.game_data(i).gs()
                assert type(gs).__name__ == "BrandubhGS" and gs.current_turn() == 0 and gs.current_player() == 0
                assert np.array_equal(np.asarray(gs.canonicalized()), batch[r])
                assert np.array_equal(np.asarray(pm.game_data(i).valid_moves()), np.asarray(gs.valid_moves()))
        v, pi = np.zeros((len(ids), 3), np.float32), np.zeros((len(ids), A), np.float32)
        for r in range(len(ids)):
            board = (batch[r, :3] > 0).astype(np.int8)
            pos = tafl_ref.position(game, board, 0 if batch[r, 3, 0, 0] > 0 else 1, 0, max_turns, 0)
            valid = pos["valid"].astype(np.float32)
            total = np.float32(int(valid.sum()) % 256)  # dumb_eval: Vector<uint8_t>::sum() wraps (game_state.h:160-173)
            pi[r] = valid / total if total > 0 else valid
            v[r] = np.float32(1.0 / 3.0)
        pm.update_inferences(0, ids, v, pi)
    t.join(timeout=60)
    assert not t.is_alive() and pm.games_completed() == G
    assert _rows(*_history(az, pm, game, G * max_turns)) == want
    assert np.array_equal(pm.scores(), pm0.scores()) and pm.avg_game_length() == pm0.avg_game_length()


def test_tafl_playmanager_validation_and_no_cpu_fallback():
    az = module("emu")
    gs = az.BrandubhGS(30)
    with pytest.raises(RuntimeError, match="You must specify MCTS visits for each player"):
        p = _params(az, 2, 1, 8, 1, True)
        p.mcts_visits = [8]
        az.PlayManager(gs, p)
    with pytest.raises(RuntimeError, match="not implemented"):
        p = _params(az, 2, 1, 8, 1, True)
        p.model_groups = [0, 2]  # three networks: the engines carry two model groups
        az.PlayManager(gs, p)
    with pytest.raises(RuntimeError, match="not implemented"):
        p = _params(az, 2, 1, 8, 1, True)
        p.eval_type = [az.EvalType.PLAYOUT, az.EvalType.PLAYOUT]
        az.PlayManager(gs, p)
    with pytest.raises(RuntimeError, match="multiple of concurrent_games"):
        p = _params(az, 2, 1, 8, 1, True)
        p.games_to_play = 3
        az.PlayManager(gs, p)
    moved = az.BrandubhGS(30)
    moved.play_move(int(np.flatnonzero(np.asarray(moved.valid_moves()))[0]))
    with pytest.raises(RuntimeError, match="initial position"):
        az.PlayManager(moved, _params(az, 2, 1, 8, 1, True))
    if not has_cuda():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            az.PlayManager(gs, _params(az, 2, 1, 8, 1, True))


# ---- Star Gambit behind the same PlayManager surface (py_alphazero.cc try_star_gambit)
@pytest.mark.parametrize("ref_game,make,G,visits,kw", [
    pytest.param(11, lambda az: az.StarGambitShowdownGS(), 3, 32, dict(epsilon=0.25, mcts_root_temp=1.25, policy_target_pruning=True),
                 marks=gpu, id="showdown-puct"),
    pytest.param(22, lambda az: az.StarGambitUnifiedClashGS(), 3, 32, dict(gumbel_enabled=True, gumbel_m=16), marks=gpu,
                 id="unified-clash-gumbel"),
])
@needs_tafl_ref
def test_star_gambit_play_equals_the_reference(ref_game, make, G, visits, kw):
    az = module("cuda")
    seed = 6100 + ref_game
    pm = az.PlayManager(make(az), _params(az, G, 1, visits, seed, True, **kw))
    pm.play()
    assert pm.games_completed() == G and pm.remaining_games() == 0
    D, A, P = tafl_ref.sg_dims(ref_game)
    cap = G * 512
    canon, v, pi = np.zeros((cap, P, D, D), np.float32), np.zeros((cap, 3), np.float32), np.zeros((cap, A), np.float32)
    n = pm.build_history_batch(canon, v, pi)
    rkw = {REF_KW.get(k, k): x for k, x in kw.items() if k != "gumbel_enabled"}
    refs = [tafl_ref.selfplay(ref_game, seed + g, 1024, 1, visits, **rkw) for g in range(G)]
    want = _rows(np.concatenate([r["canonical"] for r in refs]), np.concatenate([r["v"] for r in refs]),
                 np.concatenate([r["pi"] for r in refs]))
    assert _rows(canon[:n], v[:n], pi[:n]) == want
    assert np.array_equal(pm.scores(), np.sum([r["scores"] for r in refs], axis=0))


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
def test_star_gambit_unified_variant_mix_through_the_module(kind):
    """PlayManager(StarGambitUnifiedGS(-1, probs)): every new game draws its variant (randomize_start,
    star_gambit_gs.cc:2421-2425; play_manager.cc:515-516). The reference's draw is unseedable, so the check is on the
    distribution: only variants with weight occur, and all of them do."""
    az = module(kind)
    p = _params(az, 24, 2, 12, 9, True, gumbel_enabled=True, gumbel_m=8)
    pm = az.PlayManager(az.StarGambitUnifiedGS(-1, [0.5, 0.0, 0.25, 0.25]), p)
    pm.play()
    assert pm.games_completed() == 48
    cap = 48 * 512
    canon, v, pi = np.zeros((cap, 36, 13, 13), np.float32), np.zeros((cap, 3), np.float32), np.zeros((cap, 1709), np.float32)
    n = pm.build_history_batch(canon, v, pi)
    assert n > 48
    variants = canon[:n, 32:36].reshape(n, 4, -1).max(axis=2)  # one-hot plane per sample
    assert (variants.sum(axis=1) == 1).all()
    seen = set(np.argmax(variants, axis=1).tolist())
    assert seen == {0, 2, 3}, seen
    # the per-variant tables (play_manager.h:218-275) add up to the global ones
    assert pm.num_tracked_variants() == 4 and pm.variant_games_completed(1) == 0
    assert sum(pm.variant_games_completed(v) for v in range(4)) == 48
    assert np.array_equal(np.sum([np.asarray(pm.variant_scores(v)) for v in range(4)], axis=0), np.asarray(pm.scores()))
    lengths = sum(pm.variant_avg_game_length(v) * pm.variant_games_completed(v) for v in range(4))
    assert abs(lengths - pm.avg_game_length() * 48) < 1e-2 * lengths
    assert all(pm.variant_avg_valid_moves(v) > 1 for v in (0, 2, 3)) and pm.variant_avg_leaf_depth(0) > 0
    assert np.array_equal(np.asarray(pm.variant_perm_scores(2, 0)), np.asarray(pm.variant_scores(2)))


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
@needs_tafl_ref
def test_per_seat_overrides_through_the_module(kind):
    """PlayParams::seat_* (play_manager.h:121-153; what tournament.py sets per player) through alphazero.PlayManager on
    Brandubh: a Gumbel seat against a PUCT seat with noise, the per-seat resign rule; slot g == the unmodified reference
    PlayManager with the same overrides, seeded seed + g. Uniform overrides fold into the globals (no table)."""
    az = module(kind)
    game, G, max_turns, visits, seed = 0, 4, 40, 32, 4300
    p = _params(az, G, 2, visits, seed, True, policy_target_pruning=True)
    p.seat_gumbel_enabled = [[1, 0]]
    p.seat_gumbel_m = [[8, 16]]
    p.seat_epsilon = [[0.0, 0.25]]
    p.seat_mcts_root_temp = [[1.0, 1.25]]
    p.seat_root_fpu_zero = [[0, 1]]
    p.seat_resign_threshold = [[0.0, -2.0]]
    p.seat_resign_consecutive = [[8, 1]]
    pm = az.PlayManager(az.BrandubhGS(max_turns), p)
    pm.play()
    canon, v, pi = _history(az, pm, game, G * 2 * max_turns)
    seat = dict(seat_gumbel_enabled=(1, 0), seat_gumbel_m=(8, 16), seat_epsilon=(0.0, 0.25), seat_root_temp=(1.0, 1.25),
                seat_root_fpu_zero=(0, 1), seat_resign_threshold=(0.0, -2.0), seat_resign_consecutive=(8, 1))
    refs = [tafl_ref.selfplay(game, seed + g, max_turns, 2, visits, policy_target_pruning=True, seat_cfg=seat) for g in range(G)]
    want = _rows(np.concatenate([r["canonical"] for r in refs]), np.concatenate([r["v"] for r in refs]),
                 np.concatenate([r["pi"] for r in refs]))
    assert _rows(canon, v, pi) == want
    assert np.array_equal(pm.scores(), np.sum([r["scores"] for r in refs], axis=0))
    assert np.array_equal(pm.resign_scores(), np.sum([r["resign_scores"] for r in refs], axis=0))
    # uniform per-seat tables equal the globals
    q = _params(az, G, 1, visits, seed, True, epsilon=0.25, mcts_root_temp=1.25)
    pm0 = az.PlayManager(az.BrandubhGS(max_turns), q)
    pm0.play()
    q = _params(az, G, 1, visits, seed, True)
    q.seat_epsilon = [[0.25, 0.25]]
    q.seat_mcts_root_temp = [[1.25, 1.25]]
    pm1 = az.PlayManager(az.BrandubhGS(max_turns), q)
    pm1.play()
    assert _rows(*_history(az, pm0, game, G * max_turns)) == _rows(*_history(az, pm1, game, G * max_turns))
