"""The drop-in `alphazero` pybind11 module (csrc/py_alphazero.cc) driven the way the reference's Python drives
its own module (src/game_runner.py:648-745: play() on worker threads, a batcher calling build_batch, a result
worker calling update_inferences, a history saver calling build_history_batch).

CPU: the copy of the module linked against the host-emulation build (tests/cpp/emu) — same C++ host side, same
C ABI. GPU (-m gpu): the product module next to libb2az.so.
Known answers come from the reference's own tests (src/connect4_gs_test.cc) and from the golden traces generated
from the unmodified reference (tests/golden/, tools/make_golden.py)."""
import hashlib
import importlib.util
import os
import pickle
import threading

import numpy as np
import pytest

import parity_harness as ph
from conftest import has_cuda

ROOT = ph.ROOT


def _load(path_dir):
    import sysconfig

    path = os.path.join(path_dir, "alphazero" + sysconfig.get_config_var("EXT_SUFFIX"))
    spec = importlib.util.spec_from_file_location("alphazero", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_mods = {}


def module(kind):
    if kind not in _mods:
        _mods[kind] = _load(os.path.join(ROOT, "tests", "cpp", "emu") if kind == "emu"
                            else os.path.join(ROOT, "alphazero-pybind11_b200"))
    return _mods[kind]


def kinds():
    out = [pytest.param("emu", id="host-emulation")]
    out.append(pytest.param("cuda", id="cuda", marks=[pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]))
    return out


# ------------------------------------------------------------------------------------ surface
def test_reference_surface_is_present():
    az = module("emu")
    for name in ["PlayManager", "PlayParams", "GameData", "GameState", "Connect4GS", "BrandubhGS", "OpenTaflGS", "TawlbwrddGS",
                 "PlayHistory", "EvalType",
                 "hash_game_state", "tracy_is_enabled", "tracy_frame_mark", "_tracy_zone_begin", "_tracy_zone_end",
                 "_tracy_set_thread_name"]:
        assert hasattr(az, name), name
    pm_methods = ["game_data", "params", "scores", "resign_scores", "games_completed", "remaining_games", "stop", "stopped",
                  "awaiting_inference_count", "awaiting_mcts_count", "hist_count", "cache_hits", "cache_misses",
                  "cache_evictions", "cache_reinserts", "cache_max_size", "cache_size", "avg_game_length", "avg_leaf_depth",
                  "avg_search_entropy", "fast_avg_leaf_depth", "fast_avg_search_entropy", "avg_moves_per_turn",
                  "avg_valid_moves", "play", "pop_game", "pop_games_upto", "push_inference", "update_inferences",
                  "build_history_batch", "num_model_groups", "num_seat_perms", "perm_scores", "perm_games_completed",
                  "num_tracked_variants", "set_eager", "build_batch"]
    for m in pm_methods:  # py_wrapper.cc:362-504
        assert hasattr(az.PlayManager, m), m
    p = az.PlayParams()
    # defaults of play_manager.h:60-154
    assert (p.max_batch_size, p.max_cache_size, p.cache_shards, p.cpuct, p.start_temp, p.final_temp) == (1, 0, 1, 2.0, 1.0, 1.0)
    assert (p.tree_reuse, p.history_enabled, p.self_play, p.epsilon, p.mcts_root_temp) == (True, False, False, 0.0, 1.0)
    assert (p.playout_cap_depth, p.playout_cap_percent, p.gumbel_m, p.gumbel_c_visit) == (25, 0.75, 16, 50.0)
    p.mcts_visits = [7, 9]
    p.seat_perms = [[0, 1], [1, 0]]
    assert p.mcts_visits == [7, 9] and p.seat_perms == [[0, 1], [1, 0]]
    assert az.tracy_is_enabled() is False
    assert int(az.EvalType.RANDOM) == 1


def test_connect4_known_answers():
    """src/connect4_gs_test.cc restated through the Python surface."""
    az = module("emu")
    gs = az.Connect4GS()
    assert (az.Connect4GS.NUM_PLAYERS(), az.Connect4GS.NUM_MOVES(), az.Connect4GS.NUM_SYMMETRIES()) == (2, 7, 2)
    assert list(az.Connect4GS.CANONICAL_SHAPE()) == [4, 6, 7]
    assert gs.scores() is None and gs.valid_moves().tolist() == [1] * 7 and gs.current_player() == 0
    # a column fills after 6 stones and becomes invalid; a 7th raises (connect4_gs.cc:57)
    for _ in range(6):
        gs.play_move(2)
    assert gs.valid_moves().tolist() == [1, 1, 0, 1, 1, 1, 1] and gs.current_turn() == 6
    with pytest.raises(RuntimeError):
        gs.play_move(2)
    # horizontal win for player 0
    g = az.Connect4GS()
    for m in [0, 0, 1, 1, 2, 2, 3]:
        g.play_move(m)
    assert g.scores().tolist() == [1.0, 0.0, 0.0]
    # vertical win for player 1
    g = az.Connect4GS()
    for m in [0, 1, 0, 1, 2, 1, 3, 1]:
        g.play_move(m)
    assert g.scores().tolist() == [0.0, 1.0, 0.0]
    # both diagonals
    board = np.zeros((2, 6, 7), np.int8)
    for i in range(4):
        board[0, 5 - i, i] = 1
    assert az.Connect4GS(board, 1, 7).scores().tolist() == [1.0, 0.0, 0.0]
    board = np.zeros((2, 6, 7), np.int8)
    for i in range(4):
        board[1, 2 + i, 3 + i] = 1
    assert az.Connect4GS(board, 0, 8).scores().tolist() == [0.0, 1.0, 0.0]
    # full board without a line: draw
    board = np.zeros((2, 6, 7), np.int8)
    pat = ["0011001", "1100110", "0011001", "1100110", "0011001", "1100110"]
    for h in range(6):
        for w in range(7):
            board[int(pat[h][w]), h, w] = 1
    full = az.Connect4GS(board, 0, 42)
    if full.scores().tolist() == [0.0, 0.0, 1.0]:
        assert full.valid_moves().sum() == 0
    # canonical planes: absolute stones, side-to-move plane (connect4_gs.cc:131-149)
    g = az.Connect4GS()
    g.play_move(3)
    c = g.canonicalized()
    assert c.shape == (4, 6, 7) and c[0, 5, 3] == 1 and c[0].sum() == 1 and c[1].sum() == 0
    assert c[3].min() == 1 and c[2].max() == 0  # player 1 to move
    # equality ignores the turn counter, pickling round-trips, copy is independent
    a, b = az.Connect4GS(np.zeros((2, 6, 7), np.int8), 0, 0), az.Connect4GS(np.zeros((2, 6, 7), np.int8), 0, 5)
    assert a == b and az.hash_game_state(a) == az.hash_game_state(b)
    g2 = pickle.loads(pickle.dumps(g))
    assert g2 == g and g2.current_turn() == 1 and str(g2) == str(g)
    cp = g.copy()
    cp.play_move(0)
    assert not (cp == g) and az.hash_game_state(cp) != az.hash_game_state(g)
    with pytest.raises(RuntimeError, match="Improper connect 4 board shape"):
        az.Connect4GS(np.zeros((2, 7, 6), np.int8), 0, 0)
    # symmetries: identity + mirror (connect4_gs.cc:151-170)
    ph_ = az.PlayHistory(c, np.array([1, 0, 0], np.float32), np.arange(7, dtype=np.float32))
    syms = g.symmetries(ph_)
    assert len(syms) == 2 and np.array_equal(np.asarray(syms[1].canonical()), c[:, :, ::-1])
    assert syms[1].pi().tolist() == list(range(6, -1, -1)) and syms[1].v().tolist() == [1, 0, 0]


def test_error_conventions():
    az = module("emu")
    p = az.PlayParams()
    p.games_to_play = p.concurrent_games = 2
    with pytest.raises(RuntimeError, match="You must specify MCTS visits for each player"):  # play_manager.cc:21
        az.PlayManager(az.Connect4GS(), p)
    p.mcts_visits = [8, 8]
    p.seat_perms = [[0, 1], [1, 0], [0, 1]]  # slot g plays permutation g % 3: needs a multiple of 3 slots
    with pytest.raises(RuntimeError, match="multiple of the number of seat permutations"):
        az.PlayManager(az.Connect4GS(), p)
    p.seat_perms = [[0, 2], [2, 0]]
    with pytest.raises(RuntimeError, match="seat_perms names a model group that does not exist"):  # play_manager.cc:52
        az.PlayManager(az.Connect4GS(), p)
    p.seat_perms = [[0, 1], [1, 0]]
    assert az.PlayManager(az.Connect4GS(), p).num_seat_perms() == 2
    p.seat_perms = []
    with pytest.raises(TypeError):
        az.PlayManager(None, p)
    # the reference's dimension errors (play_manager.cc:57-68)
    p.seat_visits = [[8, 8], [8, 8]]
    with pytest.raises(RuntimeError, match="seat_visits outer dimension must match number of seat permutations"):
        az.PlayManager(az.Connect4GS(), p)
    p.seat_visits = [[8, 8, 8]]
    with pytest.raises(RuntimeError, match="seat_visits inner dimension must match number of players"):
        az.PlayManager(az.Connect4GS(), p)
    p.seat_visits = []
    p.seat_epsilon = [[0.25]]
    with pytest.raises(RuntimeError, match="seat_epsilon inner dimension must match number of players"):
        az.PlayManager(az.Connect4GS(), p)
    p.seat_epsilon = [[0.25, 0.0]]
    with pytest.raises(RuntimeError, match="not implemented"):
        az.PlayManager(az.Connect4GS(), p)
    p.seat_epsilon = []
    pm = az.PlayManager(az.Connect4GS(), p)
    assert pm.num_model_groups() == 2 and pm.num_seat_perms() == 1  # no model_groups: one group per player (play_manager.cc:25-28)
    p.model_groups = [0, 0]
    assert az.PlayManager(az.Connect4GS(), p).num_model_groups() == 1
    with pytest.raises(RuntimeError, match="Improper batch size"):  # py_wrapper.cc:474
        t = threading.Thread(target=pm.play)
        t.start()
        try:
            pm.build_batch(0, np.zeros((2, 3, 6, 7), np.float32))
        finally:
            pm.stop()
            t.join()


# ------------------------------------------------------------------------------------ the thread pipeline
def _params(az, G, games, visits, level, seed, max_batch, deterministic=True, eval_random=False):
    p = az.PlayParams()
    p.games_to_play, p.concurrent_games, p.max_batch_size = games, G, max_batch
    p.mcts_visits = [visits, visits]
    p.model_groups = [0, 0]  # what game_runner.set_model_groups builds for self-play: one network plays both seats
    p.history_enabled = p.self_play = p.tree_reuse = True
    for k, v in ph.level_params(level).items():
        setattr(p, k, bool(v) if k in ("root_fpu_zero", "shaped_dirichlet", "policy_target_pruning") else v)
    p.seed, p.deterministic = seed, deterministic
    if eval_random:
        p.eval_type = [az.EvalType.RANDOM, az.EvalType.RANDOM]
    return p


def _run_pipeline(az, p, workers=2, record=None):
    """play() on `workers` threads; this thread is batcher + evaluator + result worker (fake_net)."""
    pm = az.PlayManager(az.Connect4GS(), p)
    ths = [threading.Thread(target=pm.play) for _ in range(workers)]
    [t.start() for t in ths]
    batch = np.zeros((p.max_batch_size, 4, 6, 7), np.float32)
    gens = 0
    while pm.remaining_games() > 0:
        ids = pm.build_batch(0, batch)
        if not ids:
            continue
        canon = batch[:len(ids)]
        if record is not None:
            record(ids, canon, pm, gens)
        v, pi = ph.fake_net(canon)
        pm.update_inferences(0, ids, v, pi)
        gens += 1
    [t.join() for t in ths]
    return pm, gens


@pytest.mark.parametrize("kind", kinds())
@pytest.mark.parametrize("name", ["nn_level0", "nn_level1_100sims", "nn_level4_gumbel_full"])
def test_pipeline_reproduces_reference_golden(kind, name):
    """Deterministic mode through the Python surface == the unmodified reference, bit for bit: every leaf batch
    (ids + canonical planes), every training sample, scores and metrics (tests/golden, tools/make_golden.py)."""
    az = module(kind)
    G, games, visits, level, seed, et = ph.GOLDEN_CASES[name]
    want = dict(np.load(ph.golden_path(name)))
    digest = hashlib.sha256()

    def record(ids, canon, pm, gens):
        digest.update(np.ascontiguousarray(ids, np.uint32).tobytes())
        digest.update(np.ascontiguousarray(canon, np.float32).tobytes())

    pm, gens = _run_pipeline(az, _params(az, G, games, visits, level, seed, max_batch=G), workers=3, record=record)
    assert gens == int(want["generations"])
    assert np.array_equal(np.frombuffer(digest.digest(), np.uint8), want["leaf_digest"])
    assert pm.games_completed() == games and pm.remaining_games() == 0
    assert np.array_equal(pm.scores().astype(np.float32), want["scores"])
    n = len(want["hist_pi"])
    assert pm.hist_count() == n
    c, v, pi = np.zeros((n + 5, 4, 6, 7), np.float32), np.zeros((n + 5, 3), np.float32), np.zeros((n + 5, 7), np.float32)
    assert pm.build_history_batch(c, v, pi) == n and pm.hist_count() == 0
    assert np.array_equal(c[:n].astype(np.uint8), want["hist_canon"])
    assert np.array_equal(v[:n].view(np.uint32), want["hist_v"].view(np.uint32))
    assert np.array_equal(pi[:n].view(np.uint32), want["hist_pi"].view(np.uint32))
    got = np.array([getattr(pm, k)() for k in ph.METRIC_NAMES], np.float32)
    assert np.array_equal(got.view(np.uint32), want["metrics"].view(np.uint32))


@pytest.mark.parametrize("kind", kinds())
def test_small_batches_and_game_data_views(kind):
    """max_batch_size < concurrent games: a generation is handed out in several build_batch calls; GameData
    exposes the pending leaf (canonical / v / pi views, py_wrapper.cc:265-288) for the pop_game / push_inference
    flavour of the hand-off."""
    az = module(kind)
    p = _params(az, G=6, games=8, visits=16, level=0, seed=5, max_batch=4)
    sizes = []

    def record(ids, canon, pm, gens):
        sizes.append(len(ids))
        gd = pm.game_data(ids[0])
        assert np.array_equal(np.asarray(gd.canonical()), canon[0])
        assert gd.gs().canonicalized().shape == (4, 6, 7) and gd.valid_moves().shape == (7,)

    pm, gens = _run_pipeline(az, p, workers=1, record=record)
    assert pm.games_completed() == 8 and max(sizes) <= 4 and pm.scores().sum() == 8
    assert pm.avg_game_length() >= 7 and pm.num_model_groups() == 1 and pm.params().concurrent_games == 6

    # pop_game / push_inference
    p = _params(az, G=3, games=3, visits=8, level=0, seed=5, max_batch=3)
    pm = az.PlayManager(az.Connect4GS(), p)
    t = threading.Thread(target=pm.play)
    t.start()
    while pm.remaining_games() > 0:
        i = pm.pop_game(0)
        if i is None:
            continue
        gd = pm.game_data(i)
        v, pi = ph.fake_net(np.asarray(gd.canonical())[None])
        gd.v()[:] = v[0]
        gd.pi()[:] = pi[0]
        pm.push_inference(i)
    t.join()
    assert pm.games_completed() == 3


@pytest.mark.parametrize("kind", kinds())
def test_cache_counters_through_the_module(kind, monkeypatch):
    """max_cache_size > 0: the reference's throughput counter (cache_hits + cache_misses, network_pareto.py:415-423)
    equals the simulations, and hits answer leaves without the evaluator."""
    monkeypatch.setenv("B2AZ_EMU_FLAT", "1")
    az = module(kind)
    p = _params(az, G=16, games=16, visits=32, level=1, seed=3, max_batch=16, deterministic=False)
    p.max_cache_size = 50000
    pm, gens = _run_pipeline(az, p, workers=1)
    assert pm.games_completed() == 16
    assert pm.cache_hits() > 0 and pm.cache_hits() + pm.cache_misses() == pm.simulations()
    assert 0 < pm.cache_size() <= pm.cache_max_size() == 50000


@pytest.mark.parametrize("kind", kinds())
def test_random_eval_play_returns_when_done(kind):
    """EvalType.RANDOM: play() alone finishes the run (play_manager_test.cc); stop() ends it early."""
    az = module(kind)
    p = _params(az, G=32, games=64, visits=24, level=0, seed=9, max_batch=32, deterministic=False, eval_random=True)
    pm = az.PlayManager(az.Connect4GS(), p)
    pm.play()
    assert pm.games_completed() == 64 and pm.remaining_games() == 0 and pm.scores().sum() == 64
    assert pm.hist_count() > 64 * 7
    pm2 = az.PlayManager(az.Connect4GS(), _params(az, 4, 10 ** 6, 24, 0, 9, 4))
    t = threading.Thread(target=pm2.play)
    t.start()
    pm2.stop()
    t.join(timeout=20)
    assert not t.is_alive() and pm2.stopped() and pm2.remaining_games() == 0


@pytest.mark.parametrize("kind", kinds())
def test_two_model_groups_route_leaves_by_searching_seat(kind):
    """No model_groups = one group per player (play_manager.cc:25-28): build_batch(g) serves the leaves of the searches
    seat g runs (awaiting_inference_[seat_perm[cp]], play_manager.cc:577-598). With the same evaluator behind both
    groups the games must equal the one-group run; with the position cache on, the groups must not share entries."""
    az = module(kind)

    def run(groups, cache, two_nets=False):
        p = _params(az, G=6, games=6, visits=24, level=1, seed=11, max_batch=6, deterministic=False)
        p.model_groups = groups
        p.max_cache_size = cache
        pm = az.PlayManager(az.Connect4GS(), p)
        ths = [threading.Thread(target=pm.play) for _ in range(2)]
        [t.start() for t in ths]
        batch = np.zeros((6, 4, 6, 7), np.float32)
        served = [0] * pm.num_model_groups()
        while pm.remaining_games() > 0:
            for g in range(pm.num_model_groups()):
                ids = pm.build_batch(g, batch)
                if not ids:
                    continue
                served[g] += len(ids)
                v, pi = ph.fake_net(batch[:len(ids)])
                if g == 1 and two_nets:  # a different "network" for group 1: its evaluations must never answer group 0
                    v, pi = v[:, ::-1].copy(), pi[:, ::-1].copy()
                pm.update_inferences(g, ids, v, pi)
        [t.join() for t in ths]
        n = pm.hist_count()
        c, v, pi = np.zeros((n, 4, 6, 7), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 7), np.float32)
        assert pm.build_history_batch(c, v, pi) == n
        return pm, served, ph._sorted_rows(c, v, pi)

    one, served1, rows1 = run([0, 0], 0)
    two, served2, rows2 = run([], 0)
    assert one.num_model_groups() == 1 and two.num_model_groups() == 2
    assert min(served2) > 0 and sum(served2) == sum(served1), "both seats search, every leaf is served exactly once"
    assert np.array_equal(rows1, rows2) and np.array_equal(one.scores(), two.scores())
    with pytest.raises(RuntimeError, match="model group out of range"):
        one.build_batch(1, np.zeros((6, 4, 6, 7), np.float32))
    # a different "network" behind group 1, cache off vs cache on: the reference's cache property (test_cache.py:227-253:
    # a search with the cache is the search without it) only holds if the two groups never share an entry
    _, _, rows_off = run([], 0, two_nets=True)
    _, _, rows_on = run([], 100000, two_nets=True)
    assert np.array_equal(rows_off, rows_on) and not np.array_equal(rows_off, rows2)


@pytest.mark.parametrize("kind", kinds())
@pytest.mark.parametrize("past_is_random", [False, True])
def test_play_past_shape_seat_perms_and_mixed_evaluators(kind, past_is_random):
    """What game_runner.play_past builds (game_runner.py:2203-2242): two players = two model groups, seat_perms
    [[0, 1], [1, 0]], n = bs * cb * n_perms games, and against iteration 0 a RandPlayer (EvalType.RANDOM) for group 1.
    Then its read-out (2268-2290): perm_scores / perm_games_completed per permutation."""
    az = module(kind)
    p = _params(az, G=4, games=8, visits=16, level=1, seed=5, max_batch=4, deterministic=False)
    p.model_groups = [0, 1]
    p.seat_perms = [[0, 1], [1, 0]]
    p.eval_type = [az.EvalType.NN, az.EvalType.RANDOM if past_is_random else az.EvalType.NN]
    pm = az.PlayManager(az.Connect4GS(), p)
    assert pm.num_model_groups() == 2 and pm.num_seat_perms() == 2
    ths = [threading.Thread(target=pm.play) for _ in range(2)]
    [t.start() for t in ths]
    batch = np.zeros((4, 4, 6, 7), np.float32)
    served = [0, 0]
    while pm.remaining_games() > 0:
        for g in range(2):
            ids = pm.build_batch(g, batch)
            if not ids:
                continue
            served[g] += len(ids)
            v, pi = ph.fake_net(batch[:len(ids)])
            if g == 1:
                v, pi = v[:, ::-1].copy(), pi[:, ::-1].copy()
            pm.update_inferences(g, ids, v, pi)
    [t.join() for t in ths]
    assert pm.games_completed() == 8
    assert served[0] > 0 and (served[1] == 0) == past_is_random, served
    total = np.zeros(3, np.float32)
    for perm in range(2):
        assert pm.perm_games_completed(perm) == 4  # n / n_perms games under every seating
        sc = np.asarray(pm.perm_scores(perm))
        assert sc.sum() == 4
        total += sc
    assert np.array_equal(total, np.asarray(pm.scores()))
    with pytest.raises(IndexError):
        pm.perm_scores(2)


@pytest.mark.parametrize("kind", kinds())
def test_external_caches_size_the_device_cache(kind):
    """PlayManager(gs, params, caches=[...]) (play_manager.cc:644-649, tournament.py:224): accepted; the objects size the
    engine's device cache (they are not shared between PlayManagers, see py_alphazero.cc). Same games as max_cache_size."""
    az = module(kind)

    def run(make):
        p = _params(az, G=4, games=4, visits=24, level=1, seed=3, max_batch=4, deterministic=False)
        pm = make(p)
        ths = [threading.Thread(target=pm.play) for _ in range(2)]
        [t.start() for t in ths]
        batch = np.zeros((4, 4, 6, 7), np.float32)
        while pm.remaining_games() > 0:
            ids = pm.build_batch(0, batch)
            if ids:
                v, pi = ph.fake_net(batch[:len(ids)])
                pm.update_inferences(0, ids, v, pi)
        [t.join() for t in ths]
        n = pm.hist_count()
        c, v, pi = np.zeros((n, 4, 6, 7), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 7), np.float32)
        assert pm.build_history_batch(c, v, pi) == n
        return pm, ph._sorted_rows(c, v, pi)

    def with_param(p):
        p.max_cache_size = 4096
        return az.PlayManager(az.Connect4GS(), p)

    ext = az.ShardedS3FIFOCache(4096, 2, 3686, 7, 3)
    a, rows_a = run(with_param)
    b, rows_b = run(lambda p: az.PlayManager(az.Connect4GS(), p, caches=[ext, None]))
    assert np.array_equal(rows_a, rows_b)
    assert b.cache_hits() == a.cache_hits() > 0 and b.cache_misses() == a.cache_misses()
    assert ext.size() == 0  # the host object is not filled
    with pytest.raises(TypeError):
        az.PlayManager(az.Connect4GS(), _params(az, 2, 2, 8, 0, 1, 2), caches=[object()])


def test_playout_eval_contract():
    """playout_eval / playout_eval_batch (py_wrapper.cc:726-770, game_state.cc:10-95): uniform prior over the legal moves,
    value = one-hot outcome of a random playout (or 1/(P+1) each if the game cannot go on)."""
    az = module("emu")
    g = az.Connect4GS()
    for m in (3, 3, 3, 3, 3, 3):  # fill column 3
        g.play_move(m)
    v, pi = az.playout_eval(g)
    assert pi.shape == (7,) and pi[3] == 0 and np.allclose(pi[[0, 1, 2, 4, 5, 6]], 1 / 6)
    assert v.shape == (3,) and sorted(v.tolist()) == [0, 0, 1]
    vs, pis = az.playout_eval_batch([g, az.Connect4GS(), az.BrandubhGS(20)][:2])
    assert vs.shape == (2, 3) and pis.shape == (2, 7) and np.allclose(pis[1], 1 / 7) and np.allclose(vs.sum(1), 1)
    outcomes = np.stack([az.playout_eval(az.Connect4GS())[0] for _ in range(200)])
    assert outcomes[:, 0].sum() > 60 and outcomes[:, 1].sum() > 40  # both sides win random playouts
