"""The `MCTS` class of the drop-in `alphazero` module (csrc/py_mcts.h: one tree of the device's wide-tree search behind
the reference's single-tree API, py_wrapper.cc:191-220) driven the way the reference's Python tools drive theirs
(find_leaf -> evaluate -> process_result, counts / probs / root_value, update_root + play_move), against the UNMODIFIED
reference MCTS class over the unmodified games (oracle/_ref/libazref_tafl.so, azref_tafl_search) — same seed, same
pseudo-network: visit counts and Q values bit-exact after every move's search."""
import ctypes as C
import zlib

import numpy as np
import pytest

import tafl_ref
from conftest import has_cuda
from test_pybind_module import module

needs_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")
gpu = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def make_game(az, game):
    if game == 0:
        return az.BrandubhGS(150)
    if game == 1:
        return az.OpenTaflGS(400)
    if 10 <= game <= 13:
        return getattr(az, ["StarGambitSkirmishGS", "StarGambitShowdownGS", "StarGambitClashGS", "StarGambitBattleGS"][game - 10])()
    return az.StarGambitUnifiedGS(game - 20)


def net_for(A):
    def net(canon):
        rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(canon, np.float32).tobytes()))
        v = rng.random(3).astype(np.float32) + np.float32(0.05)
        v /= v.sum()
        pi = rng.random(A).astype(np.float32) ** 4 + np.float32(1e-3)
        pi /= pi.sum()
        return v.astype(np.float32), pi.astype(np.float32)
    return net


@pytest.mark.parametrize("game,n_moves,sims", [pytest.param(0, 10, 40, marks=gpu), pytest.param(12, 12, 40, marks=gpu),
                                               pytest.param(23, 10, 40, marks=gpu)])
@needs_ref
def test_python_driven_tree_equals_the_reference(game, n_moves, sims):
    az = module("cuda")
    gs = make_game(az, game)
    A = gs.num_moves()
    net = net_for(A)
    seed = 31337 + game
    rc, rq, rm, rd = tafl_ref.search(game, seed, n_moves, sims, 150 if game == 0 else 400, 1.25, 0.25, False, net)
    mcts = az.MCTS(1.25, 2, A, 0.0, 1.0, 0.25, game >= 10)
    mcts.seed(seed)
    for m in range(len(rm)):
        for _ in range(sims):
            leaf = mcts.find_leaf(gs)
            v, pi = net(np.asarray(leaf.canonicalized()))
            mcts.process_result(gs, v, pi, False)
        assert mcts.depth() == sims
        counts = np.asarray(mcts.counts())
        assert np.array_equal(counts, rc[m]), (game, m)
        assert np.array_equal(np.asarray(mcts.root_q_values()).view(np.uint32), rq[m].view(np.uint32)), (game, m)
        pr = np.asarray(mcts.probs(1.0))
        assert abs(float(pr.sum()) - 1.0) < 1e-5 and np.array_equal(pr > 0, counts > 0)
        pv = np.asarray(mcts.principal_variation(4))
        assert len(pv) >= 1 and counts[pv[0]] == counts.max()
        wld = np.asarray(mcts.root_value())
        assert abs(float(wld.sum()) - 1.0) < 1e-5
        mv = int(rm[m])
        mcts.update_root(gs, mv)
        gs.play_move(mv)
    assert mcts.root_n() >= 0


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
def test_wu_uct_batch_protocol_and_value_write_back(kind):
    az = module(kind)
    gs = az.StarGambitUnifiedGS(0)
    A = gs.num_moves()
    net = net_for(A)
    mcts = az.MCTS(1.25, 2, A, 0.0, 1.0, 0.25, True)
    mcts.seed(5)
    for _ in range(6):  # rounds of 4 pending leaves (play.py / mcts_analysis.py with a batched net)
        leaves = [mcts.find_leaf_batched(gs) for _ in range(4)]
        assert mcts.in_flight_count() == 4
        for i, leaf in enumerate(leaves):
            v, pi = net(np.asarray(leaf.canonicalized()))
            v_in = v.copy()
            mcts.process_result_batched(gs, i, v, pi, False)
            if leaf.scores() is None and leaf.current_player() == 1:
                assert np.array_equal(v, v_in[[1, 0, 2]])  # relative -> absolute, written through (mcts.cc:814-815)
        mcts.reset_batch()
        assert mcts.in_flight_count() == 0
    assert mcts.depth() == 24 and int(np.asarray(mcts.counts()).sum()) == mcts.root_n() - 1


def test_mcts_without_a_device_fails_loudly():
    az = module("emu")
    gs = az.BrandubhGS(150)
    mcts = az.MCTS(1.25, 2, gs.num_moves())
    with pytest.raises(RuntimeError, match="no CUDA device"):
        mcts.find_leaf(gs)
    with pytest.raises(RuntimeError, match="two-player"):
        az.MCTS(1.25, 3, 10)
    az.MCTS(1.25, 2, gs.num_moves(), 0.0, 1.0, 0.0, False, False, False, True, 16, 50.0, 1.0, True)  # gumbel_full is accepted
    assert az.MCTS.pick_move(np.array([0.0, 1.0, 0.0], np.float32)) == 1


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
def test_connect4_tree_equals_the_reference_mcts(kind):
    """MCTS over Connect4GS (game 30 of the wide-tree search) against the unmodified reference MCTS class driven through
    oracle/ref_driver.cc — same seed, same pseudo-network, a position that is not the start position."""
    import refdriver

    if not refdriver.available():
        pytest.skip("oracle/_ref/libazref.so not built")
    az = module(kind)
    L = refdriver.lib()
    net = net_for(7)
    seed, sims = 2024, 60
    gs = az.Connect4GS()
    rgs = L.azref_c4_new()
    for mv in (3, 3, 4):
        gs.play_move(mv)
        assert L.azref_c4_play(rgs, mv) == 0
    cfg = refdriver.MctsCfg(cpuct=1.25, num_players=2, num_moves=7, epsilon=0.0, root_policy_temp=1.0, fpu_reduction=0.25,
                            gumbel_m=16, gumbel_c_visit=50.0, gumbel_c_scale=1.0)
    rt = L.azref_mcts_new(C.byref(cfg))
    L.azref_seed_thread_rng(seed)
    mcts = az.MCTS(1.25, 2, 7, 0.0, 1.0, 0.25)
    mcts.seed(seed)
    canon = np.zeros((4, 6, 7), np.float32)
    for move_no in range(6):
        for _ in range(sims):
            leaf = mcts.find_leaf(gs)
            v, pi = net(np.asarray(leaf.canonicalized()))
            mcts.process_result(gs, v, pi, False)
            rleaf = L.azref_mcts_find_leaf(rt, rgs)
            L.azref_c4_canonical(rleaf, refdriver.P(canon))
            rv, rpi = net(canon)
            L.azref_mcts_process_result(rt, rgs, refdriver.P(rv), 3, refdriver.P(rpi), 7, 0)
        rc, rq = np.zeros(7, np.uint32), np.zeros(7, np.float32)
        L.azref_mcts_counts(rt, refdriver.P(rc))
        L.azref_mcts_root_q(rt, refdriver.P(rq))
        assert np.array_equal(np.asarray(mcts.counts()), rc), move_no
        assert np.array_equal(np.asarray(mcts.root_q_values()).view(np.uint32), rq.view(np.uint32)), move_no
        rwld = np.zeros(3, np.float32)
        L.azref_mcts_root_value(rt, refdriver.P(rwld))
        assert np.array_equal(np.asarray(mcts.root_value()).view(np.uint32), rwld.view(np.uint32))
        mv = int(np.argmax(rc))
        mcts.update_root(gs, mv)
        gs.play_move(mv)
        assert L.azref_mcts_update_root(rt, rgs, mv) == 0 and L.azref_c4_play(rgs, mv) == 0
        if gs.scores() is not None:
            break
    L.azref_mcts_free(rt)
    L.azref_c4_free(rgs)
