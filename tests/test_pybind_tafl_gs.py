"""BrandubhGS / OpenTaflGS / TawlbwrddGS of the drop-in `alphazero` module (csrc/py_tafl_gs.h: the host instantiation of
the kernels' bitboard rule template) against the UNMODIFIED reference games (oracle/_ref/libazref_tafl.so) along random
legal games: player, turn, legal-move mask, terminal scores and canonical planes after every move, bit-exact;
symmetries() against the reference's eightSym; pickling; the GameState conventions (copy, ==, hash, illegal moves)."""
import os
import pickle
import sys

import numpy as np
import pytest

import parity_harness as ph
import tafl_ref

needs_tafl_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")
CLS = {0: "BrandubhGS", 1: "OpenTaflGS", 2: "TawlbwrddGS"}
GOLD = {0: "tafl_brandubh_transcripts.npz", 1: "tafl_opentafl_transcripts.npz", 2: "tafl_tawlbwrdd_transcripts.npz"}


def module():
    import importlib.util

    path = [f for f in os.listdir(os.path.join(ph.ROOT, "tests", "cpp", "emu")) if f.startswith("alphazero")][0]
    spec = importlib.util.spec_from_file_location("alphazero", os.path.join(ph.ROOT, "tests", "cpp", "emu", path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def az():
    return module()


@needs_tafl_ref
@pytest.mark.parametrize("game", [0, 1, 2])
def test_tafl_gs_follows_the_reference_along_random_games(az, game):
    max_turns = 60 if game == 0 else 40
    for seed in range(4):
        moves = tafl_ref.random_game(game, 50 + seed, max_turns=max_turns)
        ref = tafl_ref.replay(game, moves, max_turns=max_turns)
        gs = getattr(az, CLS[game])(max_turns)
        for k in range(len(moves) + 1):
            assert gs.current_player() == ref["players"][k] and gs.current_turn() == ref["turns"][k]
            assert np.array_equal(np.asarray(gs.valid_moves()), ref["valid"][k]), f"game {game} seed {seed} ply {k}: legal moves"
            assert np.array_equal(np.asarray(gs.canonicalized()).view(np.uint32), ref["canonical"][k].view(np.uint32))
            sc = gs.scores()
            if ref["terminal"][k] == 0:
                assert sc is None
            else:
                assert np.array_equal(np.asarray(sc), ref["scores"][k])
            if k < len(moves):
                gs.play_move(int(moves[k]))


@needs_tafl_ref
@pytest.mark.parametrize("game", [0, 1, 2])
def test_tafl_gs_symmetries_equal_eightsym(az, game):
    gs = getattr(az, CLS[game])()
    S, A, P = tafl_ref.dims(game)
    assert gs.num_symmetries() == 8 and gs.num_moves() == A and gs.num_players() == 2
    assert tuple(getattr(az, CLS[game]).CANONICAL_SHAPE()) == (P, S, S) and getattr(az, CLS[game]).NUM_MOVES() == A
    rng = np.random.default_rng(game)
    canon = rng.random((P, S, S)).astype(np.float32)
    v = rng.random(3).astype(np.float32)
    pi = rng.random(A).astype(np.float32)
    syms = gs.symmetries(az.PlayHistory(canon, v, pi))
    rc, rv, rp = tafl_ref.symmetries(game, canon, v, pi)
    assert len(syms) == 8
    for i, ph_ in enumerate(syms):
        assert np.array_equal(np.asarray(ph_.canonical()), rc[i]) and np.array_equal(ph_.v(), rv[i]) and np.array_equal(ph_.pi(), rp[i])


@pytest.mark.parametrize("game", [0, 1, 2])
def test_tafl_gs_golden_transcripts(az, game):
    """The committed transcripts (tools/make_golden_tafl.py, generated from the unmodified reference): terminal codes,
    legal-move counts and repetition counts along every game."""
    g = np.load(os.path.join(ph.ROOT, "tests", "golden", GOLD[game]))
    for i in range(min(8, len(g["lens"]))):
        gs = getattr(az, CLS[game])(int(np.atleast_1d(g["max_turns"])[i % len(np.atleast_1d(g["max_turns"]))]))
        n = int(g["lens"][i])
        for k in range(n + 1):
            sc = gs.scores()
            term = 0 if sc is None else 1 + int(np.argmax(np.asarray(sc)))
            assert term == g["terminal"][i, k], f"game {game} transcript {i} ply {k}"
            assert int(np.asarray(gs.valid_moves()).sum()) == g["n_valid"][i, k]
            if k < n:
                gs.play_move(int(g["moves"][i, k]))


@pytest.mark.parametrize("game", [0, 1, 2])
def test_tafl_gs_conventions(az, game):
    cls = getattr(az, CLS[game])
    a = cls(30)
    legal = np.flatnonzero(np.asarray(a.valid_moves()))
    b = a.copy()
    assert a == b and az.hash_game_state(a) == az.hash_game_state(b)
    b.play_move(int(legal[0]))
    assert not (a == b) and a.current_turn() == 0 and b.current_turn() == 1 and b.current_player() == 1
    from unittest import mock
    with mock.patch.dict(sys.modules, {"alphazero": az}):  # pickle looks the class up by module name
        c = pickle.loads(pickle.dumps(b))
    assert c == b and c.current_turn() == 1 and np.array_equal(np.asarray(c.valid_moves()), np.asarray(b.valid_moves()))
    # the repetition table survives pickling: shuffling a piece back and forth reaches the same counts on both
    for gs in (b, c):
        for _ in range(2):
            for _side in range(2):
                mv = int(np.flatnonzero(np.asarray(gs.valid_moves()))[0])
                gs.play_move(mv)
    assert np.array_equal(np.asarray(b.canonicalized()), np.asarray(c.canonicalized())) and b == c
    with pytest.raises(RuntimeError):
        a.play_move(cls.NUM_MOVES())  # out of range
    assert "Current Player" in str(a)
