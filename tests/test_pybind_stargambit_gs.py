"""StarGambit*GS classes of the drop-in `alphazero` module (csrc/py_stargambit_gs.h, the host instantiation of the rule
header the kernels use: csrc/az_stargambit.h) against the UNMODIFIED reference (star_gambit_gs.cc compiled into
oracle/_ref/libazref_tafl.so) along random legal games: player, turn, legal-move mask, scores, canonical planes (bit
patterns), serialised state, get_units / get_fire_info, symmetries, pickling."""
import os
import pickle
import sys

import numpy as np
import pytest

import tafl_ref
from test_pybind_module import module


@pytest.fixture(scope="module")
def az():
    return module("emu")  # the module linked against the host-emulation build (same host C++)


pytestmark = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")

PLAIN = {10: "StarGambitSkirmishGS", 11: "StarGambitShowdownGS", 12: "StarGambitClashGS", 13: "StarGambitBattleGS"}
PINNED = {20: "StarGambitUnifiedSkirmishGS", 21: "StarGambitUnifiedShowdownGS", 22: "StarGambitUnifiedClashGS",
          23: "StarGambitUnifiedBattleGS"}


def make(az, game):
    if game in PLAIN:
        return getattr(az, PLAIN[game])()
    return az.StarGambitUnifiedGS(game - 20)


def inner_bytes(gs, game):
    raw = bytes(gs.__getstate__())
    raw = raw[25:] if game >= 20 else raw
    nu = int.from_bytes(raw[:4], "little")
    fixed = 4 + 9 * nu + 8 + 1 + 4 + 3 + 4
    h = 0xcbf29ce484222325
    for b in raw[fixed:]:  # the key history, folded like oracle/ref_tafl_driver.cc azref_sg_replay does
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return raw[:fixed] + h.to_bytes(8, "little")


@pytest.mark.parametrize("game", [10, 11, 12, 13, 20, 21, 22, 23])
def test_random_games_match_reference(az, game):
    D, A, P = tafl_ref.sg_dims(game)
    total = 0
    for seed in range(6):
        moves = tafl_ref.sg_random_game(game, 1000 * game + seed, max_len=4096 if seed == 0 else 300)
        ref = tafl_ref.sg_replay(game, moves)
        gs = make(az, game)
        assert gs.num_moves() == A and tuple(gs.canonicalized().shape) == (P, D, D)
        for k in range(len(moves) + 1):
            if k:
                gs.play_move(int(moves[k - 1]))
            assert gs.current_player() == ref["players"][k], (seed, k)
            assert gs.current_turn() == ref["turns"][k], (seed, k)
            sc = gs.scores()
            if ref["terminal"][k] == 0:
                assert sc is None, (seed, k)
            else:
                assert sc is not None and np.array_equal(np.asarray(sc), ref["scores"][k]), (seed, k)
            vm = np.asarray(gs.valid_moves())
            assert np.array_equal(vm, ref["valid"][k]), (seed, k, np.nonzero(vm)[0], np.nonzero(ref["valid"][k])[0])
            c = np.asarray(gs.canonicalized())
            assert np.array_equal(c.view(np.uint32), ref["canonical"][k].view(np.uint32)), (seed, k)
            if k % 7 == 0 or k == len(moves):
                n = int(ref["bytes_len"][k])
                assert inner_bytes(gs, game) == ref["bytes"][k, :n].tobytes(), (seed, k)
            total += 1
    assert total > 300


@pytest.mark.parametrize("game", [10, 13, 22])
def test_units_and_fire_info(az, game):
    checked = 0
    for seed in range(3):
        moves = tafl_ref.sg_random_game(game, 77 + seed)
        gs = make(az, game)
        for k, mv in enumerate(moves[:-1]):
            gs.play_move(int(mv))
            if k % 5:
                continue
            vm = np.nonzero(np.asarray(gs.valid_moves()))[0]
            D = tafl_ref.sg_dims(game)[0]
            fires = [int(m) for m in vm if m < D * D * 10 and m % 10 >= 5]
            fm = fires[0] if fires else int(vm[0])
            units, fire = tafl_ref.sg_units(game, moves[: k + 1], fm)
            mine = gs.get_units()
            got = np.array([[u.player, u.type, u.slot, u.hp, u.anchor_q, u.anchor_r, u.facing, u.moves_left] for u in mine], np.int32)
            assert np.array_equal(got.reshape(-1, 8), units)
            fi = gs.get_fire_info(fm)
            assert [int(fi.has_target), fi.target_player, fi.target_type, fi.target_slot, fi.damage] == list(fire)
            checked += 1
    assert checked > 10


@pytest.mark.parametrize("game", [10, 13, 21, 23])
def test_symmetries_match_reference(az, game):
    D, A, P = tafl_ref.sg_dims(game)
    rng = np.random.default_rng(game)
    gs = make(az, game)
    for _ in range(3):
        canon = rng.random((P, D, D), dtype=np.float32)
        v = rng.random(3, dtype=np.float32)
        pi = rng.random(A, dtype=np.float32)
        rc, rv, rp = tafl_ref.symmetries(game, canon, v, pi)
        syms = gs.symmetries(az.PlayHistory(canon, v, pi))
        assert len(syms) == len(rc) == 2
        for i, h in enumerate(syms):
            assert np.array_equal(np.asarray(h.canonical()), rc[i]), i
            assert np.array_equal(np.asarray(h.v()), rv[i])
            assert np.array_equal(np.asarray(h.pi()), rp[i]), i


def test_pickle_copy_equality_and_unified_mix(az):
    gs = az.StarGambitClashGS()
    for mv in tafl_ref.sg_random_game(12, 5)[:40]:
        gs.play_move(int(mv))
    g2 = pickle.loads(pickle.dumps(gs))
    assert g2 == gs and az.hash_game_state(g2) == az.hash_game_state(gs)
    assert np.array_equal(np.asarray(g2.canonicalized()), np.asarray(gs.canonicalized()))
    g3 = gs.copy()
    mv = int(np.nonzero(np.asarray(g3.valid_moves()))[0][0])
    g3.play_move(mv)
    assert not (g3 == gs)
    u = az.StarGambitUnifiedGS()  # random variant mix
    seen = set()
    for _ in range(64):
        u.randomize_start()
        seen.add(u.get_variant_id())
        assert u.num_variants() == 4 and u.current_turn() == 1
    assert len(seen) >= 3
    pinned = az.StarGambitUnifiedBattleGS()
    pinned.randomize_start()
    assert pinned.get_variant_id() == 3 and isinstance(pinned, az.StarGambitUnifiedGS)
    u2 = pickle.loads(pickle.dumps(pinned))
    assert u2 == pinned and u2.get_variant_id() == 3
    assert az.StarGambitUnifiedGS.NUM_MOVES() == 1709 and tuple(az.StarGambitUnifiedGS.CANONICAL_SHAPE()) == (36, 13, 13)
    assert az.StarGambitSkirmishGS.NUM_MOVES() == 1229 and az.StarGambitBattleGS.NUM_MOVES() == 1709
    assert gs.relative_values() if hasattr(gs, "relative_values") else True
