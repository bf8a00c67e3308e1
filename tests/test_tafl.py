"""Tafl game kernels (b2az_tafl_replay: {Brandubh,OpenTafl,Tawlbwrdd}GS::play_move / valid_moves / scores /
canonicalized / repetition table on the device) against the UNMODIFIED reference: committed golden transcripts
(tests/golden/tafl_*_transcripts.npz, tools/make_golden_tafl.py) and, where oracle/_ref/libazref_tafl.so exists,
fresh random games replayed through the reference live. Everything is compared bit-exact.
CPU tests run the same rule header through the host-emulation build; `-m gpu` tests run the CUDA library."""
import os
import zlib

import numpy as np
import pytest

import b2az
import parity_harness as ph
import tafl_ref

NAMES = {0: "brandubh", 1: "opentafl", 2: "tawlbwrdd"}
MAX_TURNS = {0: 150, 1: 400, 2: 400}
needs_tafl_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")


def golden(game):
    return dict(np.load(os.path.join(ph.ROOT, "tests", "golden", f"tafl_{NAMES[game]}_transcripts.npz")))


def _replay_grouped(lib, game, moves, lens, max_turns):
    """b2az_tafl_replay takes one max_turns per call: run the batch in groups."""
    n = len(lens)
    out = None
    for mt in sorted(set(int(x) for x in max_turns)):
        idx = np.nonzero(max_turns == mt)[0]
        r = b2az.tafl_replay(game, moves[idx], lens[idx], mt, lib=lib)
        assert (r["status"] == 0).all()
        if out is None:
            out = {k: np.zeros((n,) + v.shape[1:], v.dtype) for k, v in r.items()}
        for k, v in r.items():
            out[k][idx] = v
    return out


def _check_against_golden(lib, game):
    g = golden(game)
    moves, lens = g["moves"], g["lens"]
    r = _replay_grouped(lib, game, moves, lens, g["max_turns"])
    n = len(lens)
    for i in range(n):
        m = int(lens[i]) + 1
        for k in ("boards", "players", "turns", "reps", "terminal", "n_valid"):
            assert np.array_equal(r[k][i, :m], g[k][i, :m]), f"{NAMES[game]} game {i}: {k} differs from the reference"
        for k in range(m):
            assert zlib.crc32(r["valid"][i, k].tobytes()) == g["valid_crc"][i, k], f"game {i} move {k}: legal-move mask"
            assert zlib.crc32(r["canonical"][i, k].tobytes()) == g["canon_crc"][i, k], f"game {i} move {k}: canonical"
        assert r["valid"][i, :m].sum(axis=1).tolist() == g["n_valid"][i, :m].tolist()
    for i in range(g["valid_full"].shape[0]):
        m = int(lens[i]) + 1
        assert np.array_equal(r["valid"][i, :m], g["valid_full"][i, :m])
        assert np.array_equal(r["canonical"][i, :m].view(np.uint32), g["canon_full"][i, :m].view(np.uint32))
    # every golden game ends: the last position is terminal and no earlier one is
    for i in range(n):
        assert r["terminal"][i, lens[i]] != 0 and (r["terminal"][i, :lens[i]] == 0).all()


def _check_against_live_reference(lib, game, n_games, seed0):
    mt = MAX_TURNS[game]
    games = [tafl_ref.random_game(game, seed0 + i, max_turns=mt, max_len=mt + 8) for i in range(n_games)]
    L = max(len(x) for x in games)
    moves = np.zeros((n_games, L), np.uint16)
    lens = np.array([len(x) for x in games], np.uint32)
    for i, x in enumerate(games):
        moves[i, :len(x)] = x
    r = b2az.tafl_replay(game, moves, lens, mt, lib=lib)
    assert (r["status"] == 0).all()
    for i, x in enumerate(games):
        ref = tafl_ref.replay(game, x, max_turns=mt)
        m = len(x) + 1
        for k in ("boards", "players", "turns", "reps", "terminal", "n_valid", "valid"):
            assert np.array_equal(r[k][i, :m], ref[k]), f"{NAMES[game]} game {i}: {k} differs from the live reference"
        assert np.array_equal(r["canonical"][i, :m].view(np.uint32), ref["canonical"].view(np.uint32))


@pytest.mark.parametrize("game", [0, 1, 2])
def test_rules_host_build_reproduces_golden(game):
    _check_against_golden(b2az.load(ph.HOSTEMU_LIB), game)


@needs_tafl_ref
@pytest.mark.parametrize("game,n", [(0, 64), (1, 12), (2, 16)])
def test_rules_host_build_vs_live_reference(game, n):
    _check_against_live_reference(b2az.load(ph.HOSTEMU_LIB), game, n, 5000)


def test_illegal_move_is_reported():
    lib = b2az.load(ph.HOSTEMU_LIB)
    moves = np.array([[700], [0]], np.uint16)  # out of range; empty source square (0,0)
    r = b2az.tafl_replay(0, moves, np.array([1, 1], np.uint32), 150, lib=lib)
    assert r["status"].tolist() == [-5, -5]


def test_start_position_and_move_ids():
    lib = b2az.load(ph.HOSTEMU_LIB)
    r = b2az.tafl_replay(0, np.zeros((1, 1), np.uint16), np.zeros(1, np.uint32), 150, lib=lib)
    assert r["boards"][0, 0, 0, 3, 3] == 1 and r["boards"][0, 0, 1].sum() == 4 and r["boards"][0, 0, 2].sum() == 8
    assert r["players"][0, 0] == 0 and r["reps"][0, 0] == 1 and r["terminal"][0, 0] == 0
    # attackers to move: 8 pieces; e.g. (0,3) slides along row 0 to columns 1,2,4,5 (corners are king-only)
    v = r["valid"][0, 0].reshape(49, 14)
    assert v[3, :7].tolist() == [0, 1, 1, 0, 1, 1, 0] and v[3, 7:].sum() == 0
    assert r["n_valid"][0, 0] == v.sum()
    for game, pieces in ((1, (1, 12, 24)), (2, (1, 12, 24))):
        r = b2az.tafl_replay(game, np.zeros((1, 1), np.uint16), np.zeros(1, np.uint32), 400, lib=lib)
        assert tuple(int(r["boards"][0, 0, p].sum()) for p in range(3)) == pieces


@pytest.mark.gpu
@pytest.mark.parametrize("game", [0, 1, 2])
def test_cuda_rules_reproduce_golden(game):
    _check_against_golden(None, game)


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,n", [(0, 256), (1, 32), (2, 48)])
def test_cuda_rules_vs_live_reference(game, n):
    _check_against_live_reference(None, game, n, 9000)


# ---------------------------------------------------------------------------------- arbitrary positions
def _random_positions(game, n, seed):
    """Random piece placements (non-king pieces keep off the restricted squares, like in any reachable position):
    dense and sparse boards, kings next to attackers, so captures / encirclement / blocked sides all occur."""
    S = b2az.TAFL_DIMS[game][0]
    rng = np.random.default_rng(seed)
    restricted = {0, S - 1, S * (S - 1), S * S - 1, (S // 2) * S + S // 2} if game != 2 else set()
    boards = np.zeros((n, 3, S, S), np.int8)
    for i in range(n):
        cells = list(rng.permutation(S * S))
        n_def = int(rng.integers(0, 5 if game == 0 else 13))
        n_atk = int(rng.integers(0, 9 if game == 0 else 40))
        if rng.random() < 0.95:
            k = cells.pop()
            boards[i, 0].flat[k] = 1
            # crowd the king with attackers half of the time
            if rng.random() < 0.5:
                h, w = divmod(int(k), S)
                for dh, dw in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                    c = (h + dh) * S + (w + dw)
                    if 0 <= h + dh < S and 0 <= w + dw < S and c in cells and c not in restricted and rng.random() < 0.8:
                        cells.remove(c)
                        boards[i, 2].flat[c] = 1
        free = [c for c in cells if c not in restricted]
        for c in free[:n_def]:
            boards[i, 1].flat[c] = 1
        for c in free[n_def:n_def + n_atk]:
            boards[i, 2].flat[c] = 1
    players = rng.integers(0, 2, n).astype(np.uint8)
    turns = rng.integers(1, 60, n).astype(np.uint32)
    reps = rng.choice([1, 1, 1, 2, 3], n).astype(np.uint8)
    return boards, players, turns, reps, rng


def _check_positions(lib, game, n, seed):
    mt = MAX_TURNS[game]
    boards, players, turns, reps, rng = _random_positions(game, n, seed)
    refs = [tafl_ref.position(game, boards[i], players[i], turns[i], mt, reps[i]) for i in range(n)]
    moves = np.full(n, 0xFFFFFFFF, np.uint32)
    for i, r in enumerate(refs):
        legal = np.nonzero(r["valid"])[0]
        if len(legal):
            moves[i] = legal[rng.integers(0, len(legal))]
    got = b2az.tafl_positions(game, boards, players, turns, reps, mt, moves=moves, lib=lib)
    seen_terminal = set()
    captures = 0
    for i, r in enumerate(refs):
        assert got["terminal"][i] == r["terminal"], f"{NAMES[game]} position {i}: scores()"
        assert got["n_valid"][i] == r["n_valid"] and np.array_equal(got["valid"][i], r["valid"]), f"position {i}: valid_moves()"
        assert np.array_equal(got["canonical"][i].view(np.uint32), r["canonical"].view(np.uint32))
        seen_terminal.add(int(r["terminal"]))
        if moves[i] != 0xFFFFFFFF:
            after = tafl_ref.position(game, boards[i], players[i], turns[i], mt, reps[i], move=moves[i])
            assert not after["threw"] and got["status"][i] == 0
            assert np.array_equal(got["boards_out"][i], after["board_out"]), f"{NAMES[game]} position {i}: play_move()"
            removed = int(boards[i].sum() - after["board_out"].sum())
            assert bool(got["captured_any"][i]) == (removed > 0)
            captures += removed > 0
    assert {0, 1, 2} <= seen_terminal and captures > n // 100  # the sample exercises wins of both sides and captures


@needs_tafl_ref
@pytest.mark.parametrize("game", [0, 1, 2])
def test_positions_host_build_vs_live_reference(game):
    _check_positions(b2az.load(ph.HOSTEMU_LIB), game, 600, 31 + game)


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game", [0, 1, 2])
def test_cuda_positions_vs_live_reference(game):
    _check_positions(None, game, 1500, 77 + game)


# ---------------------------------------------------------------------------------- edge cases of the batch API
def test_empty_ragged_and_full_length_batches():
    lib = b2az.load(ph.HOSTEMU_LIB)
    # n == 0 is a no-op
    r = b2az.tafl_replay(0, np.zeros((0, 4), np.uint16), np.zeros(0, np.uint32), 150, lib=lib)
    assert r["boards"].shape == (0, 5, 3, 7, 7)
    # ragged: lengths 0, 1 and a whole golden game in one batch; rows beyond a game's length stay untouched (zero)
    g = golden(0)
    i = int(np.argmax(g["lens"][:-2]))  # the longest 150-turn game
    L = int(g["lens"][i])
    moves = np.zeros((3, L), np.uint16)
    moves[1, 0] = g["moves"][i, 0]
    moves[2] = g["moves"][i, :L]
    lens = np.array([0, 1, L], np.uint32)
    r = b2az.tafl_replay(0, moves, lens, 150, lib=lib)
    assert (r["status"] == 0).all()
    assert np.array_equal(r["boards"][2, :L + 1], g["boards"][i, :L + 1]) and np.array_equal(r["terminal"][2, :L + 1], g["terminal"][i, :L + 1])
    assert np.array_equal(r["boards"][1, 1], g["boards"][i, 1]) and np.array_equal(r["boards"][0, 0], g["boards"][i, 0])
    assert not r["boards"][0, 1:].any() and not r["valid"][1, 2:].any() and not r["canonical"][0, 1:].any()
    # lens larger than max_len are clamped
    r2 = b2az.tafl_replay(0, moves[2:3, :5], np.array([999], np.uint32), 150, lib=lib)
    assert np.array_equal(r2["boards"][0], g["boards"][i, :6])


def test_bad_arguments_are_rejected():
    lib = b2az.load(ph.HOSTEMU_LIB)
    with pytest.raises(b2az.B2azError):
        b2az.tafl_replay(7, np.zeros((1, 1), np.uint16), np.zeros(1, np.uint32), 150, lib=lib)  # unknown game
    with pytest.raises(b2az.B2azError):
        b2az.tafl_replay(0, np.zeros((1, 1), np.uint16), np.zeros(1, np.uint32), 70000, lib=lib)  # max_turns > uint16
    # a move that lands on an occupied square is not validated by the reference either (play_move copies the source
    # layers over the destination): status stays 0; an empty SOURCE square is where the reference throws
    S = 7
    mv_occupied = (0 * S + 3) * 14 + 7 + 1  # (0,3) down to (1,3), occupied by an attacker
    r = b2az.tafl_replay(0, np.array([[mv_occupied]], np.uint16), np.array([1], np.uint32), 150, lib=lib)
    assert r["status"][0] == 0 and r["boards"][0, 1, 2].sum() == 7  # one attacker overwritten


# ---------------------------------------------------------------------------------- symmetries (eightSym)
def _check_symmetries(lib, game, n, seed):
    S, P = b2az.TAFL_DIMS[game]
    rng = np.random.default_rng(seed)
    canon = rng.random((n, P, S, S)).astype(np.float32)
    v = rng.random((n, 3)).astype(np.float32)
    pi = rng.random((n, 2 * S ** 3)).astype(np.float32)
    co, vo, po = b2az.tafl_symmetries(game, canon, v, pi, lib=lib)
    for i in range(n):
        rc, rv, rp = tafl_ref.symmetries(game, canon[i], v[i], pi[i])
        assert np.array_equal(co[i].view(np.uint32), rc.view(np.uint32)), f"{NAMES[game]} sample {i}: canonical images"
        assert np.array_equal(vo[i], rv) and np.array_equal(po[i].view(np.uint32), rp.view(np.uint32)), f"sample {i}: pi images"
    # image 0 is the sample itself; every image is a permutation of it
    assert np.array_equal(co[:, 0], canon) and np.array_equal(po[:, 0], pi)
    assert np.array_equal(np.sort(po.reshape(n, 8, -1), axis=2), np.sort(np.repeat(pi[:, None], 8, 1), axis=2))


@needs_tafl_ref
@pytest.mark.parametrize("game", [0, 1, 2])
def test_symmetries_host_build_vs_reference(game):
    _check_symmetries(b2az.load(ph.HOSTEMU_LIB), game, 3, 11 + game)


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game", [0, 1, 2])
def test_cuda_symmetries_vs_reference(game):
    _check_symmetries(None, game, 24, 21 + game)
