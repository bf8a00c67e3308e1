"""Pins the product's shared host/device headers (csrc/az_rng.h, az_math.h, az_connect4.h), compiled
for the host by tests/cpp/shared_headers_capi.cc, against libstdc++/glibc (through the oracle port)
and against the reference's own Connect4 known-answer tests (src/connect4_gs_test.cc)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import b2az
import parity_harness as ph

ROOT = ph.ROOT
SHARED = os.path.join(ROOT, "tests", "cpp", "libazshared.so")


@pytest.fixture(scope="module")
def S():
    L = C.CDLL(SHARED)
    vp, u32 = C.c_void_p, C.c_uint32
    L.azp_rng_new.restype = vp
    L.azp_rng_new.argtypes = [C.c_uint64, C.c_int, C.c_uint64]
    L.azp_rng_free.argtypes = [vp]
    L.azp_rng_u32.argtypes = [vp]
    L.azp_rng_u32.restype = u32
    L.azp_rng_shuffle.argtypes = [vp, u32, vp]
    L.azp_rng_shuffle_discard.argtypes = [vp, u32]
    L.azp_rng_uniform01.argtypes = [vp]
    L.azp_rng_uniform01.restype = C.c_float
    L.azp_rng_gamma.argtypes = [vp, C.c_float, u32, vp]
    for n in ("azp_logf", "azp_expf"):
        getattr(L, n).argtypes = [C.c_float]
        getattr(L, n).restype = C.c_float
    L.azp_powf.argtypes = [C.c_float, C.c_float]
    L.azp_powf.restype = C.c_float
    return L


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("seed,stream", [(12345, None), (0, None), (99, 0), (7, 3), (2 ** 63 + 5, 65535)])
def test_rng_matches_libstdcxx(S, seed, stream):
    Pt = ph.port_lib()
    a = S.azp_rng_new(seed, 0 if stream is None else 1, stream or 0)
    b = Pt.azo_rng_new(seed, 0 if stream is None else 1, stream or 0)
    assert [S.azp_rng_u32(a) for _ in range(256)] == [Pt.azo_rng_u32(b) for _ in range(256)]
    for rep in range(40):
        for n in range(0, 9):
            xa = np.arange(n, dtype=np.uint32)
            xb = xa.copy()
            S.azp_rng_shuffle(a, n, ph._P(xa))
            Pt.azo_rng_shuffle(b, n, ph._P(xb))
            assert np.array_equal(xa, xb), (rep, n)
        assert S.azp_rng_uniform01(a) == Pt.azo_rng_uniform01(b)
    # shuffle_discard consumes exactly what shuffle does
    for n in range(0, 9):
        S.azp_rng_shuffle_discard(a, n)
        xb = np.arange(n, dtype=np.uint32)
        Pt.azo_rng_shuffle(b, n, ph._P(xb))
        assert S.azp_rng_u32(a) == Pt.azo_rng_u32(b)
    # gamma: every alpha the Dirichlet code can ask for (10.83/k plain; shaped alphas down to 1e-6*10.83)
    for alpha in [10.83 / k for k in range(1, 8)] + [10.83 * x for x in (1e-6, 0.01, 0.0714, 0.5, 0.9)] + [1.0, 3.7]:
        ga, gb = np.zeros(64, np.float32), np.zeros(64, np.float32)
        S.azp_rng_gamma(a, C.c_float(alpha), 64, ph._P(ga))
        Pt.azo_rng_gamma(b, C.c_float(alpha), 64, ph._P(gb))
        assert np.array_equal(_bits(ga), _bits(gb)), alpha
    S.azp_rng_free(a)
    Pt.azo_rng_free(b)


def test_math_matches_glibc_quick():
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "check_az_math"), "quick"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.slow
def test_math_matches_glibc_full():
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "check_az_math"), "full"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


# ---- Connect4 known answers, restated from the reference's gtest file (src/connect4_gs_test.cc)
def _board(rows0, rows1):
    b = np.zeros((2, 6, 7), np.int8)
    for p, rows in enumerate((rows0, rows1)):
        for h, w in rows:
            b[p, h, w] = 1
    return b


KATS = [
    # (player-0 stones, player-1 stones, expected scores or None)
    ([(5, 0), (5, 1), (5, 2), (5, 3)], [(4, 0), (4, 1), (4, 2)], [1, 0, 0]),            # horizontal (:97-118)
    ([(5, 0), (4, 1)], [(5, 6), (4, 6), (3, 6), (2, 6)], [0, 1, 0]),                    # vertical
    ([(5, 0), (4, 1), (3, 2), (2, 3)], [], [1, 0, 0]),                                  # diagonal /
    ([(2, 0), (3, 1), (4, 2), (5, 3)], [], [1, 0, 0]),                                  # diagonal \
    ([(5, 3), (4, 4), (3, 5), (2, 6)], [], [1, 0, 0]),                                  # diagonal at the right edge
    ([(5, 0), (5, 1), (5, 2)], [(4, 0), (4, 1), (4, 2)], None),                         # nobody yet
    ([(5, 4), (5, 5), (5, 6)], [(4, 0)], None),                                         # 3 at the edge + wrap guard
]


def _engine_c4(boards, players, moves=None):
    return b2az.c4_batch(boards, players, np.zeros(len(players), np.uint32), moves, lib=b2az.load(ph.HOSTEMU_LIB))


def test_connect4_known_answers():
    boards = np.stack([_board(a, b) for a, b, _ in KATS])
    out = _engine_c4(boards, np.zeros(len(KATS), np.uint8))
    for i, (_, _, want) in enumerate(KATS):
        if want is None:
            assert out["terminal"][i] == 0
        else:
            assert out["terminal"][i] == 1 and list(out["scores"][i]) == want, i
    assert np.array_equal(out["boards"], boards)


def test_connect4_draw_and_full_column():
    # a full board without four in a row (columns alternate in pairs) -> draw {0,0,1} (connect4_gs.cc:121-128)
    b = np.zeros((2, 6, 7), np.int8)
    for w in range(7):
        for h in range(6):
            p = ((h // 2) + w) % 2 if w != 6 else ((h // 2) + 1) % 2
            b[p, h, w] = 1
    Pt = ph.port_lib()
    s = np.zeros(3, np.float32)
    term = Pt.azo_c4_scores(ph._P(b), ph._P(s))
    out = _engine_c4(b[None], np.zeros(1, np.uint8), np.array([3], np.uint32))
    assert out["terminal"][0] == term and list(out["scores"][0]) == list(s)
    assert out["status"][0] == -5, "playing a full column is an error (connect4_gs.cc:57)"
    assert out["valid"][0].sum() == 0


def test_connect4_random_walks_vs_port():
    Pt = ph.port_lib()
    rng = np.random.default_rng(5)
    boards, players, turns = [], [], []
    for game in range(200):
        board = np.zeros(84, np.int8)
        player, turn = C.c_uint8(0), C.c_uint32(0)
        for ply in range(rng.integers(0, 43)):
            v = np.zeros(7, np.uint8)
            Pt.azo_c4_valid(ph._P(board), ph._P(v))
            s = np.zeros(3, np.float32)
            if Pt.azo_c4_scores(ph._P(board), ph._P(s)) or v.sum() == 0:
                break
            Pt.azo_c4_play(ph._P(board), C.byref(player), C.byref(turn), int(rng.choice(np.flatnonzero(v))))
        boards.append(board.copy()); players.append(player.value); turns.append(turn.value)
    boards = np.stack(boards)
    moves = rng.integers(0, 7, len(boards)).astype(np.uint32)
    out = b2az.c4_batch(boards, np.array(players, np.uint8), np.array(turns, np.uint32), moves,
                        lib=b2az.load(ph.HOSTEMU_LIB))
    for i in range(len(boards)):
        b = boards[i].copy()
        pl, tu = C.c_uint8(players[i]), C.c_uint32(turns[i])
        rc = Pt.azo_c4_play(ph._P(b), C.byref(pl), C.byref(tu), int(moves[i]))
        assert (out["status"][i] == 0) == (rc == 0)
        assert np.array_equal(out["boards"][i].reshape(-1), b) and out["players"][i] == pl.value
        v = np.zeros(7, np.uint8); Pt.azo_c4_valid(ph._P(b), ph._P(v))
        assert np.array_equal(out["valid"][i], v)
        s = np.zeros(3, np.float32); t = Pt.azo_c4_scores(ph._P(b), ph._P(s))
        assert out["terminal"][i] == t and np.array_equal(out["scores"][i], s)
        c = np.zeros(168, np.float32); Pt.azo_c4_canonical(ph._P(b), pl, ph._P(c))
        assert np.array_equal(out["canonical"][i].reshape(-1), c)
