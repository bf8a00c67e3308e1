"""Pins the oracle port (oracle/az_oracle.cc) against the UNMODIFIED reference (oracle/_ref/libazref.so).

Skipped where the reference build is absent; tests/test_golden.py then pins the port against fixtures
generated from the reference (tools/make_golden.py)."""
import ctypes as C

import numpy as np
import pytest

import parity_harness as ph
import refdriver
from conftest import needs_ref

pytestmark = needs_ref


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _lockstep_two_oracles(G, games, visits, level, seed):
    kw = ph.level_params(level)
    a = ph.RefPM(G=G, games_to_play=games, visits=visits, eval_type=0, rng_mode=1, seed=seed, **kw)
    b = ph.PortPM(G=G, games_to_play=games, visits=visits, eval_type=0, rng_mode=1, seed=seed, **kw)
    gens = 0
    try:
        while True:
            ma, mb = a.advance(), b.advance()
            assert ma == mb, f"generation {gens}: one side finished early"
            if not ma:
                break
            ia, ca = a.build_batch()
            ib, cb = b.build_batch()
            assert np.array_equal(ia, ib), f"generation {gens}: ids"
            assert np.array_equal(_bits(ca), _bits(cb)), f"generation {gens}: canonicals"
            if gens % 5 == 0:  # peek while the reference worker is quiescent (before the answers go in)
                g = int(ia[gens % len(ia)])
                for seat in (0, 1):
                    ph.compare_peek(b.peek(g, seat), a.peek(g, seat), f"gen {gens} game {g} seat {seat}")
            v, pi = ph.fake_net(ca)
            a.update_inferences(ia, v, pi)
            b.update_inferences(ib, v, pi)
            gens += 1
        assert np.array_equal(a.scores(), b.scores())
        ma, mb = a.metrics(), b.metrics()
        for k in ma:
            assert np.float32(ma[k]) == np.float32(mb[k]), k
        ph.compare_history(b.drain_history(), a.drain_history(), ordered=True)
    finally:
        a.close()
        b.close()
    return gens


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_port_matches_reference_lockstep(level):
    gens = _lockstep_two_oracles(G=6, games=10, visits=40, level=level, seed=12345)
    assert gens > 100


def test_port_matches_reference_lockstep_100_sims():
    # BASELINE.json configs[0]: 100 sims/move, deterministic seed
    _lockstep_two_oracles(G=3, games=4, visits=100, level=1, seed=20240601)


@pytest.mark.parametrize("level", [0, 1])
def test_port_matches_reference_random_eval(level):
    kw = ph.level_params(level)
    a = ph.RefPM(G=8, games_to_play=20, visits=50, eval_type=1, rng_mode=1, seed=99, **kw)
    b = ph.PortPM(G=8, games_to_play=20, visits=50, eval_type=1, rng_mode=1, seed=99, **kw)
    try:
        a.advance()
        b.advance()
        assert a.games_completed() == b.games_completed() == 20
        assert np.array_equal(a.scores(), b.scores())
        ma, mb = a.metrics(), b.metrics()
        for k in ma:
            assert np.float32(ma[k]) == np.float32(mb[k]), k
        ph.compare_history(b.drain_history(), a.drain_history(), ordered=True)
    finally:
        a.close()
        b.close()


def test_port_no_tree_reuse_matches_reference():
    kw = ph.level_params(1)
    a = ph.RefPM(G=4, games_to_play=6, visits=30, eval_type=1, rng_mode=1, seed=5, tree_reuse=False, **kw)
    b = ph.PortPM(G=4, games_to_play=6, visits=30, eval_type=1, rng_mode=1, seed=5, tree_reuse=False, **kw)
    try:
        a.advance()
        b.advance()
        assert np.array_equal(a.scores(), b.scores())
        ph.compare_history(b.drain_history(), a.drain_history(), ordered=True)
    finally:
        a.close()
        b.close()


def test_port_rng_matches_reference_rng():
    R, Pt = refdriver.lib(), ph.port_lib()
    for seed, stream in [(12345, None), (0, None), (7, 3), (2 ** 63 + 5, 65535)]:
        ra = R.azref_rng_new(seed, 0 if stream is None else 1, stream or 0)
        rb = Pt.azo_rng_new(seed, 0 if stream is None else 1, stream or 0)
        assert [R.azref_rng_u32(ra) for _ in range(64)] == [Pt.azo_rng_u32(rb) for _ in range(64)]
        for n in range(0, 9):
            xa = np.arange(n, dtype=np.uint32)
            xb = xa.copy()
            R.azref_rng_shuffle(ra, n, refdriver.P(xa))
            Pt.azo_rng_shuffle(rb, n, ph._P(xb))
            assert np.array_equal(xa, xb)
        assert R.azref_rng_uniform01(ra) == Pt.azo_rng_uniform01(rb)
        ga, gb = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.azref_rng_gamma(ra, C.c_float(10.83 / 7), 16, refdriver.P(ga))
        Pt.azo_rng_gamma(rb, C.c_float(10.83 / 7), 16, ph._P(gb))
        assert np.array_equal(_bits(ga), _bits(gb))
        R.azref_rng_free(ra)
        Pt.azo_rng_free(rb)


def test_port_connect4_matches_reference_random_walks():
    R, Pt = refdriver.lib(), ph.port_lib()
    rng = np.random.default_rng(3)
    for game in range(60):
        gs = R.azref_c4_new()
        board = np.zeros(84, np.int8)
        player = C.c_uint8(0)
        turn = C.c_uint32(0)
        for ply in range(43):
            va, vb = np.zeros(7, np.uint8), np.zeros(7, np.uint8)
            R.azref_c4_valid(gs, refdriver.P(va))
            Pt.azo_c4_valid(ph._P(board), ph._P(vb))
            assert np.array_equal(va, vb)
            sa, sb = np.zeros(3, np.float32), np.zeros(3, np.float32)
            ta = R.azref_c4_scores(gs, refdriver.P(sa))
            tb = Pt.azo_c4_scores(ph._P(board), ph._P(sb))
            assert ta == tb and np.array_equal(sa, sb)
            ca, cb = np.zeros(168, np.float32), np.zeros(168, np.float32)
            R.azref_c4_canonical(gs, refdriver.P(ca))
            Pt.azo_c4_canonical(ph._P(board), player, ph._P(cb))
            assert np.array_equal(ca, cb)
            if ta:
                break
            mv = int(rng.choice(np.flatnonzero(va)))
            assert R.azref_c4_play(gs, mv) == 0
            assert Pt.azo_c4_play(ph._P(board), C.byref(player), C.byref(turn), mv) == 0
        # a full column must be rejected by both
        full = np.flatnonzero(va == 0)
        if len(full) and not ta:
            assert R.azref_c4_play(gs, int(full[0])) != 0
            assert Pt.azo_c4_play(ph._P(board), C.byref(player), C.byref(turn), int(full[0])) != 0
        R.azref_c4_free(gs)
