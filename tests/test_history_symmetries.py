"""b2az_drain_history_sym = build_history_batch (py_wrapper.cc:393-424) followed by
game_runner.exploit_symmetries (game_runner.py:1050-1144) on the device: every training sample comes out together
with its images under GameState::symmetries, in that order (Connect4: the sample, then its mirror image,
connect4_gs.cc:151-170). Checked against the plain drain of an identical run and against the unmodified
reference's Connect4GS::symmetries."""
import ctypes as C

import numpy as np
import pytest

import b2az
import parity_harness as ph
import refdriver
from conftest import needs_ref


def _run(lib, sym):
    e = ph.make_engine(lib, G=32, games_to_play=48, visits=24, eval_type=b2az.EVAL_RANDOM, rng_mode=b2az.RNG_PER_GAME,
                       seed=5, history_capacity=1 << 14, **ph.level_params(1))
    while e.stats().active_games:
        e.step(64)
    out = e.drain_history_sym(1 << 13) if sym else e.drain_history(1 << 13)
    e.close()
    return out


def _check(lib):
    c0, v0, p0 = _run(lib, False)
    c2, v2, p2 = _run(lib, True)
    n = len(c0)
    assert n > 300 and len(c2) == 2 * n
    bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
    # even rows: the samples themselves, in the same order as the plain drain
    assert np.array_equal(bits(c2[0::2]), bits(c0)) and np.array_equal(bits(v2[0::2]), bits(v0))
    assert np.array_equal(bits(p2[0::2]), bits(p0))
    # odd rows: the mirror images
    assert np.array_equal(bits(c2[1::2]), bits(c0[:, :, :, ::-1])) and np.array_equal(bits(v2[1::2]), bits(v0))
    assert np.array_equal(bits(p2[1::2]), bits(p0[:, ::-1]))
    return c0, v0, p0, c2, v2, p2


def test_symmetric_drain_host_build():
    _check(ph.HOSTEMU_LIB)


@needs_ref
def test_mirror_rows_equal_reference_symmetries():
    c0, v0, p0, c2, v2, p2 = _check(ph.HOSTEMU_LIB)
    L = refdriver.lib()
    gs = L.azref_c4_new()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for i in range(0, len(c0), 7):
        co, vo, po = np.zeros((4, 6, 7), np.float32), np.zeros(3, np.float32), np.zeros(7, np.float32)
        ci, vi, pi = (np.ascontiguousarray(x[i]) for x in (c0, v0, p0))
        L.azref_c4_mirror(gs, p(ci), p(vi), p(pi), p(co), p(vo), p(po))
        assert np.array_equal(co.view(np.uint32), np.ascontiguousarray(c2[2 * i + 1]).view(np.uint32))
        assert np.array_equal(vo, v2[2 * i + 1]) and np.array_equal(po.view(np.uint32), np.ascontiguousarray(p2[2 * i + 1]).view(np.uint32))
    L.azref_c4_free(gs)


@pytest.mark.gpu
def test_symmetric_drain_cuda():
    _check(None)
