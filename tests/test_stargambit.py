"""Star Gambit game kernels (csrc/az_stargambit_kernels.h, b2az_sg_replay: one warp per game) and the host classes against
golden transcripts generated from the UNMODIFIED reference (tools/make_golden_stargambit.py ->
tests/golden/stargambit_transcripts.npz): after every move player, turn, terminal code, number of legal moves and the
CRC-32 of the legal-move mask, of the canonical planes (float32 bytes: bit-exact) and of the serialised unit list."""
import os
import zlib

import numpy as np
import pytest

import parity_harness as ph
from conftest import has_cuda

GOLD = os.path.join(ph.ROOT, "tests", "golden", "stargambit_transcripts.npz")
GAMES = [10, 11, 12, 13, 20, 21, 22, 23]
CLS = {10: "StarGambitSkirmishGS", 11: "StarGambitShowdownGS", 12: "StarGambitClashGS", 13: "StarGambitBattleGS"}


def state_blob(row):
    """device SGState record -> the reference's to_bytes() prefix [num_units | units | reserves | player | turn | flags]"""
    b = row.tobytes()
    nu = b[180]
    return (int(nu).to_bytes(4, "little") + b[: 9 * nu] + b[181:189] + b[189:190] + b[196:200] + b[190:191] + b[191:192] + b[192:193])


@pytest.mark.parametrize("game", GAMES)
def test_host_classes_match_golden(game):
    from test_pybind_module import module

    az = module("emu")
    g = np.load(GOLD)
    moves, lens, meta, crcs = (g[f"g{game}_{k}"] for k in ("moves", "lens", "meta", "crcs"))
    for i in range(len(lens)):
        gs = getattr(az, CLS[game])() if game < 20 else az.StarGambitUnifiedGS(game - 20)
        for k in range(int(lens[i]) + 1):
            if k:
                gs.play_move(int(moves[i, k - 1]))
            vm = np.asarray(gs.valid_moves())
            sc = gs.scores()
            term = 0 if sc is None else (1 + int(np.argmax(np.asarray(sc))) if np.asarray(sc).max() == 1.0 else 4)
            assert (gs.current_player(), gs.current_turn(), term, int(vm.sum())) == tuple(int(x) for x in meta[i, k]), (i, k)
            assert zlib.crc32(vm.tobytes()) == crcs[i, k, 0], (i, k)
            if k % 3 == 0 or k == lens[i]:
                assert zlib.crc32(np.ascontiguousarray(gs.canonicalized()).tobytes()) == crcs[i, k, 1], (i, k)


@pytest.mark.gpu
@pytest.mark.parametrize("game", GAMES)
def test_replay_kernel_matches_golden(game):
    if not has_cuda():
        pytest.skip("no CUDA device")
    import b2az

    g = np.load(GOLD)
    moves, lens, meta, crcs = (g[f"g{game}_{k}"] for k in ("moves", "lens", "meta", "crcs"))
    out = b2az.sg_replay(game, moves, lens)
    assert not out["status"].any()
    for i in range(len(lens)):
        for k in range(int(lens[i]) + 1):
            st = out["states"][i, k]
            got = (int(st[189]), int(np.frombuffer(st[196:200].tobytes(), np.uint32)[0]), int(out["terminal"][i, k]),
                   int(out["n_valid"][i, k]))
            assert got == tuple(int(x) for x in meta[i, k]), (i, k)
            assert int(out["valid"][i, k].sum()) == got[3], (i, k)
            assert zlib.crc32(out["valid"][i, k].tobytes()) == crcs[i, k, 0], (i, k)
            assert zlib.crc32(out["canonical"][i, k].tobytes()) == crcs[i, k, 1], (i, k)
            assert zlib.crc32(state_blob(st)) == crcs[i, k, 2], (i, k)


@pytest.mark.gpu
def test_replay_kernel_many_games_equal_their_first_copy():
    """full-size launch: 4,096 games (256 distinct transcripts repeated) — every copy must equal the first one"""
    if not has_cuda():
        pytest.skip("no CUDA device")
    import b2az

    g = np.load(GOLD)
    moves, lens = g["g23_moves"][:, :96], np.minimum(g["g23_lens"], 96)
    reps = 512
    out = b2az.sg_replay(23, np.tile(moves, (reps, 1)), np.tile(lens, reps), want_valid=False)
    n = len(lens)
    c = out["canonical"].reshape(reps, n, *out["canonical"].shape[1:])
    assert np.array_equal(c, np.broadcast_to(c[:1], c.shape))
    assert np.array_equal(out["n_valid"].reshape(reps, n, -1), np.broadcast_to(out["n_valid"].reshape(reps, n, -1)[:1], (reps, n, 97)))


@pytest.mark.gpu
@pytest.mark.parametrize("game", [10, 13, 22])
def test_symmetry_kernel_matches_reference(game):
    """b2az_sg_symmetries (identity + NW-axis mirror on the device, float32 and the fp16 packing of save_compressed) against
    GameState::symmetries of the unmodified reference."""
    if not has_cuda():
        pytest.skip("no CUDA device")
    import b2az
    import tafl_ref

    if not tafl_ref.available():
        pytest.skip("oracle/_ref/libazref_tafl.so not built")
    D, A, P = b2az.sg_dims(game)
    rng = np.random.default_rng(game)
    n = 5
    canon = rng.random((n, P, D, D), dtype=np.float32)
    v = rng.random((n, 3), dtype=np.float32)
    pi = rng.random((n, A), dtype=np.float32)
    co, vo, po = b2az.sg_symmetries(game, canon, v, pi)
    ch, vh, ph_ = b2az.sg_symmetries(game, canon, v, pi, fp16=True)
    for i in range(n):
        rc, rv, rp = tafl_ref.symmetries(game, canon[i], v[i], pi[i])
        assert np.array_equal(co[i].view(np.uint32), rc.view(np.uint32)), i
        assert np.array_equal(vo[i], rv) and np.array_equal(po[i].view(np.uint32), rp.view(np.uint32)), i
        assert np.array_equal(ch[i], rc.astype(np.float16)) and np.array_equal(ph_[i], rp.astype(np.float16)), i
