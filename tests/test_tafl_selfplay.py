"""PlayManager::play over the tafl games on the device (b2az_tafl_selfplay_*) against the UNMODIFIED reference
PlayManager (oracle/_ref/libazref_tafl.so: azref_tafl_selfplay — play_manager.cc + mcts.cc + the tafl games compiled in
place). Slot g of the device run == a reference PlayManager with concurrent_games = 1, games_to_play = games_per_slot,
EvalType::RANDOM, run on one thread after MCTS::seed_thread_rng(seed + g). Compared bit for bit (floats by bit
pattern): the training samples in history_ order (canonical planes, outcome, policy target — Gumbel improved policy /
probs_pruned(1) / probs(1)), scores, games completed, game length, and the metric sums (leaf depth, root entropy,
legal moves). Golden CRCs (tools/make_golden_tafl_selfplay.py) cover the same cases without the reference."""
import os
import zlib

import numpy as np
import pytest

import b2az
import parity_harness as ph
import tafl_ref

NAMES = {0: "brandubh", 1: "opentafl", 2: "tawlbwrdd"}
needs_tafl_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")
GOLDEN = os.path.join(ph.ROOT, "tests", "golden", "tafl_selfplay.npz")

# self-play settings of configs/brandubh.yaml-style runs, small: name -> (game, slots, games per slot, max_turns, visits, kwargs)
CASES = {
    "brandubh_puct_plain": (0, 6, 2, 40, 40, dict()),
    "brandubh_puct_selfplay": (0, 6, 2, 40, 40, dict(epsilon=0.25, root_policy_temp=1.25, shaped_dirichlet=True,
                                                     policy_target_pruning=True, root_fpu_zero=True, start_temp=1.0,
                                                     final_temp=0.2, temp_decay_half_life=10.0)),
    "brandubh_gumbel": (0, 6, 2, 40, 48, dict(gumbel_m=16, root_policy_temp=1.25)),
    # epsilon > 0 with Gumbel: no noise at the first root evaluation (mcts.cc:514-518) but PlayManager still re-noises the
    # reused root after every move (play_manager.cc:546-553)
    "brandubh_gumbel_with_noise": (0, 4, 1, 40, 48, dict(gumbel_m=16, epsilon=0.25, root_policy_temp=1.25, shaped_dirichlet=True)),
    "brandubh_no_tree_reuse": (0, 4, 2, 30, 32, dict(tree_reuse=False, epsilon=0.25, gumbel_m=8)),
    "opentafl_gumbel": (1, 3, 1, 24, 40, dict(gumbel_m=16, cpuct=2.0)),
    "tawlbwrdd_puct_selfplay": (2, 3, 1, 24, 40, dict(epsilon=0.25, root_policy_temp=1.1, policy_target_pruning=True,
                                                      start_temp=1.0, final_temp=0.5, temp_decay_half_life=6.0)),
}


def run_device(game, slots, per_slot, max_turns, visits, kw, seed):
    words = 2 * (1 + (max_turns + 2) * visits * (1 + 8 * (64 if game == 0 else 200)))
    sp = b2az.TaflSelfplay(game, slots, max_turns, visits, games_per_slot=per_slot, seed=seed, words_per_tree=words,
                           hist_capacity=slots * per_slot * max_turns, **kw)  # drained once, at the end
    active, rounds = slots, 0
    while active:
        active = sp.play(8)
        rounds += 1
        assert rounds < 10000
    canon, v, pi, slot = sp.drain_history()
    st, err = sp.slots()
    sp.close()
    assert (err == 0).all() and (st["error"] == 0).all()
    return canon, v, pi, slot, st


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def f32(x):
    return np.float32(x)


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_selfplay_equals_the_reference_playmanager(name):
    game, slots, per_slot, max_turns, visits, kw = CASES[name]
    seed = 1000 + 17 * sorted(CASES).index(name)
    canon, v, pi, slot, st = run_device(game, slots, per_slot, max_turns, visits, kw, seed)
    for g in range(slots):
        ref = tafl_ref.selfplay(game, seed + g, max_turns, per_slot, visits, **kw)
        rows = slot == g
        assert rows.sum() == len(ref["v"]), f"{name} slot {g}: {rows.sum()} samples vs {len(ref['v'])}"
        assert np.array_equal(canon[rows].view(np.uint32), ref["canonical"].view(np.uint32)), f"{name} slot {g}: canonical"
        assert np.array_equal(v[rows].view(np.uint32), ref["v"].view(np.uint32)), f"{name} slot {g}: outcomes"
        assert np.array_equal(pi[rows].view(np.uint32), ref["pi"].view(np.uint32)), f"{name} slot {g}: policy targets"
        s = st[g]
        assert s["games_completed"] == ref["games_completed"] == per_slot and s["active"] == 0
        assert np.array_equal(s["scores"], ref["scores"])
        # the getters of play_manager.h:288-316 on the slot's sums
        assert f32(f32(s["game_length"]) / f32(s["games_completed"])) == ref["avg_game_length"]
        assert f32(s["leaf_depth"] / float(s["total_full_move_count"])) == ref["avg_leaf_depth"]
        assert f32(s["entropy"] / float(s["total_full_move_count"])) == ref["avg_search_entropy"]
        assert f32(s["valid_moves"] / float(s["total_move_count"])) == ref["avg_valid_moves"]
        assert s["simulations"] == visits * s["total_move_count"]


# Per-seat budgets, playout-cap randomisation and resignation against the unmodified reference. The reference flips its
# playout-cap / playthrough coins with an unseedable engine (play_manager.cc:261-262), so the bit-exact cases sit at
# the deterministic corners: every search capped (percent 1) or none (0), every resignation honoured (playthrough 0) or
# none (1). In between only the coin VALUES differ (this engine has its own coin stream per slot, apart from the search's).
FEATURE_CASES = {
    "brandubh_seat_visits": (0, 5, 2, 40, dict(seat_visits=(48, 24), epsilon=0.25, root_policy_temp=1.25, policy_target_pruning=True)),
    "brandubh_all_capped_puct": (0, 4, 2, 40, dict(playout_cap_randomization=True, playout_cap_percent=1.0, playout_cap_depth=12,
                                                   epsilon=0.25, start_temp=1.0, final_temp=0.3, temp_decay_half_life=8.0)),
    "brandubh_all_capped_gumbel_fast": (0, 4, 2, 40, dict(playout_cap_randomization=True, playout_cap_percent=1.0,
                                                          seat_cap_visits=(16, 20), gumbel_m=8, fast_search_uses_gumbel=True)),
    "brandubh_all_capped_gumbel_puct_fast": (0, 4, 1, 40, dict(playout_cap_randomization=True, playout_cap_percent=1.0,
                                                               playout_cap_depth=14, gumbel_m=8, epsilon=0.25)),
    "brandubh_never_capped": (0, 4, 1, 40, dict(playout_cap_randomization=True, playout_cap_percent=0.0, playout_cap_depth=12,
                                                gumbel_m=16)),
    "brandubh_resign_always": (0, 6, 2, 60, dict(resign_percent=0.62, resign_playthrough_percent=0.0, epsilon=0.25)),
    "brandubh_resign_playthrough_always": (0, 4, 1, 40, dict(resign_percent=0.62, resign_playthrough_percent=1.0)),
    "tawlbwrdd_seat_visits_resign": (2, 3, 1, 24, dict(seat_visits=(40, 28), resign_percent=0.62, resign_playthrough_percent=0.0)),
}


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("name", sorted(FEATURE_CASES))
def test_selfplay_features_equal_the_reference_at_the_deterministic_corners(name):
    game, slots, per_slot, max_turns, kw = FEATURE_CASES[name]
    visits = 40
    seed = 4000 + 13 * sorted(FEATURE_CASES).index(name)
    canon, v, pi, slot, st = run_device(game, slots, per_slot, max_turns, visits, kw, seed)
    resigned = 0.0
    for g in range(slots):
        ref = tafl_ref.selfplay(game, seed + g, max_turns, per_slot, visits, **kw)
        rows = slot == g
        assert rows.sum() == len(ref["v"]), f"{name} slot {g}: {rows.sum()} samples vs {len(ref['v'])}"
        assert np.array_equal(canon[rows].view(np.uint32), ref["canonical"].view(np.uint32)), f"{name} slot {g}: canonical"
        assert np.array_equal(v[rows].view(np.uint32), ref["v"].view(np.uint32)), f"{name} slot {g}: outcomes"
        assert np.array_equal(pi[rows].view(np.uint32), ref["pi"].view(np.uint32)), f"{name} slot {g}: policy targets"
        s = st[g]
        assert s["games_completed"] == ref["games_completed"] == per_slot and s["active"] == 0
        assert np.array_equal(s["scores"], ref["scores"]) and np.array_equal(s["resign_scores"], ref["resign_scores"])
        assert f32(f32(s["game_length"]) / f32(s["games_completed"])) == ref["avg_game_length"]
        if s["total_full_move_count"]:
            assert f32(s["leaf_depth"] / float(s["total_full_move_count"])) == ref["avg_leaf_depth"]
            assert f32(s["entropy"] / float(s["total_full_move_count"])) == ref["avg_search_entropy"]
        if s["total_fast_move_count"]:
            assert f32(s["fast_leaf_depth"] / float(s["total_fast_move_count"])) == ref["fast_avg_leaf_depth"]
            assert f32(s["fast_entropy"] / float(s["total_fast_move_count"])) == ref["fast_avg_search_entropy"]
        assert f32(s["valid_moves"] / float(s["total_move_count"])) == ref["avg_valid_moves"]
        resigned += float(s["resign_scores"].sum())
    if "all_capped" in name:
        assert len(v) == 0 and (st["total_fast_move_count"] == st["total_move_count"]).all(), "capped searches record nothing"
    if name == "brandubh_resign_always":
        assert resigned > 0, "pick a resign_percent that triggers"
    if "playthrough_always" in name:
        assert resigned == 0 and (st["playthrough"] == 1).any()


@pytest.mark.gpu
def test_selfplay_playout_cap_mix_statistics():
    """Between the corners the coins are this engine's own: 75 % of the searches are fast ones (no sample, cap budget),
    and the simulation count is exactly what the full / fast move counts say."""
    game, slots, max_turns, visits, cap = 0, 256, 40, 40, 10
    kw = dict(playout_cap_randomization=True, playout_cap_percent=0.75, playout_cap_depth=cap, epsilon=0.25, gumbel_m=0)
    canon, v, pi, slot, st = run_device(game, slots, 1, max_turns, visits, kw, 99)
    full, fast, total = int(st["total_full_move_count"].sum()), int(st["total_fast_move_count"].sum()), int(st["total_move_count"].sum())
    assert full + fast == total and 0.70 < fast / total < 0.80
    assert len(v) == full and int(st["simulations"].sum()) == full * visits + fast * cap


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_selfplay_equals_the_golden_fixture(name):
    game, slots, per_slot, max_turns, visits, kw = CASES[name]
    seed = 1000 + 17 * sorted(CASES).index(name)
    gold = np.load(GOLDEN)
    canon, v, pi, slot, st = run_device(game, slots, per_slot, max_turns, visits, kw, seed)
    for g in range(slots):
        rows = slot == g
        want = gold[name][g]
        got = (rows.sum(), crc(canon[rows]), crc(v[rows]), crc(pi[rows]), int(st[g]["game_length"]), crc(st[g]["scores"]))
        assert tuple(int(x) for x in want) == tuple(int(x) for x in got), f"{name} slot {g}"


@pytest.mark.gpu
@needs_tafl_ref
def test_selfplay_with_a_host_evaluator_in_the_middle():
    """EvalType::NN form: find_leaf -> evaluator -> process_result per simulation. With the evaluator answering
    dumb_eval's numbers (uniform over the legal moves of the leaf, 1/3 each) the run equals the fused RANDOM one."""
    game, slots, max_turns, visits = 0, 4, 30, 24
    words = 2 * (1 + (max_turns + 2) * visits * (1 + 8 * 64))
    kw = dict(epsilon=0.25, root_policy_temp=1.25, policy_target_pruning=True)
    fused = run_device(game, slots, 1, max_turns, visits, kw, 77)
    import ctypes as C
    cudart = C.CDLL("libcudart.so.12")
    sp = b2az.TaflSelfplay(game, slots, max_turns, visits, games_per_slot=1, seed=77, words_per_tree=words, **kw)
    S, P, A = sp.S, sp.P, sp.A
    active, steps = slots, 0
    canon = np.zeros((slots, P, S, S), np.float32)
    while active:
        ptr = sp.find_leaf()
        assert cudart.cudaMemcpy(canon.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), canon.nbytes, 2) == 0  # D2H, syncs
        vs, pis = np.zeros((slots, 3), np.float32), np.zeros((slots, A), np.float32)
        for g in range(slots):
            board = np.zeros((3, S, S), np.int8)
            board[:] = canon[g, :3] > 0
            player = 0 if canon[g, 3, 0, 0] > 0 else 1
            pos = tafl_ref.position(game, board, player, 0, max_turns, 0)
            valid = pos["valid"].astype(np.float32)
            # dumb_eval (game_state.h:160-173): Vector<uint8_t>::sum() wraps mod 256
            total = np.float32(int(valid.sum()) % 256)
            pis[g] = valid / total if total > 0 else valid
            vs[g] = np.float32(1.0 / 3.0)
        active = sp.process_result(vs, pis, want_active=True)
        steps += 1
        assert steps < 200000
    canon, v, pi, slot = sp.drain_history()
    sp.close()
    # outcomes and sample counts: the same games were played (policy targets too, when no leaf had > 255 legal moves)
    assert np.array_equal(slot, fused[3]) and np.array_equal(v, fused[1])
    assert np.array_equal(canon.view(np.uint32), fused[0].view(np.uint32))
    assert np.array_equal(pi.view(np.uint32), fused[2].view(np.uint32))


@pytest.mark.gpu
def test_sharded_slots_equal_the_unsharded_run():
    """Multi-GPU sharding: slots are independent, rank r takes slots [r*n, (r+1)*n) with seed + r*n — the
    union of two half-size runs is the full run, sample for sample."""
    game, slots, max_turns, visits = 0, 8, 30, 32
    kw = dict(gumbel_m=16)
    full = run_device(game, slots, 1, max_turns, visits, kw, 500)
    for r in range(2):
        part = run_device(game, slots // 2, 1, max_turns, visits, kw, 500 + r * (slots // 2))
        for g in range(slots // 2):
            a, b = full[3] == g + r * (slots // 2), part[3] == g
            assert a.sum() == b.sum() > 0
            for k in range(3):
                assert np.array_equal(full[k][a].view(np.uint32), part[k][b].view(np.uint32))
            assert full[4][g + r * (slots // 2)]["game_length"] == part[4][g]["game_length"]


def test_selfplay_fails_loudly_without_a_gpu_and_on_bad_arguments():
    from conftest import has_cuda
    with pytest.raises(b2az.B2azError):
        b2az.TaflSelfplay(0, 0, 30, 8)  # no slots
    with pytest.raises(b2az.B2azError):
        b2az.TaflSelfplay(0, 2, 30, 0)  # no visits
    if not has_cuda():
        with pytest.raises(b2az.B2azError) as ei:
            b2az.TaflSelfplay(0, 2, 30, 8)
        assert "no CPU fallback" in str(ei.value)


def _nn_run(game, slots, max_turns, visits, seed, cache_entries, kw):
    """EvalType::NN form driven with a deterministic pure-function evaluator (a function of the canonical planes only)."""
    import ctypes as C
    import zlib

    cudart = C.CDLL("libcudart.so.12")
    sp = b2az.TaflSelfplay(game, slots, max_turns, visits, games_per_slot=1, seed=seed, cache_entries=cache_entries,
                           hist_capacity=slots * max_turns, **kw)
    S, P, A = sp.S, sp.P, sp.A
    canon = np.zeros((slots, P, S, S), np.float32)
    active, steps = slots, 0
    while active:
        ptr = sp.find_leaf()
        assert cudart.cudaMemcpy(canon.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), canon.nbytes, 2) == 0
        vs, pis = np.zeros((slots, 3), np.float32), np.zeros((slots, A), np.float32)
        for g in range(slots):
            rng = np.random.default_rng(zlib.crc32(canon[g].tobytes()))
            v = rng.random(3).astype(np.float32) + np.float32(0.05)
            vs[g] = v / v.sum()
            pi = rng.random(A).astype(np.float32) ** 4 + np.float32(1e-3)
            pis[g] = pi / pi.sum()
        active = sp.process_result(vs, pis, want_active=True)
        steps += 1
        assert steps < 400000
    out = sp.drain_history()
    st = sp.stats()
    slots_, err = sp.slots()
    sp.close()
    assert (err == 0).all() and (slots_["error"] == 0).all()
    return out, st, steps


@pytest.mark.gpu
@pytest.mark.parametrize("game,slots,max_turns,visits,kw", [
    (0, 5, 40, 32, dict(epsilon=0.25, root_policy_temp=1.25, policy_target_pruning=True)),
    (1, 3, 16, 24, dict(gumbel_m=8)),
    (10, 3, 600, 16, dict(gumbel_m=8)),
])
def test_position_cache_on_equals_cache_off(game, slots, max_turns, visits, kw):
    """The device position cache of the wide-tree engine (az_selfplay.h SpCache; S3FIFOCache's role for the tafl games and
    Star Gambit): with a pure-function evaluator a run with the cache equals the run without it, sample for sample — the
    reference's own criterion for a cache (test_cache.py:227-253) — while the evaluator is asked less often."""
    def same(x, y):  # slot by slot: the order in which slots finish (and append their samples) is not part of the result
        for g in range(slots):
            rx, ry = x[3] == g, y[3] == g
            assert rx.sum() == ry.sum() and rx.sum() > 0
            for a, b in zip(x[:3], y[:3]):
                assert np.array_equal(a[rx].view(np.uint32), b[ry].view(np.uint32)), g

    off, st0, steps0 = _nn_run(game, slots, max_turns, visits, 321, 0, kw)
    on, st1, steps1 = _nn_run(game, slots, max_turns, visits, 321, 1 << 14, kw)
    same(off, on)
    assert st0.cache_hits == 0 and st0.cache_max_size == 0
    assert st1.cache_hits > 0 and st1.cache_misses > 0 and st1.cache_max_size == 1 << 14
    assert st1.simulations == st0.simulations and st1.cache_hits + st1.cache_misses == st1.simulations
    assert steps1 <= steps0
    tiny, st2, _ = _nn_run(game, slots, max_turns, visits, 321, 16, kw)  # 16 entries: evictions all the time, same games
    same(off, tiny)
    assert st2.cache_evictions > 0
