"""Seat permutations, two model groups and mixed evaluator types (PlayParams::seat_perms / model_groups / eval_type:
play_manager.cc:24-90, 213-221, 466-474, 577-598) — what game_runner.play_past (game_runner.py:2183-2260) and
tournament.py build — on the Connect4 engine, pinned to the UNMODIFIED reference PlayManager.

The engine hands the permutations out per slot (slot g plays permutation g % n in every one of its games: the
reference's round-robin hand-out when the slots finish in order), so slot g must equal a reference PlayManager with
concurrent_games = 1, seat_perms = [perms[g % n]], seeded MCTS::seed_thread_rng(seed + g): training samples, scores and
the per-permutation score tables. The two groups search with different budgets (mcts_visits is per group:
seat_visits_[p][s] = mcts_visits_[seat_perms_[p][s]], play_manager.cc:70-80) and, in the NN cases, different
evaluators, so a row routed to the wrong group changes the games.

CPU: the host-emulation build of the engine logic (tests/cpp). GPU: the product library."""
import os

import numpy as np
import pytest

import parity_harness as ph
import refdriver

b2az = ph.b2az

pytestmark = pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref/libazref.so not built")

PERMS = [[0, 1], [1, 0]]
RANDOM, NN = b2az.EVAL_RANDOM, b2az.EVAL_NN


def net_of_group(group, canon):
    """group 0: fake_net; group 1: fake_net with the value and policy rows rolled — a different 'network'."""
    v, pi = ph.fake_net(canon)
    if group == 1:
        v, pi = np.roll(v, 1, axis=1), np.roll(pi, 3, axis=1)
    return np.ascontiguousarray(v), np.ascontiguousarray(pi)


def run_engine(engine_lib, G, quota, group_visits, evals, seed, level):
    lib = b2az.load(engine_lib) if engine_lib else b2az.load()
    all_random = all(e == RANDOM for e in evals)
    kw = ph.level_params(level)
    p = b2az.default_params(lib, games_to_play=G * quota, concurrent_games=G, mcts_visits=group_visits, history_enabled=1,
                            self_play=0, tree_reuse=1, eval_type=RANDOM if all_random else NN, rng_mode=b2az.RNG_PER_GAME,
                            seed=seed, per_slot_quota=1, history_capacity=max(1 << 12, G * quota * 42), **kw)
    b2az.fill_perms(p, PERMS, [[group_visits[g] for g in perm] for perm in PERMS], None,
                    None if all_random else [e == RANDOM for e in evals])
    eng = b2az.Engine(p, lib=lib)
    rows_by_group = [0, 0]
    try:
        for _ in range(10 ** 6):
            if all_random:
                eng.step(64)
            else:
                eng.step(1)
                ids, canon = eng.leaf_batch_host()
                if len(ids):
                    groups = eng.leaf_groups_host(len(ids))
                    v = np.empty((len(ids), 3), np.float32)
                    pi = np.empty((len(ids), 7), np.float32)
                    for g in (0, 1):
                        m = groups == g
                        if m.any():
                            assert evals[g] == NN, "a RANDOM group's leaf reached the evaluator batch"
                            v[m], pi[m] = net_of_group(g, canon[m])
                            rows_by_group[g] += int(m.sum())
                    eng.submit_eval_host(ids, v, pi)
            st = eng.stats()
            if st.active_games == 0:
                break
        assert st.device_error == 0 and st.games_completed == G * quota
        return eng.drain_history(1 << 20), st, eng.perm_scores(), rows_by_group
    finally:
        eng.close()


def run_reference_slot(g, quota, group_visits, evals, seed, level):
    kw = ph.level_params(level)
    cfg = refdriver.play_cfg(games_to_play=quota, concurrent_games=1, max_batch_size=1, mcts_visits=group_visits,
                             history_enabled=1, self_play=0, tree_reuse=1, **kw)
    cfg.has_groups = 1
    cfg.model_groups[0], cfg.model_groups[1] = 0, 1
    cfg.n_seat_perms = 1
    cfg.seat_perms[0][0], cfg.seat_perms[0][1] = PERMS[g % len(PERMS)]
    cfg.group_eval[0], cfg.group_eval[1] = evals
    pm = refdriver.RefPlayManager(cfg)
    try:
        if all(e == RANDOM for e in evals):
            pm.play_here(seed + g)
        else:
            pm.start_workers(1, seed + g, True)
            for _ in range(10 ** 7):
                st = pm.wait_quiescent(60000)
                assert st >= 0, "reference lock-step harness timed out"
                if st == 0:
                    break
                for grp in (0, 1):
                    if evals[grp] != NN:
                        continue
                    ids, canon = pm.build_batch(grp, 1)
                    if len(ids):
                        v, pi = net_of_group(grp, canon)
                        pm.update_inferences(ids, v, pi, group=grp)
            pm.join()
        hist = pm.drain_history(quota * 42 + 1)
        return hist, pm.scores(), pm.perm_scores(0)
    finally:
        pm.close()


def check(engine_lib, G, quota, group_visits, evals, seed, level):
    he, st, perm, rows = run_engine(engine_lib, G, quota, group_visits, evals, seed, level)
    canon, v, pi = [], [], []
    scores = np.zeros(3, np.float64)
    perm_scores = np.zeros((len(PERMS), 3), np.float64)
    perm_games = [0] * len(PERMS)
    for g in range(G):
        h, sc, (psc, pn) = run_reference_slot(g, quota, group_visits, evals, seed, level)
        canon.append(h[0]); v.append(h[1]); pi.append(h[2])
        scores += sc
        perm_scores[g % len(PERMS)] += psc
        perm_games[g % len(PERMS)] += pn
    ph.compare_history(he, (np.concatenate(canon), np.concatenate(v), np.concatenate(pi)), ordered=False)
    assert np.array_equal(np.array(st.scores[:], np.float64), scores)
    assert len(perm) == len(PERMS)
    for p in range(len(PERMS)):
        assert np.array_equal(perm[p][0].astype(np.float64), perm_scores[p]) and perm[p][1] == perm_games[p] == G // len(PERMS) * quota
    for g in (0, 1):
        assert (rows[g] > 0) == (evals[g] == NN and not all(e == RANDOM for e in evals))
    return len(he[0])


CASES = [
    pytest.param((RANDOM, RANDOM), 0, id="random-random-puct"),
    pytest.param((NN, RANDOM), 0, id="nn-random-puct"),       # play_past against iteration 0 (RandPlayer)
    pytest.param((NN, NN), 1, id="nn-nn-connect4-yaml"),      # play_past: two networks, root noise + temperature
    pytest.param((RANDOM, NN), 1, id="random-nn-connect4-yaml"),
]


@pytest.mark.skipif(not os.path.exists(ph.HOSTEMU_LIB), reason="host-emulation library not built")
@pytest.mark.parametrize("evals,level", CASES)
def test_seat_perms_hostemu_vs_reference(evals, level):
    assert check(ph.HOSTEMU_LIB, G=4, quota=2, group_visits=(24, 12), evals=evals, seed=4242, level=level) > 40


@pytest.mark.gpu
@pytest.mark.parametrize("evals,level", CASES)
def test_seat_perms_gpu_vs_reference(evals, level):
    assert check(None, G=8, quota=2, group_visits=(40, 20), evals=evals, seed=777, level=level) > 100


def test_seat_perm_validation():
    lib = b2az.load(ph.HOSTEMU_LIB) if os.path.exists(ph.HOSTEMU_LIB) else None
    if lib is None:
        pytest.skip("host-emulation library not built")
    p = b2az.default_params(lib, games_to_play=3, concurrent_games=3, mcts_visits=(8, 8), eval_type=RANDOM)
    b2az.fill_perms(p, PERMS)
    with pytest.raises(b2az.B2azError, match="multiple of the number of seat permutations"):
        b2az.Engine(p, lib=lib)
    p = b2az.default_params(lib, games_to_play=2, concurrent_games=2, mcts_visits=(8, 8), eval_type=RANDOM)
    b2az.fill_perms(p, [[0, 2], [1, 0]])
    with pytest.raises(b2az.B2azError, match="group index must be 0 or 1"):
        b2az.Engine(p, lib=lib)


# ------------------------------------------------------------------------------------ the wide-tree engine (tafl games)
import tafl_ref  # noqa: E402

needs_tafl_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")


def _wide(game, slots, per_slot, max_turns, group_visits, seed, kw, **extra):
    words = 2 * (1 + (max_turns + 2) * max(group_visits) * (1 + 8 * (64 if game == 0 else 200)))
    return b2az.TaflSelfplay(game, slots, max_turns, max(group_visits), games_per_slot=per_slot, seed=seed, words_per_tree=words,
                             hist_capacity=slots * per_slot * max_turns, seat_perms=PERMS,
                             perm_seat_visits=[[group_visits[g] for g in perm] for perm in PERMS], **kw, **extra)


def _finish(sp):
    canon, v, pi, slot = sp.drain_history()
    st, err = sp.slots()
    perm = sp.perm_stats()
    sp.close()
    assert (err == 0).all() and (st["error"] == 0).all()
    return canon, v, pi, slot, st, perm


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,slots,per_slot,max_turns,kw", [
    (0, 6, 2, 40, dict(epsilon=0.25, root_policy_temp=1.25, policy_target_pruning=True)),
    (0, 4, 1, 40, dict(gumbel_m=16, root_policy_temp=1.25)),
    (2, 4, 1, 24, dict()),
])
def test_wide_engine_seat_perms_vs_reference(game, slots, per_slot, max_turns, kw):
    """Brandubh / Tawlbwrdd self-play with two model groups under two seat permutations, every group EvalType::RANDOM with
    its own visit budget: slot g == the unmodified reference PlayManager with seat_perms = [PERMS[g % 2]]."""
    group_visits, seed = (40, 24), 5150
    sp = _wide(game, slots, per_slot, max_turns, group_visits, seed, kw)
    active, rounds = slots, 0
    while active:
        active = sp.play(8)
        rounds += 1
        assert rounds < 10000
    canon, v, pi, slot, st, perm = _finish(sp)
    want = np.zeros((len(PERMS), 3), np.float32)
    games = [0] * len(PERMS)
    for g in range(slots):
        ref = tafl_ref.selfplay(game, seed + g, max_turns, per_slot, max(group_visits), seat_perm=PERMS[g % 2],
                                group_visits=group_visits, **kw)
        rows = slot == g
        assert rows.sum() == len(ref["v"]), f"slot {g}: {rows.sum()} samples vs {len(ref['v'])}"
        assert np.array_equal(canon[rows].view(np.uint32), ref["canonical"].view(np.uint32)), f"slot {g}: canonical"
        assert np.array_equal(v[rows].view(np.uint32), ref["v"].view(np.uint32)), f"slot {g}: outcomes"
        assert np.array_equal(pi[rows].view(np.uint32), ref["pi"].view(np.uint32)), f"slot {g}: policy targets"
        assert np.array_equal(st[g]["scores"], ref["scores"])
        want[g % 2] += ref["scores"]
        games[g % 2] += ref["games_completed"]
    assert len(perm) == 2
    for p in range(2):
        assert np.array_equal(perm[p]["scores"], want[p]) and perm[p]["games_completed"] == games[p] == slots // 2 * per_slot


def _dumb_eval(game, canon, max_turns):
    S = canon.shape[-1]
    board = np.zeros((3, S, S), np.int8)
    board[:] = canon[:3] > 0
    player = 0 if canon[3, 0, 0] > 0 else 1
    valid = tafl_ref.position(game, board, player, 0, max_turns, 0)["valid"].astype(np.float32)
    total = np.float32(int(valid.sum()) % 256)  # dumb_eval (game_state.h:160-173): Vector<uint8_t>::sum() wraps mod 256
    return np.full(3, np.float32(1.0 / 3.0)), (valid / total if total > 0 else valid)


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("group_random,cache", [((0, 1), 0), ((1, 0), 0), ((0, 0), 0), ((0, 1), 4096)])
def test_wide_engine_mixed_evaluators_route_by_group(group_random, cache):
    """EvalType::NN for one model group next to EvalType::RANDOM for the other (play_past against iteration 0), and two NN
    groups: rows reach the evaluator of the group that searches for the seat under the slot's permutation, a RANDOM
    group's searches never do. With every NN evaluator answering dumb_eval's numbers the run equals the fused all-RANDOM
    run of the same permutations (itself pinned to the reference above), with and without the position cache."""
    game, slots, max_turns, group_visits, seed = 0, 4, 30, (24, 16), 99
    kw = dict(epsilon=0.25, root_policy_temp=1.25, policy_target_pruning=True)
    sp = _wide(game, slots, 1, max_turns, group_visits, seed, kw)
    active = slots
    while active:
        active = sp.play(8)
    fused = _finish(sp)
    sp = _wide(game, slots, 1, max_turns, group_visits, seed, kw, group_random=group_random, cache_entries=cache)
    rows_by_group = [0, 0]
    for rounds in range(200000):
        ids, canon = sp.leaf_batch_host()
        if len(ids) == 0 and sp.stats().active_games == 0:
            break
        groups = sp.leaf_groups_host(len(ids))
        vs, pis = np.zeros((len(ids), 3), np.float32), np.zeros((len(ids), sp.A), np.float32)
        for i, g in enumerate(ids):
            # the searching seat is the slot's side to move; its group under permutation g % 2 must be the row's group
            assert not group_random[groups[i]], "a RANDOM group's leaf reached the evaluator"
            rows_by_group[groups[i]] += 1
            vs[i], pis[i] = _dumb_eval(game, canon[i], max_turns)
        sp.submit_eval_host(ids, vs, pis)
    else:
        raise AssertionError("did not finish")
    canon, v, pi, slot, st, perm = _finish(sp)
    assert np.array_equal(slot, fused[3]) and np.array_equal(v, fused[1])
    assert np.array_equal(canon.view(np.uint32), fused[0].view(np.uint32))
    assert np.array_equal(pi.view(np.uint32), fused[2].view(np.uint32))
    for p in range(2):
        assert np.array_equal(perm[p]["scores"], fused[5][p]["scores"]) and perm[p]["games_completed"] == fused[5][p]["games_completed"]
    for g in (0, 1):
        assert (rows_by_group[g] > 0) == (not group_random[g])


# per-seat search settings (PlayParams::seat_epsilon / seat_mcts_root_temp / seat_root_fpu_zero / seat_gumbel_*: every seat's
# MCTS object is built with its own, play_manager.cc:92-164, 602-617) and the per-seat resign rule (:335-366)
SEAT_CASES = {
    "gumbel_seat_vs_puct_seat": (0, 4, 2, 40, 40, dict(policy_target_pruning=True), dict(
        seat_gumbel_enabled=(1, 0), seat_gumbel_m=(8, 16), seat_epsilon=(0.0, 0.25), seat_root_temp=(1.0, 1.25),
        seat_root_fpu_zero=(0, 1))),
    "two_gumbel_seats_different_scales": (0, 4, 1, 40, 48, dict(), dict(
        seat_gumbel_enabled=(1, 1), seat_gumbel_m=(16, 4), seat_gumbel_c_visit=(50.0, 20.0), seat_gumbel_c_scale=(1.0, 0.5),
        seat_epsilon=(0.25, 0.0))),
    "seat_resign_rule": (0, 6, 2, 60, 32, dict(epsilon=0.25), dict(
        seat_resign_threshold=(0.0, -2.0), seat_resign_consecutive=(8, 1))),  # W - L = 2 q_best - 1 <= 0: no winning line found, eight own moves in a row
    "tawlbwrdd_seat_resign_both": (2, 3, 1, 24, 32, dict(), dict(
        seat_resign_threshold=(0.0, 0.0), seat_resign_consecutive=(5, 8), seat_epsilon=(0.25, 0.0))),
}


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("name", sorted(SEAT_CASES))
def test_wide_engine_per_seat_settings_vs_reference(name):
    game, slots, per_slot, max_turns, visits, kw, seat = SEAT_CASES[name]
    seed = 31337 + 7 * sorted(SEAT_CASES).index(name)
    words = 2 * (1 + (max_turns + 2) * visits * (1 + 8 * (64 if game == 0 else 200)))
    sp = b2az.TaflSelfplay(game, slots, max_turns, visits, games_per_slot=per_slot, seed=seed, words_per_tree=words,
                           hist_capacity=slots * per_slot * max_turns, seat_search={k: [v] for k, v in seat.items()}, **kw)
    active, rounds = slots, 0
    while active:
        active = sp.play(8)
        rounds += 1
        assert rounds < 10000
    canon, v, pi, slot, st, _ = _finish(sp)
    resigned = 0
    for g in range(slots):
        ref = tafl_ref.selfplay(game, seed + g, max_turns, per_slot, visits, seat_cfg=seat, **kw)
        rows = slot == g
        assert rows.sum() == len(ref["v"]), f"{name} slot {g}: {rows.sum()} samples vs {len(ref['v'])}"
        assert np.array_equal(canon[rows].view(np.uint32), ref["canonical"].view(np.uint32)), f"{name} slot {g}: canonical"
        assert np.array_equal(v[rows].view(np.uint32), ref["v"].view(np.uint32)), f"{name} slot {g}: outcomes"
        assert np.array_equal(pi[rows].view(np.uint32), ref["pi"].view(np.uint32)), f"{name} slot {g}: policy targets"
        assert np.array_equal(st[g]["scores"], ref["scores"])
        assert np.array_equal(st[g]["resign_scores"], ref["resign_scores"])
        assert np.float32(np.float32(st[g]["game_length"]) / np.float32(st[g]["games_completed"])) == ref["avg_game_length"]
        resigned += ref["resign_scores"].sum()
    if "resign" in name:
        assert resigned > 0, "the case never exercised the per-seat resign rule"
