"""Star Gambit under the wide-tree search (b2az_forest_*, one warp per tree: csrc/az_forest.h FGame<B2AZ_FOREST_SG>) and
under the device PlayManager (b2az_tafl_selfplay_*) against the UNMODIFIED reference: its MCTS class and its PlayManager
driven over the unmodified star_gambit_gs.cc (oracle/_ref/libazref_tafl.so). Tree i == a reference MCTS run after
MCTS::seed_thread_rng(seed + i); slot g == a reference PlayManager with concurrent_games = 1 after
seed_thread_rng(seed + g). relative_values (mcts.cc:522-524, play_manager.cc:451-455) is exercised by the host
pseudo-network (its v is not symmetric) and by the stored outcomes. Everything bit-exact, floats by bit pattern."""
import zlib

import numpy as np
import pytest

import b2az
import tafl_ref
from test_forest import _compare, run_forest

needs_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")


def pseudo_net_for(game):
    A = b2az.game_dims(game)[2]

    def net(canon):
        rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(canon, np.float32).tobytes()))
        v = rng.random(3).astype(np.float32) + np.float32(0.05)
        v /= v.sum()
        pi = rng.random(A).astype(np.float32) ** 4 + np.float32(1e-3)
        pi /= pi.sum()
        return v.astype(np.float32), pi.astype(np.float32)

    return net


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("game,trees,n_moves,sims", [(10, 6, 40, 60), (13, 4, 30, 60), (22, 4, 30, 50), (23, 4, 24, 80)])
def test_forest_random_eval_vs_reference(game, trees, n_moves, sims):
    _compare(game, trees, n_moves, sims, 5151 + game, 1.25, 0.25, False, None)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("game,trees,n_moves,sims,rfz", [(12, 3, 16, 40, True), (21, 3, 16, 40, False)])
def test_forest_host_evaluator_relative_values_vs_reference(game, trees, n_moves, sims, rfz):
    _compare(game, trees, n_moves, sims, 99, 1.5, 0.3, rfz, pseudo_net_for(game))


@pytest.mark.gpu
@needs_ref
def test_forest_gumbel_vs_reference():
    game, trees, n_moves, sims, seed = 23, 4, 20, 64, 777
    gum = (16, 50.0, 1.0)
    refs = [tafl_ref.search(game, seed + i, n_moves, sims, 768, 1.25, 0.25, False, None, gumbel_m=gum[0],
                            gumbel_c_visit=gum[1], gumbel_c_scale=gum[2]) for i in range(trees)]
    got = run_forest(game, trees, n_moves, sims, seed, 1.25, 0.25, False, None, moves_ref=[r[2] for r in refs], gumbel=gum)
    for i, (rc, rq, rm, rd, rp) in enumerate(refs):
        for m in range(len(rm)):
            counts, q, info, action, policy = got[m]
            assert np.array_equal(counts[i], rc[m]), (i, m)
            assert np.array_equal(q[i].view(np.uint32), rq[m].view(np.uint32)), (i, m)
            assert action[i] == rm[m], (i, m)
            assert np.array_equal(policy[i].view(np.uint32), rp[m].view(np.uint32)), (i, m)


SP_CASES = {
    # name -> (game, slots, games per slot, visits, kwargs)
    "skirmish_puct": (10, 4, 1, 32, dict(epsilon=0.25, root_policy_temp=1.25, shaped_dirichlet=True, policy_target_pruning=True,
                                         start_temp=1.0, final_temp=0.2, temp_decay_half_life=10.0)),
    "unified_clash_gumbel": (22, 3, 2, 32, dict(gumbel_m=16, root_policy_temp=1.25)),
    "unified_battle_gumbel": (23, 3, 1, 40, dict(gumbel_m=16)),
}
STAGE_ROWS = 1024


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", sorted(SP_CASES))
def test_selfplay_equals_the_reference_playmanager(name):
    game, slots, per_slot, visits, kw = SP_CASES[name]
    seed = 4000 + 13 * sorted(SP_CASES).index(name)
    sp = b2az.TaflSelfplay(game, slots, STAGE_ROWS, visits, games_per_slot=per_slot, seed=seed,
                           hist_capacity=slots * per_slot * STAGE_ROWS, **kw)
    active, rounds = slots, 0
    while active:
        active = sp.play(16)
        rounds += 1
        assert rounds < 5000
    canon, v, pi, slot = sp.drain_history()
    st, err = sp.slots()
    sp.close()
    assert (err == 0).all() and (st["error"] == 0).all()
    f32 = np.float32
    for g in range(slots):
        ref = tafl_ref.selfplay(game, seed + g, STAGE_ROWS, per_slot, visits, **kw)
        rows = slot == g
        assert rows.sum() == len(ref["v"]), f"{name} slot {g}: {rows.sum()} samples vs {len(ref['v'])}"
        assert np.array_equal(v[rows].view(np.uint32), ref["v"].view(np.uint32)), f"{name} slot {g}: outcomes"
        assert np.array_equal(canon[rows].view(np.uint32), ref["canonical"].view(np.uint32)), f"{name} slot {g}: canonical"
        assert np.array_equal(pi[rows].view(np.uint32), ref["pi"].view(np.uint32)), f"{name} slot {g}: policy targets"
        s = st[g]
        assert s["games_completed"] == ref["games_completed"] == per_slot and s["active"] == 0
        assert np.array_equal(s["scores"], ref["scores"])
        assert f32(f32(s["game_length"]) / f32(s["games_completed"])) == ref["avg_game_length"]
        assert f32(s["leaf_depth"] / float(s["total_full_move_count"])) == ref["avg_leaf_depth"]
        assert f32(s["valid_moves"] / float(s["total_move_count"])) == ref["avg_valid_moves"]
    assert len(np.unique(v, axis=0)) >= 2  # both frames of the outcome occur


@pytest.mark.gpu
def test_selfplay_variant_mix_with_one_weighted_variant_equals_the_pinned_run():
    """game 24 (StarGambitUnifiedGS with the variant mix) draws the variant of every new game from the slot's coin stream —
    apart from the search's generator, so a mix that can only draw Clash is the pinned-Clash run, sample for sample."""
    def run(game, **kw):
        sp = b2az.TaflSelfplay(game, 4, STAGE_ROWS, 24, games_per_slot=2, seed=31, hist_capacity=8 * STAGE_ROWS, gumbel_m=8, **kw)
        active = 4
        while active:
            active = sp.play(16)
        out = sp.drain_history()
        st, err = sp.slots()
        sp.close()
        assert (err == 0).all() and (st["error"] == 0).all()
        return out
    a = run(22)
    b = run(24, variant_probs=[0.0, 0.0, 1.0, 0.0])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
