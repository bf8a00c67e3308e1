"""The batched single-tree MCTS over the tafl games (b2az_forest_*: MCTS::find_leaf / process_result / update_root /
counts for many trees, one warp per tree) against the UNMODIFIED reference's MCTS class driven over the unmodified
tafl games (oracle/_ref/libazref_tafl.so: azref_tafl_search). Tree i of the forest == a reference MCTS run after
MCTS::seed_thread_rng(seed + i). Visit counts and Q values after every move's search are compared bit-exact, with
the reference's dumb_eval evaluator (fused device launches) and with a deterministic pseudo-network supplied by
the host on both sides (leaf canonical planes out, v / pi in). Golden fixtures (tools/make_golden_forest.py) make
the dumb_eval cases runnable without the reference."""
import os
import zlib

import numpy as np
import pytest

import b2az
import parity_harness as ph
import tafl_ref

NAMES = {0: "brandubh", 1: "opentafl", 2: "tawlbwrdd"}
MAX_TURNS = {0: 150, 1: 400, 2: 400}
for _g in (10, 11, 12, 13, 20, 21, 22, 23):  # Star Gambit (tests/test_stargambit_search.py): max_turns is unused by the search
    NAMES[_g] = f"stargambit{_g}"
    MAX_TURNS[_g] = 768
needs_tafl_ref = pytest.mark.skipif(not tafl_ref.available(), reason="oracle/_ref/libazref_tafl.so not built")
GOLDEN = os.path.join(ph.ROOT, "tests", "golden", "forest_random_eval.npz")
# (game, trees, moves, sims per move, seed, cpuct, fpu_reduction, root_fpu_zero)
GOLDEN_CASES = {"brandubh": (0, 6, 12, 60, 100, 1.25, 0.25, False), "opentafl": (1, 3, 6, 50, 200, 2.0, 0.2, True),
                "tawlbwrdd": (2, 3, 6, 50, 300, 1.25, 0.0, False)}


def pseudo_net(canon):
    """Deterministic evaluator: a function of the canonical planes' bytes only (so both sides see identical v, pi)."""
    A = 2 * canon.shape[1] ** 3
    rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(canon, np.float32).tobytes()))
    v = rng.random(3).astype(np.float32) + np.float32(0.05)
    v /= v.sum()
    pi = rng.random(A).astype(np.float32) ** 4 + np.float32(1e-3)
    pi /= pi.sum()
    return v.astype(np.float32), pi.astype(np.float32)


def run_forest(game, trees, n_moves, sims, seed, cpuct, fpu, rfz, evaluator, moves_ref=None, slab_moves=None, gumbel=None,
               noise=None, width=0):
    """Drives the forest like the reference run; plays `moves_ref[i][m]` when given (else the most visited move)."""
    # each half of the slab must hold the kept subtree + one move's search (1 + 8k words per expanded node);
    # slab_moves = how many moves' worth of nodes one half can hold (re-rooting compacts into the other half)
    words = 2 * (1 + (slab_moves or n_moves + 1) * sims * (1 + 8 * (64 if game == 0 else 200)))
    gkw = dict(gumbel_m=gumbel[0], gumbel_c_visit=gumbel[1], gumbel_c_scale=gumbel[2], gumbel_full=len(gumbel) > 3 and gumbel[3]) if gumbel else {}
    if noise:  # (epsilon, root_policy_temp, shaped_dirichlet)
        gkw.update(epsilon=noise[0], root_policy_temp=noise[1], shaped_dirichlet=noise[2])
    rn = bool(noise and noise[0] > 0)
    f = b2az.Forest(game, trees, MAX_TURNS[game], cpuct=cpuct, fpu_reduction=fpu, root_fpu_zero=rfz, seed=seed,
                    words_per_tree=words, max_in_flight=width, **gkw)
    out = []
    alive = np.ones(trees, bool)
    for m in range(n_moves):
        if gumbel:
            f.set_gumbel_num_sims(sims)
        if width and evaluator is None:
            f.simulate_batched(sims // width, width)
        elif width:  # WU-UCT with the host evaluator: `width` pending leaves per tree and round
            for _ in range(sims // width):
                for _i in range(width):
                    f.find_leaf_batched()
                canon = f.leaf_canon()
                for li in range(width):
                    ev = [evaluator(canon[li, i]) for i in range(trees)]
                    f.process_result_batched(li, np.stack([e[0] for e in ev]), np.stack([e[1] for e in ev]))
                f.reset_batch()
        elif evaluator is None:
            f.simulate(sims, root_noise=rn)
        else:
            for _ in range(sims):
                f.find_leaf()
                canon = f.leaf_canon()
                ev = [evaluator(canon[i]) for i in range(trees)]
                f.process_result(np.stack([e[0] for e in ev]), np.stack([e[1] for e in ev]), root_noise=rn)
        counts, q, info = f.counts()
        assert (info["error"] == 0).all(), info["error"]
        out.append((counts, q, info) + (f.gumbel_result() if gumbel else ()))
        mv = np.full(trees, 0xFFFFFFFF, np.uint32)
        for i in range(trees):
            if moves_ref is not None:
                if m < len(moves_ref[i]):
                    mv[i] = moves_ref[i][m]
                else:
                    alive[i] = False
            else:
                mv[i] = int(np.argmax(counts[i]))
        f.update_root(mv)
        if noise:
            f.root_noise(add_noise=rn)  # the reused root: temperature again + fresh noise (play_manager.cc:546-553)
    f.close()
    return out


def _compare(game, trees, n_moves, sims, seed, cpuct, fpu, rfz, evaluator, slab_moves=None):
    refs = [tafl_ref.search(game, seed + i, n_moves, sims, MAX_TURNS[game], cpuct, fpu, rfz, evaluator) for i in range(trees)]
    got = run_forest(game, trees, n_moves, sims, seed, cpuct, fpu, rfz, evaluator, moves_ref=[r[2] for r in refs],
                     slab_moves=slab_moves)
    compared = 0
    for i, (rc, rq, rm, rd) in enumerate(refs):
        for m in range(len(rm)):
            counts, q, info = got[m][:3]
            assert np.array_equal(counts[i], rc[m]), f"{NAMES[game]} tree {i} move {m}: visit counts differ"
            assert np.array_equal(q[i].view(np.uint32), rq[m].view(np.uint32)), f"tree {i} move {m}: Q values differ"
            assert info["total_leaf_depth"][i] == rd[m]
            compared += 1
    assert compared >= trees


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,trees,n_moves,sims", [(0, 8, 20, 80), (1, 4, 8, 60), (2, 4, 8, 60)])
def test_forest_random_eval_vs_reference(game, trees, n_moves, sims):
    _compare(game, trees, n_moves, sims, 4242, 1.25, 0.25, False, None)


@pytest.mark.gpu
@needs_tafl_ref
def test_forest_long_game_in_a_small_slab():
    """60 moves in a slab that holds two moves' worth of nodes per half: only works because re-rooting compacts."""
    _compare(0, 6, 60, 60, 999, 1.25, 0.25, False, None, slab_moves=2)


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,trees,n_moves,sims,rfz", [(0, 4, 8, 40, True), (2, 2, 4, 30, False)])
def test_forest_host_evaluator_vs_reference(game, trees, n_moves, sims, rfz):
    _compare(game, trees, n_moves, sims, 77, 1.5, 0.3, rfz, pseudo_net)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_forest_reproduces_golden(name):
    game, trees, n_moves, sims, seed, cpuct, fpu, rfz = GOLDEN_CASES[name]
    g = np.load(GOLDEN)
    moves, lens = g[f"{name}_moves"], g[f"{name}_lens"]
    got = run_forest(game, trees, n_moves, sims, seed, cpuct, fpu, rfz, None,
                     moves_ref=[moves[i, :lens[i]] for i in range(trees)])
    for i in range(trees):
        for m in range(int(lens[i])):
            counts, q, _ = got[m][:3]
            assert zlib.crc32(counts[i].tobytes()) == g[f"{name}_counts_crc"][i, m], f"{name} tree {i} move {m}: counts"
            assert zlib.crc32(q[i].tobytes()) == g[f"{name}_q_crc"][i, m], f"{name} tree {i} move {m}: Q"


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,trees,n_moves,sims,m", [(0, 8, 16, 120, 16), (0, 4, 10, 30, 4), (1, 3, 5, 120, 16), (2, 3, 5, 64, 8)])
def test_forest_gumbel_root_search_vs_reference(game, trees, n_moves, sims, m):
    """configs/brandubh.yaml-style search: Gumbel top-m + sequential halving at the root, PUCT below; the move
    played is gumbel_final_action, the training target gumbel_improved_policy."""
    gum = (m, 50.0, 1.0)
    refs = [tafl_ref.search(game, 31 + i, n_moves, sims, MAX_TURNS[game], 1.25, 0.25, False, None, *gum) for i in range(trees)]
    got = run_forest(game, trees, n_moves, sims, 31, 1.25, 0.25, False, None, moves_ref=[r[2] for r in refs], gumbel=gum)
    compared = 0
    for i, (rc, rq, rm, rd, rp) in enumerate(refs):
        for mv in range(len(rm)):
            counts, q, info, action, policy = got[mv]
            assert np.array_equal(counts[i], rc[mv]), f"{NAMES[game]} tree {i} move {mv}: visit counts differ"
            assert np.array_equal(q[i].view(np.uint32), rq[mv].view(np.uint32)), f"tree {i} move {mv}: Q values differ"
            assert action[i] == rm[mv], f"tree {i} move {mv}: gumbel_final_action {action[i]} != {rm[mv]}"
            assert np.array_equal(policy[i].view(np.uint32), rp[mv].view(np.uint32)), f"tree {i} move {mv}: improved policy"
            compared += 1
    assert compared >= trees * 3


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,trees,n_moves,sims,noise,evaluator", [
    (0, 6, 12, 60, (0.25, 1.25, False), None), (0, 4, 8, 50, (0.25, 1.25, True), pseudo_net),
    (0, 4, 8, 40, (0.0, 1.4, False), pseudo_net), (1, 2, 4, 40, (0.25, 1.25, True), None), (2, 2, 4, 40, (0.3, 1.0, False), None)])
def test_forest_root_noise_and_temperature_vs_reference(game, trees, n_moves, sims, noise, evaluator):
    """Self-play settings of the non-Gumbel configs: root policy temperature, (shaped) Dirichlet noise at the first
    root evaluation and again on every reused root (gamma / normal / pow / log restated bit-exactly on the device)."""
    refs = [tafl_ref.search(game, 555 + i, n_moves, sims, MAX_TURNS[game], 1.25, 0.25, True, evaluator, epsilon=noise[0],
                            root_policy_temp=noise[1], shaped_dirichlet=noise[2]) for i in range(trees)]
    got = run_forest(game, trees, n_moves, sims, 555, 1.25, 0.25, True, evaluator, moves_ref=[r[2] for r in refs], noise=noise)
    for i, (rc, rq, rm, rd) in enumerate(refs):
        for m in range(len(rm)):
            counts, q, info = got[m][:3]
            assert np.array_equal(counts[i], rc[m]), f"{NAMES[game]} tree {i} move {m}: visit counts differ"
            assert np.array_equal(q[i].view(np.uint32), rq[m].view(np.uint32)), f"tree {i} move {m}: Q values differ"


@pytest.mark.gpu
@pytest.mark.parametrize("game", [0, 1])
def test_parallel_shuffle_draws_equal_sequential(game):
    """std::shuffle's draws come from a pcg32 jump-ahead table on all lanes (and fall back to the sequential path when
    Lemire's rejection could fire): forcing the sequential path everywhere must give the same trees."""
    outs = []
    for serial in (False, True):
        f = b2az.Forest(game, 64, MAX_TURNS[game], seed=9, words_per_tree=2 * (1 + 4 * 100 * (1 + 7 * 200)), serial_shuffle=serial)
        for _ in range(3):
            f.simulate(100)
            f.advance()
        f.simulate(100)
        outs.append(f.counts())
        f.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
    assert (outs[0][2]["error"] == 0).all() and outs[0][0].sum() > 0


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,trees,n_moves,sims,width,evaluator", [
    (0, 6, 10, 64, 8, None), (0, 4, 6, 48, 4, pseudo_net), (1, 2, 4, 64, 16, None), (2, 2, 4, 60, 6, None)])
def test_forest_wu_uct_batched_vs_reference(game, trees, n_moves, sims, width, evaluator):
    """find_leaf_batched / process_result_batched / reset_batch (virtual loss through n_in_flight, mcts.cc:752-851):
    rounds of `width` pending leaves per tree, answered in order."""
    refs = [tafl_ref.search(game, 808 + i, n_moves, sims, MAX_TURNS[game], 1.25, 0.25, False, evaluator, batch_width=width)
            for i in range(trees)]
    got = run_forest(game, trees, n_moves, sims, 808, 1.25, 0.25, False, evaluator, moves_ref=[r[2] for r in refs], width=width)
    for i, (rc, rq, rm, rd) in enumerate(refs):
        for m in range(len(rm)):
            counts, q, info = got[m][:3]
            assert np.array_equal(counts[i], rc[m]), f"{NAMES[game]} tree {i} move {m}: visit counts differ"
            assert np.array_equal(q[i].view(np.uint32), rq[m].view(np.uint32)), f"tree {i} move {m}: Q values differ"
            assert info["total_leaf_depth"][i] == rd[m]


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,temp,sims", [(0, 1.0, 60), (0, 0.5, 60), (0, 0.0, 40), (2, 0.8, 40), (0, 1.0, 1)])
def test_forest_probs_and_pick_move_vs_reference(game, temp, sims):
    """PlayManager's PUCT acting rule: probs(temp) over the dense move vector (sums in move order, pow through the
    restated powf) and pick_move (one uniform draw). sims = 1 exercises the raw-policy branch (no child visited)."""
    trees, n_moves = 5, 8
    ptemp = {1.0: 1.0, 0.5: 0.7, 0.0: 0.0, 0.8: 1.0}[temp]
    refs = [tafl_ref.search(game, 1300 + i, n_moves, sims, MAX_TURNS[game], 1.25, 0.25, False, pseudo_net, act_temp=temp,
                            pruned_temp=ptemp) for i in range(trees)]
    f = b2az.Forest(game, trees, MAX_TURNS[game], cpuct=1.25, fpu_reduction=0.25, seed=1300,
                    words_per_tree=2 * (1 + (n_moves + 1) * sims * (1 + 8 * (64 if game == 0 else 200))))
    for m in range(n_moves):
        for _ in range(sims):
            f.find_leaf()
            canon = f.leaf_canon()
            ev = [pseudo_net(canon[i]) for i in range(trees)]
            f.process_result(np.stack([e[0] for e in ev]), np.stack([e[1] for e in ev]))
        pruned = f.probs(ptemp, pruned=True)  # the policy target (no draw), then the acting rule (one draw)
        probs, picked = f.probs(temp, pick=True)
        mv = np.full(trees, 0xFFFFFFFF, np.uint32)
        for i, r in enumerate(refs):
            if m < len(r[2]):
                assert np.array_equal(pruned[i].view(np.uint32), r[5][m].view(np.uint32)), f"tree {i} move {m}: probs_pruned({ptemp})"
                assert np.array_equal(probs[i].view(np.uint32), r[4][m].view(np.uint32)), f"tree {i} move {m}: probs({temp})"
                assert picked[i] == r[2][m], f"tree {i} move {m}: pick_move {picked[i]} != {r[2][m]}"
                mv[i] = r[2][m]
        f.update_root(mv)
    f.close()


def test_forest_refuses_without_cuda_or_unsupported_params():
    lib = b2az.load(ph.HOSTEMU_LIB)
    with pytest.raises(b2az.B2azError) as ei:
        b2az.Forest(0, 4, 150, gumbel_m=100, lib=lib)
    assert "gumbel_m" in str(ei.value)
    with pytest.raises(b2az.B2azError) as ei:
        b2az.Forest(0, 4, 150, lib=lib)
    assert ei.value.code == -2


@pytest.mark.gpu
@needs_tafl_ref
@pytest.mark.parametrize("game,trees,n_moves,sims,m,evaluator", [(0, 4, 8, 64, 8, None), (0, 3, 6, 48, 16, pseudo_net), (23, 3, 8, 48, 8, None)])
def test_forest_gumbel_full_vs_reference(game, trees, n_moves, sims, m, evaluator):
    """gumbel_full (mcts.cc:285-334, 479-481): pi'-matching at the interior nodes as well — argmax of pi'(a) - N(a) / (1 + sum N)
    with pi' = softmax(log prior + sigma * completedQ) instead of PUCT below the root."""
    if evaluator is not None and game >= 10:
        pytest.skip("pseudo_net is sized for the tafl games")
    gum = (m, 50.0, 1.0)
    tafl_ref.lib().azref_tafl_set_gumbel_full(1)
    try:
        refs = [tafl_ref.search(game, 61 + i, n_moves, sims, MAX_TURNS[game], 1.25, 0.25, False, evaluator, *gum) for i in range(trees)]
    finally:
        tafl_ref.lib().azref_tafl_set_gumbel_full(0)
    got = run_forest(game, trees, n_moves, sims, 61, 1.25, 0.25, False, evaluator, moves_ref=[r[2] for r in refs], gumbel=gum + (True,))
    plain = [tafl_ref.search(game, 61 + i, n_moves, sims, MAX_TURNS[game], 1.25, 0.25, False, evaluator, *gum) for i in range(trees)]
    differs = False
    for i, (rc, rq, rm, rd, rp) in enumerate(refs):
        for mv in range(len(rm)):
            counts, q, info, action, policy = got[mv]
            assert np.array_equal(counts[i], rc[mv]), f"{NAMES[game]} tree {i} move {mv}: visit counts differ"
            assert np.array_equal(q[i].view(np.uint32), rq[mv].view(np.uint32)), f"tree {i} move {mv}: Q values differ"
            assert action[i] == rm[mv] and np.array_equal(policy[i].view(np.uint32), rp[mv].view(np.uint32))
            assert info["total_leaf_depth"][i] == rd[mv]
        differs |= len(plain[i][0]) != len(rc) or not np.array_equal(plain[i][0][: len(rc)], rc[: len(plain[i][0])])
    assert differs  # (the option really changes the search)
