"""CPU suite: the product's engine logic (csrc/az_engine_logic.h) compiled for the host (W = 1 lane,
tests/cpp/libb2az_hostemu.so, -DB2AZ_HOST_EMU) against the oracle. This checks the host logic of the
C ABI and the SoA/paged tree bookkeeping without a GPU; the CUDA build of the same code is checked by
tests/test_gpu_parity.py. The host-emulation library is test scaffolding and is never shipped."""
import numpy as np
import pytest

import b2az
import parity_harness as ph
from conftest import needs_ref

EMU = ph.HOSTEMU_LIB


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("rng_mode", [b2az.RNG_GLOBAL, b2az.RNG_PER_GAME])
def test_lockstep_vs_port(level, rng_mode):
    r = ph.run_lockstep_parity(EMU, G=5, games_to_play=9, visits=36, level=level, seed=4242, oracle="port",
                               rng_mode=rng_mode)
    assert r["games"] == 9 and r["moves_compared"] > 50


@needs_ref
@pytest.mark.parametrize("level", [0, 1])
def test_lockstep_vs_reference(level):
    r = ph.run_lockstep_parity(EMU, G=4, games_to_play=6, visits=50, level=level, seed=12345, oracle="ref")
    assert r["games"] == 6


@needs_ref
def test_config0_100_sims_vs_reference():
    # BASELINE.json configs[0]: Connect4 via PlayManager, 100 sims/move, deterministic seed
    r = ph.run_lockstep_parity(EMU, G=2, games_to_play=3, visits=100, level=1, seed=12345, oracle="ref")
    assert r["games"] == 3


@pytest.mark.parametrize("rng_mode", [b2az.RNG_GLOBAL, b2az.RNG_PER_GAME])
def test_random_eval_vs_port(rng_mode):
    r = ph.run_random_parity(EMU, G=16, games_to_play=40, visits=48, seed=11, oracle="port", rng_mode=rng_mode,
                             level=1)
    assert r["games"] == 40


@needs_ref
def test_random_eval_vs_reference():
    r = ph.run_random_parity(EMU, G=8, games_to_play=20, visits=64, seed=12345, oracle="ref", level=0)
    assert r["games"] == 20


@needs_ref
@pytest.mark.parametrize("level", [0, 1])
def test_per_game_slots_equal_reference_single_game_runs(level):
    # B2AZ_RNG_PER_GAME: slot g == the unmodified reference PlayManager(concurrent_games=1) seeded seed + g
    r = ph.run_slots_vs_reference(EMU, G=6, quota=3, visits=48, level=level, seed=991, chunk=37)
    assert r["games"] == 18


def test_no_tree_reuse_vs_port():
    ph.run_lockstep_parity(EMU, G=3, games_to_play=5, visits=24, level=1, seed=3, oracle="port", tree_reuse=False)
    ph.run_random_parity(EMU, G=4, games_to_play=8, visits=24, seed=3, oracle="port", level=2, tree_reuse=False)


def test_single_game_many_restarts():
    # one slot, many sequential games: exercises page recycling (Cheney re-root + free) over and over
    r = ph.run_random_parity(EMU, G=1, games_to_play=30, visits=32, seed=8, oracle="port", level=1)
    assert r["games"] == 30


def test_pool_pages_are_recycled():
    lib = b2az.load(EMU)
    p = b2az.default_params(lib, games_to_play=64, concurrent_games=4, mcts_visits=(40, 40), eval_type=b2az.EVAL_RANDOM,
                            pool_nodes=4 * 2 * 2048, seed=1)
    e = b2az.Engine(p, lib=lib)
    for _ in range(4000):
        e.step(16)
        st = e.stats()
        if st.active_games == 0:
            break
    assert st.games_completed == 64 and st.device_error == 0
    assert st.pool_pages_free == st.pool_pages_total, "every page must be back on the free stacks at the end"
    e.close()


@pytest.mark.parametrize("compact_pages", [1, 2, 50000])
def test_compaction_threshold_does_not_change_results(compact_pages):
    """Re-rooting re-points the tree header; the Cheney copy only runs when the arena is over budget. Whatever
    the budget (compact at every move ... never), the run must equal the oracle bit for bit."""
    r = ph.run_lockstep_parity(EMU, G=5, games_to_play=9, visits=48, level=1, seed=4242, oracle="port",
                               rng_mode=b2az.RNG_GLOBAL, compact_pages=compact_pages)
    assert r["games"] == 9
    if compact_pages == 1:
        assert r["compactions"] > 50
    if compact_pages == 50000:
        assert r["compactions"] == 0


@pytest.mark.parametrize("chunk", [1, 7, 400])
def test_flattened_loop_matches_port(chunk, monkeypatch):
    """run_flat() — the fused kernel's one-tree-level-per-iteration loop — run game by game on the CPU
    (B2AZ_EMU_FLAT=1) must produce what the step-by-step loop produces. No slot retires, so the comparison is
    order independent."""
    monkeypatch.setenv("B2AZ_EMU_FLAT", "1")
    r = ph.run_random_parity(EMU, G=24, games_to_play=10 ** 6, visits=60, seed=77, oracle="port",
                             rng_mode=b2az.RNG_PER_GAME, level=1, chunk=chunk, steps=1200, ordered=False)
    assert r["games"] > 5 and r["samples"] > 100


CAP_RESIGN = dict(playout_cap_randomization=1, playout_cap_depth=6, playout_cap_percent=0.6, resign_percent=0.35,
                  resign_playthrough_percent=0.3)


@pytest.mark.parametrize("flat", ["0", "1"])
def test_playout_cap_and_resign_vs_port(flat, monkeypatch):
    """Playout-cap randomisation (capped searches: cap visits, no noise, not recorded, fast_* metrics) and
    resign_percent / playthrough (play_manager.cc:305-337, 440-444, 486-488, 523-524, 559-560) against the port;
    the coin flips come from each game's own stream on both sides."""
    monkeypatch.setenv("B2AZ_EMU_FLAT", flat)
    r = ph.run_random_parity(EMU, G=20, games_to_play=10 ** 6, visits=40, seed=8, oracle="port", rng_mode=b2az.RNG_PER_GAME,
                             level=1, chunk=37, steps=1500, ordered=False, extra=CAP_RESIGN)
    assert r["games"] > 20 and sum(r["resign_scores"]) > 0 and r["fast_avg_leaf_depth"] > 0
    r = ph.run_lockstep_parity(EMU, G=6, games_to_play=10 ** 6, visits=24, level=1, seed=8, oracle="port",
                               rng_mode=b2az.RNG_PER_GAME, max_generations=1500, extra=CAP_RESIGN)
    assert r["games"] > 6


@needs_ref
@pytest.mark.parametrize("level", [3, 4])
def test_gumbel_vs_reference_live(level):
    """Gumbel root search (init / halving phases / next root child / interior select / improved policy / final action,
    mcts.cc:175-401) against the UNMODIFIED reference, lock-step, bit for bit."""
    r = ph.run_lockstep_parity(EMU, G=5, games_to_play=9, visits=40, level=level, seed=99, oracle="ref")
    assert r["games"] == 9


@pytest.mark.parametrize("flat", ["0", "1"])
def test_gumbel_with_playout_cap_vs_port(flat, monkeypatch):
    monkeypatch.setenv("B2AZ_EMU_FLAT", flat)
    extra = dict(fast_search_uses_gumbel=1, playout_cap_randomization=1, playout_cap_depth=12, playout_cap_percent=0.5)
    for level in (3, 4):
        r = ph.run_random_parity(EMU, G=16, games_to_play=10 ** 6, visits=40, seed=5, oracle="port", rng_mode=b2az.RNG_PER_GAME,
                                 level=level, steps=1200, chunk=29, ordered=False, extra=extra)
        assert r["games"] > 16
    r = ph.run_lockstep_parity(EMU, G=6, games_to_play=10 ** 6, visits=40, level=4, seed=99, oracle="port",
                               rng_mode=b2az.RNG_PER_GAME, max_generations=1200, extra=dict(extra, fast_search_uses_gumbel=0))
    assert r["games"] > 6


def test_cap_and_resign_rejected_in_reference_parity_mode():
    lib = b2az.load(EMU)
    for kw in (dict(playout_cap_randomization=1), dict(resign_percent=0.1)):
        with pytest.raises(b2az.B2azError, match="unseedable"):
            b2az.Engine(b2az.default_params(lib, rng_mode=b2az.RNG_GLOBAL, **kw), lib=lib)


def test_pool_exhaustion_is_reported():
    lib = b2az.load(EMU)
    p = b2az.default_params(lib, games_to_play=8, concurrent_games=8, mcts_visits=(400, 400), eval_type=b2az.EVAL_NN,
                            pool_nodes=64 * 256, seed=1)
    e = b2az.Engine(p, lib=lib)
    with pytest.raises(b2az.B2azError) as ei:
        for _ in range(3000):
            e.step(1)
            ids, canon = e.leaf_batch_host()
            v, pi = ph.fake_net(canon)
            e.submit_eval_host(ids, v, pi)
    assert ei.value.code == -3
    e.close()


def test_error_conventions():
    lib = b2az.load(EMU)
    with pytest.raises(b2az.B2azError, match="MCTS visits"):  # play_manager.cc:21
        b2az.Engine(b2az.default_params(lib, mcts_visits=(0, 10)), lib=lib)
    with pytest.raises(b2az.B2azError):
        b2az.Engine(b2az.default_params(lib, concurrent_games=0), lib=lib)
    with pytest.raises(b2az.B2azError, match="gumbel_m"):
        b2az.Engine(b2az.default_params(lib, gumbel_enabled=1, gumbel_m=0), lib=lib)
    e = b2az.Engine(b2az.default_params(lib, concurrent_games=2, games_to_play=2, mcts_visits=(8, 8)), lib=lib)
    e.step(1)
    with pytest.raises(b2az.B2azError, match="still waiting"):  # stepping with unanswered leaves
        e.step(1)
    ids, canon = e.leaf_batch_host()
    assert list(ids) == [0, 1] and canon.shape == (2, 4, 6, 7)
    # the first leaf of a fresh game is the empty board seen by player 0 (connect4_gs.cc:131-149)
    assert canon[0, :2].sum() == 0 and canon[0, 2].min() == 1 and canon[0, 3].max() == 0
    e.close()


def test_legacy_partial_batches():
    """build_batch may hand out the leaves in several pieces and update_inferences may answer them in any
    grouping (py_wrapper.cc:449-504, play_manager.cc:619-642)."""
    lib = b2az.load(EMU)
    kw = ph.level_params(0)
    mk = lambda: b2az.Engine(b2az.default_params(lib, concurrent_games=7, games_to_play=7, mcts_visits=(12, 12),
                                                 history_enabled=1, seed=5, **kw), lib=lib)
    a, b = mk(), mk()
    for _ in range(2000):
        a.step(1)
        b.step(1)
        ids, canon = a.leaf_batch_host()
        if len(ids) == 0:
            break
        v, pi = ph.fake_net(canon)
        a.submit_eval_host(ids, v, pi)
        got = []
        while True:
            i2, c2 = b.leaf_batch_host(max_rows=3)
            if len(i2) == 0:
                break
            got.append((i2.copy(), c2.copy()))
        assert np.array_equal(np.concatenate([g[0] for g in got]), ids)
        for i2, c2 in reversed(got):
            v2, p2 = ph.fake_net(c2)
            b.submit_eval_host(i2[::-1], v2[::-1], p2[::-1])
    ha, hb = a.drain_history(1000), b.drain_history(1000)
    ph.compare_history(ha, hb, ordered=True)
    assert a.stats().games_completed == 7
    a.close()
    b.close()
