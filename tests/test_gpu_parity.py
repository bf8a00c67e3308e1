"""GPU parity tests proper: the CUDA library (libb2az.so, sm_100a) through the C ABI vs the oracle.

  * serial kernel (B2AZ_RNG_GLOBAL) vs golden traces generated from the unmodified reference — bit-exact
    move lists, leaf batches, visit counts, Q, history targets;
  * serial kernel vs the reference itself when oracle/_ref/libazref.so travelled with the snapshot;
  * parallel kernels (B2AZ_RNG_PER_GAME, every lane width) vs the oracle port with per-game streams;
  * bitboard game kernels vs the port on random walks and the reference's known answers;
  * size-independent properties at the full BASELINE size (65,536 games).
Nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

import b2az
import parity_harness as ph
from conftest import has_cuda, needs_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


@pytest.mark.parametrize("name", sorted(ph.GOLDEN_CASES))
def test_serial_kernel_reproduces_golden(name):
    G, games, visits, level, seed, et = ph.GOLDEN_CASES[name]
    pm = ph.EnginePM(None, G=G, games_to_play=games, visits=visits, eval_type=et, rng_mode=b2az.RNG_GLOBAL, seed=seed,
                     **ph.level_params(level))
    got = ph.trace_run(pm, et)
    pm.close()
    ph.compare_trace(got, dict(np.load(ph.golden_path(name))), f"CUDA serial kernel vs golden {name}")


@needs_ref
@pytest.mark.parametrize("level", [0, 1])
def test_serial_kernel_vs_reference_live(level):
    r = ph.run_lockstep_parity(None, G=6, games_to_play=9, visits=64, level=level, seed=777, oracle="ref")
    assert r["games"] == 9


@pytest.mark.parametrize("compact_pages", [0, 1, 50000])
@pytest.mark.parametrize("level", [0, 1])
def test_parallel_kernel_lockstep_vs_port(compact_pages, level):
    # games_to_play is out of reach, so no slot retires: which slot takes the LAST games of a run depends on
    # completion order (atomics here, thread timing in the reference) and is not a parity target
    r = ph.run_lockstep_parity(None, G=48, games_to_play=10 ** 6, visits=40, level=level, seed=2024, oracle="port",
                               rng_mode=b2az.RNG_PER_GAME, compact_pages=compact_pages, peek_every=13,
                               max_generations=1600)
    assert r["games"] > 48 and r["moves_compared"] > 500
    assert (r["compactions"] > 0) == (compact_pages != 50000)


@pytest.mark.parametrize("chunk,compact_pages", [(128, 0), (1, 2), (400, 50000)])
def test_parallel_kernel_random_eval_vs_port(chunk, compact_pages):
    # fused launches (the game slot, tree header and RNG stay in registers across `chunk` generations)
    r = ph.run_random_parity(None, G=512, games_to_play=10 ** 6, visits=100, seed=31337, oracle="port",
                             rng_mode=b2az.RNG_PER_GAME, level=1, chunk=chunk, steps=4000 if chunk > 1 else 1500,
                             compact_pages=compact_pages,
                             pool_nodes=40_000_000 if compact_pages == 50000 else 0)  # never compacting needs room
    assert r["games"] > (512 if chunk > 1 else 100)


def test_parallel_kernel_400_sims_vs_port():
    # the headline search size (400 sims/move) on a batch the port finishes in seconds
    r = ph.run_random_parity(None, G=256, games_to_play=10 ** 6, visits=400, seed=5, oracle="port",
                             rng_mode=b2az.RNG_PER_GAME, level=0, chunk=400, steps=12000)
    assert r["games"] > 100


@needs_ref
@pytest.mark.parametrize("step_kernel", [b2az.STEP_QUEUE, b2az.STEP_FLAT, b2az.STEP_WAVES, b2az.STEP_SYNC])
def test_fused_kernel_slots_equal_reference_single_game_runs(step_kernel):
    """The TIMED kernel (fused launches of 400 generations, per-game RNG, 400 sims/move) against the UNMODIFIED
    reference: slot g == PlayManager(concurrent_games=1) after MCTS::seed_thread_rng(seed + g); samples, scores,
    game lengths and search metrics."""
    r = ph.run_slots_vs_reference(None, G=256, quota=2, visits=400, level=1, seed=20260, chunk=400,
                                  step_kernel=step_kernel)
    assert r["games"] == 512 and r["samples"] > 5000
    r = ph.run_slots_vs_reference(None, G=512, quota=3, visits=100, level=0, seed=7, chunk=400, step_kernel=step_kernel)
    assert r["games"] == 1536


@pytest.mark.parametrize("G", [1, 33, 448, 2000])
def test_queue_kernel_equals_flat_kernel(G):
    """k_step_q (work queues, state in shared memory) and k_step (thread per game) run the same per-game state machine:
    identical samples, scores, simulation and move counts for every group size (partial groups, one game, several
    groups per CTA), with the NN-free evaluator and fused launches."""
    out = []
    for kern in (b2az.STEP_QUEUE, b2az.STEP_FLAT, b2az.STEP_WAVES, b2az.STEP_SYNC):
        e = ph.make_engine(None, G, G * 2, 60, b2az.EVAL_RANDOM, b2az.RNG_PER_GAME, 99, per_slot_quota=1,
                           step_kernel=kern, history_capacity=G * 2 * 42, **ph.level_params(1))
        for _ in range(10 ** 5):
            e.step(173)
            st = e.stats()
            if st.active_games == 0:
                break
        assert st.device_error == 0 and st.games_completed == 2 * G
        out.append((st.simulations, st.moves, list(st.scores), e.drain_history(G * 2 * 42)))
        e.close()
    assert out[0][:3] == out[1][:3] == out[2][:3] == out[3][:3]
    for i in (0, 2, 3):
        ph.compare_history(out[i][3], out[1][3], ordered=False)


def test_playout_cap_and_resign_gpu():
    """Playout-cap randomisation + resign_percent / playthrough on the device vs the port (coins from each game's own
    stream on both sides): fused RANDOM-eval launches and the lock-step NN loop."""
    extra = dict(playout_cap_randomization=1, playout_cap_depth=25, playout_cap_percent=0.75, resign_percent=0.35,
                 resign_playthrough_percent=0.3)
    r = ph.run_random_parity(None, G=384, games_to_play=10 ** 6, visits=100, seed=12, oracle="port",
                             rng_mode=b2az.RNG_PER_GAME, level=1, chunk=96, steps=3000, extra=extra)
    assert r["games"] > 384 and sum(r["resign_scores"]) > 0 and r["fast_avg_leaf_depth"] > 0
    r = ph.run_lockstep_parity(None, G=32, games_to_play=10 ** 6, visits=40, level=1, seed=12, oracle="port",
                               rng_mode=b2az.RNG_PER_GAME, max_generations=1200, extra=extra)
    assert r["games"] > 32


@pytest.mark.parametrize("level", [3, 4])
def test_gumbel_parallel_kernel_vs_port(level):
    """Gumbel root search on the device (per-game streams) vs the port: fused RANDOM-eval launches incl. Gumbel
    fast searches under playout-cap randomisation, and the lock-step NN loop."""
    extra = dict(fast_search_uses_gumbel=1, playout_cap_randomization=1, playout_cap_depth=16, playout_cap_percent=0.5)
    r = ph.run_random_parity(None, G=256, games_to_play=10 ** 6, visits=64, seed=77, oracle="port",
                             rng_mode=b2az.RNG_PER_GAME, level=level, chunk=64, steps=2500, extra=extra)
    assert r["games"] > 256
    r = ph.run_lockstep_parity(None, G=32, games_to_play=10 ** 6, visits=48, level=level, seed=77, oracle="port",
                               rng_mode=b2az.RNG_PER_GAME, max_generations=1500)
    assert r["games"] > 32


@needs_ref
def test_gumbel_serial_kernel_vs_reference_live():
    r = ph.run_lockstep_parity(None, G=5, games_to_play=8, visits=48, level=4, seed=31, oracle="ref")
    assert r["games"] == 8


def test_no_tree_reuse_gpu():
    ph.run_random_parity(None, G=64, games_to_play=10 ** 6, visits=50, seed=3, oracle="port", level=2, tree_reuse=False,
                         steps=3000)
    # and to the very end in the serial (reference-order) mode
    ph.run_random_parity(None, G=8, games_to_play=20, visits=50, seed=3, oracle="port", level=2, tree_reuse=False,
                         rng_mode=b2az.RNG_GLOBAL)


def test_c4_kernels_random_walks_vs_port():
    Pt = ph.port_lib()
    rng = np.random.default_rng(11)
    boards, players, turns = [], [], []
    for game in range(2000):
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
    out = b2az.c4_batch(boards, np.array(players, np.uint8), np.array(turns, np.uint32), moves)
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


def test_device_zero_copy_path_matches_host_path():
    """b2az_leaf_batch / b2az_submit_eval (device pointers, the DLPack-style feed) must drive the engine to the
    same result as the legacy host-buffer calls."""
    import torch

    kw = ph.level_params(1)
    mk = lambda: ph.make_engine(None, 64, 10 ** 6, 32, b2az.EVAL_NN, b2az.RNG_PER_GAME, 17, **kw)
    a, b = mk(), mk()
    mix = torch.from_numpy(ph._MIX).cuda()
    for _ in range(1500):
        a.step(1)
        b.step(1)
        ids, canon = a.leaf_batch_host()
        n, cptr, iptr = b.leaf_batch()
        assert n == len(ids)
        if n == 0:
            break
        v, pi = ph.fake_net(canon)
        a.submit_eval_host(ids, v, pi)
        # evaluate b's batch on the device, rows in b's own order
        cb = torch.empty((n, 168), dtype=torch.float32, device="cuda")
        C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(cb.data_ptr()), C.c_void_p(cptr), C.c_size_t(n * 168 * 4), 3)
        h = (cb.double() @ mix.double()).to(torch.int64)  # integers < 2^53: exact in any summation order
        wp = (1 + (h[:, :7] % 13) ** 2).to(torch.float32)
        wv = (1 + (h[:, 7:] % 17)).to(torch.float32)
        pi_d = (wp / wp.sum(1, keepdim=True)).contiguous()
        v_d = (wv / wv.sum(1, keepdim=True)).contiguous()
        # float32 sums of <= 7 small integers are exact, so the device evaluator equals fake_net bit for bit
        torch.cuda.synchronize()
        b.submit_eval(v_d.data_ptr(), pi_d.data_ptr(), n)
        b._keep = (v_d, pi_d)
    sa, sb = a.stats(), b.stats()
    assert sa.games_completed == sb.games_completed > 64 and list(sa.scores) == list(sb.scores)
    ph.compare_history(a.drain_history(1 << 16), b.drain_history(1 << 16), ordered=False)
    a.close()
    b.close()


def test_sync_free_generation_loop_matches_host_path():
    """b2az_leaf_batch_device / b2az_submit_eval_all: no host synchronisation per generation (the evaluator runs
    on all concurrent_games rows, the row count stays on the device). Same games as the host-buffer path; with
    the position cache on in both."""
    import torch

    G = 96
    kw = ph.level_params(1)
    mk = lambda: ph.make_engine(None, G, G, 48, b2az.EVAL_NN, b2az.RNG_PER_GAME, 23, max_cache_size=100000,
                                history_capacity=G * 42, **kw)
    a, b = mk(), mk()
    mix = torch.from_numpy(ph._MIX).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    # host path to the end
    while True:
        a.step(1)
        ids, canon = a.leaf_batch_host()
        if len(ids) == 0:
            break
        v, pi = ph.fake_net(canon)
        a.submit_eval_host(ids, v, pi)
    # device path: fixed number of generations, everything stream ordered
    v_d = torch.empty((G, 3), dtype=torch.float32, device="cuda")
    pi_d = torch.empty((G, 7), dtype=torch.float32, device="cuda")
    cudart = C.CDLL("libcudart.so")
    cb = torch.zeros((G, 168), dtype=torch.float32, device="cuda")
    for gen in range(4000):
        b.step(1, stream)
        cptr, iptr, nptr = b.leaf_batch_device(stream)
        cudart.cudaMemcpyAsync(C.c_void_p(cb.data_ptr()), C.c_void_p(cptr), C.c_size_t(G * 168 * 4), 3, C.c_void_p(stream))
        h = (cb.double() @ mix.double()).to(torch.int64)
        wp = (1 + (h[:, :7] % 13) ** 2).to(torch.float32)
        wv = (1 + (h[:, 7:] % 17)).to(torch.float32)
        pi_d.copy_(wp / wp.sum(1, keepdim=True))
        v_d.copy_(wv / wv.sum(1, keepdim=True))
        b.submit_eval_all(v_d.data_ptr(), pi_d.data_ptr())
        if gen % 64 == 63 and b.stats(stream).active_games == 0:
            break
    sa, sb = a.stats(), b.stats(stream)
    assert sb.active_games == 0 and sb.device_error == 0
    assert sa.games_completed == sb.games_completed == G and list(sa.scores) == list(sb.scores)
    assert sa.simulations == sb.simulations and sb.cache_hits > 0
    ph.compare_history(a.drain_history(G * 42), b.drain_history(G * 42), ordered=False)
    a.close()
    b.close()


def test_full_size_properties():
    """BASELINE.json configs[1] size (65,536 concurrent games, 400 sims/move): properties that do not need the
    oracle — simulation/move accounting, legal finished samples, pool accounting, determinism across runs."""
    G = 65536

    def run(steps):
        e = ph.make_engine(None, G, 2 * G, 400, b2az.EVAL_RANDOM, b2az.RNG_PER_GAME, 1, history=True,
                           history_capacity=G * 4, **ph.level_params(0))
        e.step(steps)
        st = e.stats()
        hist = e.drain_history(G * 4)
        e.close()
        return st, hist

    st, (canon, v, pi) = run(1201)
    assert st.device_error == 0
    assert st.simulations == G * 1200, "every active slot finishes exactly one simulation per step"
    assert st.moves == G * 3, "a move every 400 simulations"
    assert st.games_completed == 0 and len(canon) == 0
    st2, _ = run(1201)
    assert (st2.simulations, st2.moves, st2.pool_pages_free) == (st.simulations, st.moves, st.pool_pages_free)


def test_engine_driven_from_a_fresh_thread_on_another_device():
    """The CUDA current device is per host thread (a new Python thread starts on device 0): every C-ABI entry point binds
    the engine's device itself. Create the engine on the LAST device and drive it — steps, leaf batches, evaluations,
    history, statistics, peeks — from a thread that never called cudaSetDevice."""
    import threading

    import torch

    dev = torch.cuda.device_count() - 1
    eng = ph.make_engine(None, 96, 96, 24, b2az.EVAL_NN, b2az.RNG_PER_GAME, 5, device=dev, history_capacity=96 * 42,
                         max_cache_size=50000, **ph.level_params(1))
    out = {}

    def drive():
        try:
            while True:
                eng.step(1)
                ids, canon = eng.leaf_batch_host()
                if len(ids) == 0:
                    break
                v, pi = ph.fake_net(canon)
                eng.submit_eval_host(ids, v, pi)
            out["peek"] = eng.peek(0, 0)
            out["stats"] = eng.stats()
            out["hist"] = eng.drain_history(96 * 42)
        except Exception as ex:  # noqa: BLE001
            out["error"] = repr(ex)

    t = threading.Thread(target=drive)
    t.start()
    t.join(timeout=300)
    assert not t.is_alive() and "error" not in out, out.get("error")
    assert out["stats"].games_completed == 96 and out["stats"].device_error == 0 and len(out["hist"][0]) > 500
    eng.close()
    ref = ph.make_engine(None, 96, 96, 24, b2az.EVAL_NN, b2az.RNG_PER_GAME, 5, device=0, history_capacity=96 * 42,
                         max_cache_size=50000, **ph.level_params(1))
    while True:
        ref.step(1)
        ids, canon = ref.leaf_batch_host()
        if len(ids) == 0:
            break
        v, pi = ph.fake_net(canon)
        ref.submit_eval_host(ids, v, pi)
    ph.compare_history(out["hist"], ref.drain_history(96 * 42), ordered=False)
    ref.close()
