"""Shared parity drivers: run the B200 engine (through the C ABI) and an oracle side by side.

Oracles (test infrastructure, never the product):
  PortPM  oracle/_ref/libazoracle.so  — the CPU restatement (oracle/az_oracle.cc); both RNG modes
  RefPM   oracle/_ref/libazref.so     — the UNMODIFIED reference compiled from /root/reference
                                        (tests/refdriver.py); global RNG only (mcts.cc:19-21)
Evaluator: `fake_net` — a pure integer-hash function of the canonical planes, exact in float32 and
independent of batch composition, so every engine sees bit-identical (v, pi) for the same position
(the role DeterministicAgent plays in the reference's src/test_cache.py:173-187).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import b2az  # noqa: E402
import refdriver  # noqa: E402

PORT_LIB = os.path.join(ROOT, "oracle", "_ref", "libazoracle.so")
HOSTEMU_LIB = os.path.join(ROOT, "tests", "cpp", "libb2az_hostemu.so")


# ------------------------------------------------------------------------------------ evaluator
def _mix_matrix():
    m = np.empty((168, 10), np.int64)
    x = 0x9E3779B97F4A7C15
    for i in range(168):
        for j in range(10):
            x = (x * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            m[i, j] = (x >> 40) % 1000003
    return m


_MIX = _mix_matrix()


def fake_net(canon):
    """canon float32[B,4,6,7] (0/1 planes) -> (v float32[B,3], pi float32[B,7]); rows sum to 1."""
    x = np.ascontiguousarray(canon, np.float32).reshape(len(canon), 168).astype(np.int64)
    h = x @ _MIX  # exact integer arithmetic: identical for any batch split
    wp = (1 + (h[:, :7] % 13) ** 2).astype(np.float32)
    wv = (1 + (h[:, 7:] % 17)).astype(np.float32)
    pi = wp / wp.sum(axis=1, keepdims=True, dtype=np.float32)
    v = wv / wv.sum(axis=1, keepdims=True, dtype=np.float32)
    return v.astype(np.float32), pi.astype(np.float32)


# ------------------------------------------------------------------------------------ configs
def level_params(level):
    """SURVEY.md §8(d) parity levels. 0 = plain PUCT; 1 = connect4.yaml self-play settings."""
    if level == 0:
        return dict(cpuct=1.25, fpu_reduction=0.25, epsilon=0.0, mcts_root_temp=1.0, start_temp=1.0, final_temp=1.0,
                    temp_decay_half_life=0.0, root_fpu_zero=0, shaped_dirichlet=0, policy_target_pruning=0)
    if level == 1:
        return dict(cpuct=1.25, fpu_reduction=0.25, epsilon=0.25, mcts_root_temp=1.25, start_temp=1.0, final_temp=0.2,
                    temp_decay_half_life=10.0, root_fpu_zero=1, shaped_dirichlet=1, policy_target_pruning=1)
    if level == 2:  # plain (unshaped) Dirichlet, no root temperature, fast temperature decay
        return dict(cpuct=2.0, fpu_reduction=0.0, epsilon=0.4, mcts_root_temp=1.0, start_temp=1.5, final_temp=0.3,
                    temp_decay_half_life=3.0, root_fpu_zero=0, shaped_dirichlet=0, policy_target_pruning=1)
    if level == 3:  # Gumbel AlphaZero root search (+ improved-policy targets), no Dirichlet
        return dict(cpuct=1.25, fpu_reduction=0.25, epsilon=0.0, mcts_root_temp=1.0, start_temp=1.0, final_temp=1.0,
                    temp_decay_half_life=0.0, root_fpu_zero=0, shaped_dirichlet=0, policy_target_pruning=0,
                    gumbel_enabled=1, gumbel_m=16, gumbel_c_visit=50.0, gumbel_c_scale=1.0)
    if level == 4:  # Gumbel everywhere: interior pi'-matching, m < legal moves, epsilon > 0 (noise only on reused roots)
        return dict(cpuct=1.25, fpu_reduction=0.25, epsilon=0.25, mcts_root_temp=1.25, start_temp=1.0, final_temp=0.2,
                    temp_decay_half_life=10.0, root_fpu_zero=1, shaped_dirichlet=1, policy_target_pruning=1,
                    gumbel_enabled=1, gumbel_m=4, gumbel_c_visit=50.0, gumbel_c_scale=0.5, gumbel_full=1)
    raise ValueError(level)


# ------------------------------------------------------------------------------------ oracle: port
class AzoCfg(C.Structure):
    _fields_ = [
        ("games_to_play", C.c_uint32), ("concurrent_games", C.c_uint32), ("mcts_visits", C.c_uint32 * 2),
        ("cpuct", C.c_float), ("start_temp", C.c_float), ("final_temp", C.c_float), ("temp_decay_half_life", C.c_float),
        ("history_enabled", C.c_uint8), ("tree_reuse", C.c_uint8), ("root_fpu_zero", C.c_uint8),
        ("shaped_dirichlet", C.c_uint8), ("policy_target_pruning", C.c_uint8), ("eval_type", C.c_uint8),
        ("rng_mode", C.c_uint8), ("pad_", C.c_uint8), ("epsilon", C.c_float), ("mcts_root_temp", C.c_float),
        ("fpu_reduction", C.c_float), ("seed", C.c_uint64),
        ("playout_cap_randomization", C.c_uint8), ("pad2_", C.c_uint8 * 3), ("playout_cap_depth", C.c_uint32),
        ("playout_cap_percent", C.c_float), ("resign_percent", C.c_float), ("resign_playthrough_percent", C.c_float),
        ("gumbel_enabled", C.c_uint8), ("gumbel_full", C.c_uint8), ("fast_search_uses_gumbel", C.c_uint8),
        ("pad3_", C.c_uint8), ("gumbel_m", C.c_uint32), ("gumbel_c_visit", C.c_float), ("gumbel_c_scale", C.c_float),
    ]


_port = None


def port_lib():
    global _port
    if _port is None:
        L = C.CDLL(PORT_LIB)
        vp, u32 = C.c_void_p, C.c_uint32
        L.azo_pm_new.restype = vp
        L.azo_pm_new.argtypes = [C.POINTER(AzoCfg)]
        L.azo_pm_free.argtypes = [vp]
        L.azo_pm_run.argtypes = [vp]
        L.azo_pm_run.restype = u32
        L.azo_pm_run_iterations.argtypes = [vp, C.c_uint64]
        L.azo_pm_run_iterations.restype = u32
        L.azo_pm_build_batch.argtypes = [vp, u32, vp, vp]
        L.azo_pm_build_batch.restype = u32
        L.azo_pm_update_inferences.argtypes = [vp, vp, u32, vp, vp]
        L.azo_pm_drain_history.argtypes = [vp, u32, vp, vp, vp]
        L.azo_pm_drain_history.restype = u32
        for n in ("azo_pm_hist_count", "azo_pm_games_completed", "azo_pm_remaining_games"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = u32
        for n in ("azo_pm_simulations", "azo_pm_moves"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = C.c_uint64
        L.azo_pm_scores.argtypes = [vp, vp]
        L.azo_pm_resign_scores.argtypes = [vp, vp]
        L.azo_pm_metrics.argtypes = [vp, vp]
        L.azo_pm_peek.argtypes = [vp, u32, u32, vp, vp, vp, vp, C.POINTER(u32), C.POINTER(u32), vp]
        L.azo_c4_play.argtypes = [vp, vp, vp, u32]
        L.azo_c4_valid.argtypes = [vp, vp]
        L.azo_c4_scores.argtypes = [vp, vp]
        L.azo_c4_canonical.argtypes = [vp, C.c_uint8, vp]
        L.azo_rng_new.restype = vp
        L.azo_rng_new.argtypes = [C.c_uint64, C.c_int, C.c_uint64]
        L.azo_rng_free.argtypes = [vp]
        L.azo_rng_u32.argtypes = [vp]
        L.azo_rng_u32.restype = u32
        L.azo_rng_shuffle.argtypes = [vp, u32, vp]
        L.azo_rng_uniform01.argtypes = [vp]
        L.azo_rng_uniform01.restype = C.c_float
        L.azo_rng_gamma.argtypes = [vp, C.c_float, u32, vp]
        _port = L
    return _port


def _P(a):
    return a.ctypes.data_as(C.c_void_p)


METRIC_NAMES = ["avg_game_length", "avg_leaf_depth", "avg_search_entropy", "fast_avg_leaf_depth",
                "fast_avg_search_entropy", "avg_moves_per_turn", "avg_valid_moves"]


class PortPM:
    """oracle/az_oracle.cc behind the same small interface as refdriver.RefPlayManager."""

    def __init__(self, G, games_to_play, visits, eval_type, rng_mode, seed, history=True, tree_reuse=True, **kw):
        self.L = port_lib()
        c = AzoCfg(games_to_play=games_to_play, concurrent_games=G, history_enabled=int(history),
                   tree_reuse=int(tree_reuse), eval_type=eval_type, rng_mode=rng_mode, seed=seed, **kw)
        c.mcts_visits[0] = c.mcts_visits[1] = visits
        self.h = self.L.azo_pm_new(C.byref(c))
        assert self.h
        self.G = G

    def close(self):
        if self.h:
            self.L.azo_pm_free(self.h)
            self.h = None

    def advance(self):
        """Drain awaiting_mcts_. Returns False when the run is over."""
        self.L.azo_pm_run(self.h)
        return self.L.azo_pm_remaining_games(self.h) > 0

    def run_iterations(self, n):
        self.L.azo_pm_run_iterations(self.h, n)

    def build_batch(self):
        ids = np.empty(self.G, np.uint32)
        canon = np.empty((self.G, 4, 6, 7), np.float32)
        n = self.L.azo_pm_build_batch(self.h, self.G, _P(ids), _P(canon))
        return ids[:n], canon[:n]

    def update_inferences(self, ids, v, pi):
        ids = np.ascontiguousarray(ids, np.uint32)
        v = np.ascontiguousarray(v, np.float32)
        pi = np.ascontiguousarray(pi, np.float32)
        self.L.azo_pm_update_inferences(self.h, _P(ids), len(ids), _P(v), _P(pi))

    def drain_history(self, max_rows=1 << 20):
        n = min(max_rows, self.L.azo_pm_hist_count(self.h))
        canon = np.empty((n, 4, 6, 7), np.float32)
        v = np.empty((n, 3), np.float32)
        pi = np.empty((n, 7), np.float32)
        got = self.L.azo_pm_drain_history(self.h, n, _P(canon), _P(v), _P(pi)) if n else 0
        return canon[:got], v[:got], pi[:got]

    def scores(self):
        s = np.zeros(3, np.float32)
        self.L.azo_pm_scores(self.h, _P(s))
        return s

    def resign_scores(self):
        s = np.zeros(3, np.float32)
        self.L.azo_pm_resign_scores(self.h, _P(s))
        return s

    def metrics(self):
        m = np.zeros(7, np.float32)
        self.L.azo_pm_metrics(self.h, _P(m))
        return dict(zip(METRIC_NAMES, m.tolist()))

    def games_completed(self):
        return self.L.azo_pm_games_completed(self.h)

    def simulations(self):
        return self.L.azo_pm_simulations(self.h)

    def moves(self):
        return self.L.azo_pm_moves(self.h)

    def peek(self, game, seat):
        state = np.zeros(89, np.uint8)
        counts = np.zeros(7, np.uint32)
        q = np.zeros(7, np.float32)
        pol = np.zeros(7, np.float32)
        rv = np.zeros(3, np.float32)
        depth, root_n = C.c_uint32(), C.c_uint32()
        self.L.azo_pm_peek(self.h, game, seat, _P(state), _P(counts), _P(q), _P(rv), C.byref(depth), C.byref(root_n),
                           _P(pol))
        return dict(state=state, counts=counts, q=q, policy=pol, root_value=rv, depth=depth.value, root_n=root_n.value)


class RefPM:
    """The unmodified reference PlayManager (global thread-local RNG, one worker thread)."""

    def __init__(self, G, games_to_play, visits, eval_type, rng_mode, seed, history=True, tree_reuse=True, **kw):
        assert rng_mode == b2az.RNG_GLOBAL, "the reference only has the global thread-local stream"
        cfg = refdriver.play_cfg(games_to_play=games_to_play, concurrent_games=G, max_batch_size=G,
                                 mcts_visits=(visits, visits), history_enabled=int(history), self_play=1,
                                 tree_reuse=int(tree_reuse), eval_type=eval_type, **kw)
        self.pm = refdriver.RefPlayManager(cfg)
        self.eval_type = eval_type
        self.seed = seed
        self.G = G
        self.started = False

    def close(self):
        self.pm.close()

    def advance(self):
        if self.eval_type == b2az.EVAL_RANDOM:
            self.pm.play_here(self.seed)
            return False
        if not self.started:
            self.pm.start_workers(1, self.seed, True)
            self.started = True
        st = self.pm.wait_quiescent(60000)
        assert st >= 0, "reference lock-step harness timed out"
        if st == 0:
            self.pm.join()
        return st == 1

    def build_batch(self):
        return self.pm.build_batch(0, self.G)

    def update_inferences(self, ids, v, pi):
        self.pm.update_inferences(ids, v, pi)

    def drain_history(self, max_rows=1 << 20):
        n = self.pm.L.azref_pm_hist_count(self.pm.h)
        return self.pm.drain_history(max(1, min(n, max_rows))) if n else (np.empty((0, 4, 6, 7), np.float32),
                                                                           np.empty((0, 3), np.float32),
                                                                           np.empty((0, 7), np.float32))

    def scores(self):
        return self.pm.scores()

    def metrics(self):
        return self.pm.metrics()

    def games_completed(self):
        return self.pm.games_completed()

    def peek(self, game, seat):
        return self.pm.peek(game, seat)


def make_oracle(kind, **kw):
    return {"port": PortPM, "ref": RefPM}[kind](**kw)


def make_engine(engine_lib, G, games_to_play, visits, eval_type, rng_mode, seed, history=True, tree_reuse=True,
                lanes=0, device=0, **kw):
    lib = b2az.load(engine_lib) if engine_lib else b2az.load()
    p = b2az.default_params(lib, games_to_play=games_to_play, concurrent_games=G, mcts_visits=(visits, visits),
                            history_enabled=int(history), self_play=1, tree_reuse=int(tree_reuse), eval_type=eval_type,
                            rng_mode=rng_mode, seed=seed, lanes_per_game=lanes, **kw)
    return b2az.Engine(p, device=device, lib=lib)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _sorted_rows(canon, v, pi):
    rows = np.concatenate([_bits(canon).reshape(len(canon), -1), _bits(v), _bits(pi)], axis=1)
    order = np.lexsort(rows.T[::-1])
    return rows[order]


def compare_history(eng_hist, ora_hist, ordered):
    ce, ve, pe = eng_hist
    co, vo, po = ora_hist
    assert len(ce) == len(co), f"history length {len(ce)} != {len(co)}"
    if ordered:
        assert np.array_equal(_bits(ce), _bits(co)), "history canonicals differ"
        assert np.array_equal(_bits(ve), _bits(vo)), "history values differ"
        assert np.array_equal(_bits(pe), _bits(po)), "history policy targets differ (bit-exact)"
    else:
        assert np.array_equal(_sorted_rows(ce, ve, pe), _sorted_rows(co, vo, po)), "history sample multisets differ"


def compare_peek(pe, po, where):
    assert np.array_equal(pe["state"], po["state"]), f"{where}: game state differs"
    assert np.array_equal(pe["counts"], po["counts"]), f"{where}: visit counts differ {pe['counts']} vs {po['counts']}"
    assert pe["depth"] == po["depth"] and pe["root_n"] == po["root_n"], f"{where}: depth/root_n differ"
    assert np.array_equal(_bits(pe["q"]), _bits(po["q"])), f"{where}: root Q differ"
    assert np.array_equal(_bits(pe["root_value"]), _bits(po["root_value"])), f"{where}: root_value differs"


def run_lockstep_parity(engine_lib, G, games_to_play, visits, level, seed, oracle="port", rng_mode=None,
                        tree_reuse=True, lanes=0, peek_every=7, max_generations=10 ** 7, compact_pages=0, extra=None):
    """NN-eval lock-step run (SURVEY.md Appendix A 'Scheduling order'): every generation both sides expose
    their leaf batch (ids + canonical planes, compared bit for bit), get the same fake_net answers and
    advance. Visit counts / Q / root values are peeked every `peek_every` generations and the finished
    training samples (root canonical, final score, policy target) are compared at the end."""
    if rng_mode is None:
        rng_mode = b2az.RNG_GLOBAL if oracle == "ref" else b2az.RNG_PER_GAME
    ordered = rng_mode == b2az.RNG_GLOBAL or engine_lib is not None
    kw = dict(level_params(level), **(extra or {}))
    eng = make_engine(engine_lib, G, games_to_play, visits, b2az.EVAL_NN, rng_mode, seed, tree_reuse=tree_reuse,
                      lanes=lanes, compact_pages=compact_pages, **kw)
    ora = make_oracle(oracle, G=G, games_to_play=games_to_play, visits=visits, eval_type=b2az.EVAL_NN,
                      rng_mode=rng_mode, seed=seed, tree_reuse=tree_reuse, **kw)
    gens = 0
    leaves = 0
    try:
        while gens < max_generations:
            more = ora.advance()
            eng.step(1)
            ids_e, canon_e = eng.leaf_batch_host()
            if not more:
                assert len(ids_e) == 0, "engine still has leaves after the oracle finished"
                break
            ids_o, canon_o = ora.build_batch()
            if ordered:
                assert np.array_equal(ids_e, ids_o), f"generation {gens}: leaf ids differ"
                assert np.array_equal(_bits(canon_e), _bits(canon_o)), f"generation {gens}: leaf canonicals differ"
            else:  # parallel kernel: batch row order is whatever the atomics gave; compare per slot
                oe, oo = np.argsort(ids_e, kind="stable"), np.argsort(ids_o, kind="stable")
                assert np.array_equal(ids_e[oe], ids_o[oo]), f"generation {gens}: leaf id sets differ"
                assert np.array_equal(_bits(canon_e[oe]), _bits(canon_o[oo])), \
                    f"generation {gens}: leaf canonicals differ"
            if peek_every and gens % peek_every == 0:  # while the oracle is quiescent
                g = int(ids_o[(gens // peek_every) % len(ids_o)])
                for seat in (0, 1):
                    compare_peek(eng.peek(g, seat), ora.peek(g, seat), f"generation {gens} game {g} seat {seat}")
            v, pi = fake_net(canon_o)
            ora.update_inferences(ids_o, v, pi)
            ve, pie = fake_net(canon_e)  # pure function of the position: same answers, engine row order
            eng.submit_eval_host(ids_e, ve, pie)
            leaves += len(ids_o)
            gens += 1
        st = eng.stats()
        assert st.device_error == 0, f"device error bits {st.device_error}"
        assert st.games_completed == ora.games_completed()
        if gens < max_generations:
            assert st.games_completed == games_to_play
        assert np.array_equal(np.array(st.scores[:], np.float32), ora.scores()), "scores differ"
        if hasattr(ora, "resign_scores"):
            assert np.array_equal(np.array(st.resign_scores[:], np.float32), ora.resign_scores()), "resign scores differ"
        if st.games_completed == 0:
            return dict(generations=gens, leaves=leaves, games=0, moves_compared=0, scores=[0, 0, 0],
                        simulations=int(st.simulations), compactions=int(st.compactions))
        m = ora.metrics()
        for name in ("avg_game_length", "avg_moves_per_turn", "avg_valid_moves"):
            assert np.float32(getattr(st, name)) == np.float32(m[name]), name
        for name in ("avg_leaf_depth", "avg_search_entropy"):
            if ordered:
                assert np.float32(getattr(st, name)) == np.float32(m[name]), name
            else:
                assert abs(getattr(st, name) - m[name]) <= 1e-5 * max(1.0, abs(m[name])), name
        he = eng.drain_history(1 << 20)
        ho = ora.drain_history(1 << 20)
        compare_history(he, ho, ordered)
        return dict(generations=gens, leaves=leaves, games=int(st.games_completed), moves_compared=len(ho[0]),
                    scores=[float(x) for x in st.scores[:]], simulations=int(st.simulations),
                    compactions=int(st.compactions))
    finally:
        eng.close()
        ora.close()


def run_random_parity(engine_lib, G, games_to_play, visits, seed, oracle="port", rng_mode=None, level=0,
                      tree_reuse=True, lanes=0, chunk=64, steps=None, compact_pages=0, pool_nodes=0, ordered=None,
                      extra=None):
    """RANDOM-eval run (EvalType::RANDOM, the reference's own fake backend: play_manager_test.cc): the engine
    fuses `chunk` loop iterations per launch; the oracle plays to the end; final scores, metrics and the
    training samples must agree."""
    if rng_mode is None:
        rng_mode = b2az.RNG_GLOBAL if oracle == "ref" else b2az.RNG_PER_GAME
    if ordered is None:
        ordered = rng_mode == b2az.RNG_GLOBAL or engine_lib is not None
    kw = dict(level_params(level), **(extra or {}))
    eng = make_engine(engine_lib, G, games_to_play, visits, b2az.EVAL_RANDOM, rng_mode, seed, tree_reuse=tree_reuse,
                      lanes=lanes, history_capacity=max(1 << 16, games_to_play * 42), compact_pages=compact_pages,
                      pool_nodes=pool_nodes, **kw)
    ora = make_oracle(oracle, G=G, games_to_play=games_to_play, visits=visits, eval_type=b2az.EVAL_RANDOM,
                      rng_mode=rng_mode, seed=seed, tree_reuse=tree_reuse, **kw)
    try:
        if steps is None:  # play every game to the end
            ora.advance()
            for _ in range(10 ** 6):
                eng.step(chunk)
                st = eng.stats()
                if st.active_games == 0:
                    break
            assert st.games_completed == games_to_play
        else:  # a fixed number of generations while no slot can retire (order-independent: see DESIGN.md)
            ora.run_iterations(G * steps)
            done = 0
            while done < steps:
                n = min(chunk, steps - done)
                eng.step(n)
                done += n
            st = eng.stats()
            assert st.active_games == G, "pick games_to_play large enough that no slot retires"
        assert st.device_error == 0, f"device error bits {st.device_error}"
        assert st.games_completed == ora.games_completed()
        assert np.array_equal(np.array(st.scores[:], np.float32), ora.scores()), \
            f"scores differ {st.scores[:]} vs {ora.scores()}"
        m = ora.metrics()
        for name in ("avg_game_length", "avg_moves_per_turn", "avg_valid_moves"):
            assert np.float32(getattr(st, name)) == np.float32(m[name]), name
        for name in ("avg_leaf_depth", "avg_search_entropy", "fast_avg_leaf_depth", "fast_avg_search_entropy"):
            assert abs(getattr(st, name) - m[name]) <= 1e-5 * max(1.0, abs(m[name])), name
        if hasattr(ora, "resign_scores"):
            assert np.array_equal(np.array(st.resign_scores[:], np.float32), ora.resign_scores()), "resign scores differ"
        if hasattr(ora, "simulations"):
            assert st.simulations == ora.simulations() and st.moves == ora.moves()
        he = eng.drain_history(1 << 20)
        ho = ora.drain_history(1 << 20)
        compare_history(he, ho, ordered)
        return dict(games=int(st.games_completed), scores=[float(x) for x in st.scores[:]],
                    simulations=int(st.simulations), moves=int(st.moves), samples=len(ho[0]),
                    avg_game_length=float(st.avg_game_length), avg_leaf_depth=float(st.avg_leaf_depth),
                    compactions=int(st.compactions), resign_scores=[float(x) for x in st.resign_scores[:]],
                    fast_avg_leaf_depth=float(st.fast_avg_leaf_depth))
    finally:
        eng.close()
        ora.close()


def run_slots_vs_reference(engine_lib, G, quota, visits, level, seed, chunk=400, extra=None, step_kernel=0):
    """The fused per-game-RNG kernel pinned DIRECTLY to the unmodified reference: slot g of one engine run
    (per_slot_quota: every slot plays `quota` games) must equal a reference PlayManager with concurrent_games = 1,
    games_to_play = quota, RANDOM eval, run on a thread seeded MCTS::seed_thread_rng(seed + g). Compared: the
    multiset of training samples (canonical, outcome, policy target — bit patterns), scores, total game length /
    move counts (exact), and the leaf-depth / entropy means (double sums added in another order: 1e-5)."""
    kw = dict(level_params(level), **(extra or {}))
    eng = make_engine(engine_lib, G, G * quota, visits, b2az.EVAL_RANDOM, b2az.RNG_PER_GAME, seed,
                      history_capacity=max(1 << 16, G * quota * 42), per_slot_quota=1, step_kernel=step_kernel, **kw)
    try:
        for _ in range(10 ** 6):
            eng.step(chunk)
            st = eng.stats()
            if st.active_games == 0:
                break
        assert st.device_error == 0 and st.games_completed == G * quota
        he = eng.drain_history(1 << 22)
    finally:
        eng.close()
    canon, v, pi = [], [], []
    scores = np.zeros(3, np.float64)
    length = moves = 0.0
    depth_sum = ent_sum = 0.0
    for g in range(G):
        cfg = refdriver.play_cfg(games_to_play=quota, concurrent_games=1, max_batch_size=1,
                                 mcts_visits=(visits, visits), history_enabled=1, self_play=1, tree_reuse=1,
                                 eval_type=b2az.EVAL_RANDOM, **kw)
        pm = refdriver.RefPlayManager(cfg)
        pm.play_here(seed + g)
        c, vv, pp = pm.drain_history(quota * 42 + 1)
        canon.append(c); v.append(vv); pi.append(pp)
        scores += pm.scores()
        m = pm.metrics()
        length += m["avg_game_length"] * quota
        # avg_moves_per_turn == 1 for Connect4, so total_move_count == game_length and every move is a full search here
        depth_sum += m["avg_leaf_depth"] * m["avg_game_length"] * quota
        ent_sum += m["avg_search_entropy"] * m["avg_game_length"] * quota
        moves += m["avg_game_length"] * quota
        pm.close()
    compare_history(he, (np.concatenate(canon), np.concatenate(v), np.concatenate(pi)), ordered=False)
    assert np.array_equal(np.array(st.scores[:], np.float64), scores), f"scores differ {st.scores[:]} vs {scores}"
    assert abs(st.avg_game_length * G * quota - length) < 0.5, "total game length differs"
    assert st.moves == round(moves)
    if not kw.get("playout_cap_randomization"):
        assert abs(st.avg_leaf_depth - depth_sum / moves) <= 1e-5 * max(1.0, depth_sum / moves)
        assert abs(st.avg_search_entropy - ent_sum / moves) <= 1e-5 * max(1.0, ent_sum / moves)
    return dict(games=int(st.games_completed), samples=len(he[0]), simulations=int(st.simulations))


# ------------------------------------------------------------------------------------ golden traces
class EnginePM:
    """The engine behind the oracle-shaped interface, so one tracer serves every side."""

    def __init__(self, engine_lib, G, games_to_play, visits, eval_type, rng_mode, seed, history=True, tree_reuse=True,
                 lanes=0, **kw):
        self.e = make_engine(engine_lib, G, games_to_play, visits, eval_type, rng_mode, seed, history=history,
                             tree_reuse=tree_reuse, lanes=lanes, history_capacity=max(1 << 16, games_to_play * 42), **kw)
        self.eval_type = eval_type
        self.G = G

    def close(self):
        self.e.close()

    def advance(self):
        if self.eval_type == b2az.EVAL_RANDOM:
            while self.e.stats().active_games:
                self.e.step(64)
            return False
        self.e.step(1)
        return self.e.stats().active_games > 0

    def build_batch(self):
        return self.e.leaf_batch_host()

    def update_inferences(self, ids, v, pi):
        self.e.submit_eval_host(ids, v, pi)

    def drain_history(self, max_rows=1 << 20):
        return self.e.drain_history(max_rows)

    def scores(self):
        return np.array(self.e.stats().scores[:], np.float32)

    def metrics(self):
        st = self.e.stats()
        return {k: float(getattr(st, k)) for k in METRIC_NAMES}

    def games_completed(self):
        return self.e.stats().games_completed

    def peek(self, game, seat):
        return self.e.peek(game, seat)


def trace_run(pm, eval_type, peek_every=11):
    """Run `pm` to the end with fake_net and record everything a parity check looks at."""
    import hashlib

    digest = hashlib.sha256()
    peeks = []
    gens = 0
    if eval_type == b2az.EVAL_RANDOM:
        pm.advance()
    else:
        while pm.advance():
            ids, canon = pm.build_batch()
            digest.update(np.ascontiguousarray(ids, np.uint32).tobytes())
            digest.update(np.ascontiguousarray(canon, np.float32).tobytes())
            if gens % peek_every == 0:
                g = int(ids[(gens // peek_every) % len(ids)])
                for seat in (0, 1):
                    pk = pm.peek(g, seat)
                    peeks.append(np.concatenate([[gens, g, seat, pk["depth"], pk["root_n"]], pk["counts"],
                                                 _bits(pk["q"]), _bits(pk["root_value"])]).astype(np.int64))
            v, pi = fake_net(canon)
            pm.update_inferences(ids, v, pi)
            gens += 1
    canon, v, pi = pm.drain_history()
    m = pm.metrics()
    return dict(generations=np.int64(gens), leaf_digest=np.frombuffer(digest.digest(), np.uint8).copy(),
                peeks=np.stack(peeks) if peeks else np.zeros((0, 22), np.int64),
                hist_canon=canon.astype(np.uint8), hist_v=v, hist_pi=pi, scores=pm.scores(),
                metrics=np.array([m[k] for k in METRIC_NAMES], np.float32),
                games_completed=np.int64(pm.games_completed()))


GOLDEN_CASES = {
    # name: (G, games_to_play, visits, level, seed, eval_type)
    "nn_level0": (4, 6, 50, 0, 12345, 0),
    "nn_level1": (4, 6, 50, 1, 12345, 0),
    "nn_level1_100sims": (2, 3, 100, 1, 20240601, 0),   # BASELINE.json configs[0] settings
    "nn_level2": (3, 5, 30, 2, 7, 0),
    "random_level0": (8, 20, 64, 0, 12345, 1),
    "random_level1": (8, 16, 40, 1, 99, 1),
    "nn_level3_gumbel": (4, 6, 50, 3, 4242, 0),
    "nn_level4_gumbel_full": (3, 5, 40, 4, 777, 0),
    "random_level3_gumbel": (8, 16, 48, 3, 31, 1),
}


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", f"c4_{name}.npz")


def compare_trace(got, want, what):
    for k in want:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, f"{what}: {k} shape {a.shape} vs {b.shape}"
        if a.dtype.kind == "f":
            assert np.array_equal(a.astype(np.float32).view(np.uint32), b.astype(np.float32).view(np.uint32)), \
                f"{what}: {k} differs (bit-exact)"
        else:
            assert np.array_equal(a, b), f"{what}: {k} differs"
