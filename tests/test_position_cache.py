"""The device position cache (replaces S3FIFOCache / ShardedS3FIFOCache, src/s3fifo_cache.h) inside the NN
evaluation loop. The reference's own end-to-end property (src/test_cache.py:227-253): with a deterministic
evaluator a search with the cache on is IDENTICAL to the search with the cache off. Plus the counter semantics:
every leaf — terminal ones included — is looked up (play_manager.cc:589-598), so hits + misses == simulations
(what src/network_pareto.py:415-423 uses as the simulation counter).
CPU: host-emulation build running the fused kernel's loop game by game (B2AZ_EMU_FLAT=1); GPU: libb2az.so."""
import numpy as np
import pytest

import b2az
import parity_harness as ph
from conftest import has_cuda

LIBS = [pytest.param(ph.HOSTEMU_LIB, id="host-emulation"),
        pytest.param(None, id="cuda", marks=[pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")])]


def _play_out(lib_path, G, visits, cache, level=1, seed=21):
    """Every slot plays exactly ONE game (games_to_play == concurrent_games), so the set of training samples does not
    depend on how far each slot gets per generation."""
    eng = ph.make_engine(lib_path, G, G, visits, b2az.EVAL_NN, b2az.RNG_PER_GAME, seed, history_capacity=G * 42,
                         max_cache_size=cache, **ph.level_params(level))
    gens = 0
    while True:
        eng.step(1)
        ids, canon = eng.leaf_batch_host()
        if len(ids) == 0:
            break
        v, pi = ph.fake_net(canon)
        eng.submit_eval_host(ids, v, pi)
        gens += 1
    st = eng.stats()
    hist = eng.drain_history(G * 42)
    eng.close()
    return st, hist, gens


@pytest.mark.parametrize("lib_path", LIBS)
def test_cache_on_equals_cache_off(lib_path, monkeypatch):
    monkeypatch.setenv("B2AZ_EMU_FLAT", "1")
    G, visits = (24, 40) if lib_path else (256, 64)
    off, h_off, gens_off = _play_out(lib_path, G, visits, cache=0)
    on, h_on, gens_on = _play_out(lib_path, G, visits, cache=4000000)  # sparse: only set conflicts can evict
    assert off.device_error == 0 and on.device_error == 0
    assert on.games_completed == off.games_completed == G
    assert list(on.scores) == list(off.scores) and on.simulations == off.simulations and on.moves == off.moves
    ph.compare_history(h_on, h_off, ordered=False)
    assert on.cache_hits > 0 and on.cache_hits + on.cache_misses == on.simulations
    assert off.cache_hits == off.cache_misses == 0 and off.cache_max_size == 0
    assert 0 < on.cache_size <= on.cache_max_size == 4000000
    assert on.cache_evictions <= on.cache_size // 100  # 4-way buckets at < 4 % load: set conflicts are rare
    assert gens_on < gens_off, "hits must save evaluator round trips"


@pytest.mark.parametrize("lib_path", LIBS)
def test_tiny_cache_evicts_and_stays_correct(lib_path, monkeypatch):
    monkeypatch.setenv("B2AZ_EMU_FLAT", "1")
    G, visits = (16, 32) if lib_path else (128, 48)
    off, h_off, _ = _play_out(lib_path, G, visits, cache=0, level=0, seed=5)
    on, h_on, _ = _play_out(lib_path, G, visits, cache=64, level=0, seed=5)
    ph.compare_history(h_on, h_off, ordered=False)
    assert list(on.scores) == list(off.scores)
    assert on.cache_max_size == 64 and on.cache_size <= 64 and on.cache_evictions > 0
    assert on.cache_hits + on.cache_misses == on.simulations
    assert on.cache_reinserts <= on.cache_misses


def test_cache_rejected_in_parity_mode():
    lib = b2az.load(ph.HOSTEMU_LIB)
    with pytest.raises(b2az.B2azError, match="parity"):
        b2az.Engine(b2az.default_params(lib, max_cache_size=100, rng_mode=b2az.RNG_GLOBAL), lib=lib)


@pytest.mark.parametrize("lib_path", LIBS)
def test_overlapped_history_drain_equals_the_plain_drain(lib_path):
    """b2az_history_mark + b2az_drain_history_marked (samples of step k leave on a second stream while step k + 1 runs)
    deliver exactly the samples, in exactly the order, of b2az_drain_history after every step."""
    G = 64 if lib_path else 2048
    kw = dict(history_capacity=G * 42 * 2, per_slot_quota=1, **ph.level_params(1))
    a = ph.make_engine(lib_path, G, 3 * G, 40, b2az.EVAL_RANDOM, b2az.RNG_PER_GAME, 77, **kw)
    b = ph.make_engine(lib_path, G, 3 * G, 40, b2az.EVAL_RANDOM, b2az.RNG_PER_GAME, 77, **kw)
    s1 = s2 = None
    if lib_path is None:
        import torch

        t1, t2 = torch.cuda.Stream(), torch.cuda.Stream()
        s1, s2 = t1.cuda_stream, t2.cuda_stream
    cap = G * 42
    bufs = [np.zeros((cap, 4, 6, 7), np.float32), np.zeros((cap, 3), np.float32), np.zeros((cap, 7), np.float32)]
    got_a, got_b = [], []
    marked = False
    for it in range(10 ** 5):
        a.step(57)
        h = a.drain_history(cap)
        if len(h[0]):
            got_a.append(h)
        b.step(57, s1)          # step k + 1 is enqueued ...
        if marked:              # ... while the samples of step k are drained on the second stream
            n = b.drain_history_marked_into(*[x.ctypes.data for x in bufs], cap, s2)
            if n:
                got_b.append(tuple(x[:n].copy() for x in bufs))
        b.history_mark(s1)
        marked = True
        if a.stats().active_games == 0 and b.stats(s1).active_games == 0:
            break
    n = b.drain_history_marked_into(*[x.ctypes.data for x in bufs], cap, s2)
    if n:
        got_b.append(tuple(x[:n].copy() for x in bufs))
    assert a.stats().games_completed == b.stats(s1).games_completed == 3 * G and b.stats(s1).device_error == 0
    cat = lambda parts: tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
    ph.compare_history(cat(got_a), cat(got_b), ordered=lib_path is not None)  # (atomics order the ring on the device)
    a.close()
    b.close()
