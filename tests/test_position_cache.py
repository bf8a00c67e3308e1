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
