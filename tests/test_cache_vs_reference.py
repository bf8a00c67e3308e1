"""The position caches against the reference's S3FIFOCache (src/s3fifo_cache.h, the container behind
ShardedS3FIFOCache; its 26 gtest scenarios in src/s3fifo_cache_test.cc pass against oracle/_ref, see
tests/test_oracle_vs_reference.py) on recorded insert / find traces:

  * the module's host classes S3FIFOCache / ShardedS3FIFOCache (csrc/py_s3fifo.h): every find() answer and all six
    counters identical, under heavy eviction pressure (tiny capacities), with and without a ghost list;
  * the DEVICE table of the engine (4-way set associative, csrc/az_engine_logic.h cache_find / cache_insert) through
    b2az_cache_insert_host / b2az_cache_find_host: identical answers and counters while no set overflows (the regime
    of self-play: a 200 k-entry table, a few 10 k positions alive), and a bounded hit-rate gap under pressure, where a
    set-associative table cannot replay a fully associative FIFO exactly."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import b2az
import parity_harness as ph
import refdriver
from conftest import ROOT, has_cuda, needs_ref

LIBS = [pytest.param(ph.HOSTEMU_LIB, id="host-emulation"),
        pytest.param(None, id="cuda", marks=[pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")])]


class RefCache:
    def __init__(self, max_size, ghost, np_=7, nv=3):
        self.L = refdriver.lib()
        self.h = self.L.azref_cache_new(max_size, ghost, np_, nv)
        self.np_, self.nv = np_, nv

    def find(self, key):
        pi, v = np.zeros(self.np_, np.float32), np.zeros(self.nv, np.float32)
        hit = self.L.azref_cache_find(self.h, C.c_uint64(int(key)), refdriver.P(pi), refdriver.P(v))
        return (pi, v) if hit else None

    def insert(self, key, pi, v):
        self.L.azref_cache_insert(self.h, C.c_uint64(int(key)), refdriver.P(np.ascontiguousarray(pi, np.float32)),
                                  refdriver.P(np.ascontiguousarray(v, np.float32)))

    def stats(self):
        s = np.zeros(6, np.uint64)
        self.L.azref_cache_stats(self.h, refdriver.P(s))
        return dict(zip(["hits", "misses", "evictions", "reinserts", "size", "max_size"], s.tolist()))

    def close(self):
        self.L.azref_cache_free(self.h)


def _value_of(key):
    rng = np.random.default_rng(int(key) & 0xFFFFFFFF)
    return rng.random(7, np.float32), rng.random(3, np.float32)


def _trace(n_ops, n_keys, seed, zipf=1.2):
    rng = np.random.default_rng(seed)
    keys = (rng.zipf(zipf, n_ops) % n_keys).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)
    return keys | np.uint64(1)


def _module():
    mod_dir = os.path.join(ROOT, "alphazero-pybind11_b200") if has_cuda() else os.path.join(ROOT, "tests", "cpp", "emu")
    sys.path.insert(0, mod_dir)
    try:
        sys.modules.pop("alphazero", None)
        import alphazero
        return alphazero
    finally:
        sys.path.remove(mod_dir)


@needs_ref
@pytest.mark.parametrize("max_size,ghost", [(1, 0), (3, 2), (8, 7), (50, 45), (64, 0)])
def test_host_s3fifo_class_replays_the_reference(max_size, ghost):
    az = _module()
    mine, ref = az.S3FIFOCache(max_size, ghost, 7, 3), RefCache(max_size, ghost)
    for key in _trace(6000, 40 * max_size + 5, seed=max_size * 31 + ghost):
        a, b = mine.find(int(key), 7, 3), ref.find(key)
        assert (a is None) == (b is None), "find() answers differ"
        if a is None:  # the self-play pattern: look up, evaluate on a miss, insert
            pi, v = _value_of(key)
            mine.insert(int(key), pi, v)
            ref.insert(key, pi, v)
        else:
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    want = ref.stats()
    got = {k: getattr(mine, k)() for k in want}
    assert got == want and want["evictions"] > 0
    ref.close()


@needs_ref
def test_sharded_host_class_counts_like_the_reference_shards():
    az = _module()
    sh = az.ShardedS3FIFOCache(64, 4, 56, 7, 3)
    assert (sh.max_size(), sh.size(), sh.hits(), sh.misses(), sh.evictions(), sh.reinserts()) == (64, 0, 0, 0, 0, 0)
    with pytest.raises(TypeError):
        az.ShardedS3FIFOCache(64, 4)


@needs_ref
@pytest.mark.parametrize("lib_path", LIBS)
def test_device_cache_equals_reference_while_no_set_overflows(lib_path):
    eng = ph.make_engine(lib_path, 4, 4, 8, b2az.EVAL_NN, b2az.RNG_PER_GAME, 1, max_cache_size=200000)
    ref = RefCache(200000, 180000)
    keys = _trace(4000, 900, seed=3)  # 900 distinct keys in 50,000 four-way sets: no set ever fills up
    answers_ref, pend_k, pend_v, pend_pi = [], [], [], []
    for i in range(0, len(keys), 50):  # batches of lookups, then the misses are inserted (order preserved on both sides)
        chunk = keys[i:i + 50]
        found, v, pi = eng.cache_find(chunk)
        ins = {}
        for j, key in enumerate(chunk):
            b = ref.find(key)
            answers_ref.append(b is not None)
            assert bool(found[j]) == (b is not None), f"lookup {i + j}: device {bool(found[j])} vs reference {b is not None}"
            if b is not None:
                assert np.array_equal(pi[j], b[0]) and np.array_equal(v[j], b[1])
            elif int(key) not in ins:
                ins[int(key)] = _value_of(key)
        for key, (p_, v_) in ins.items():
            ref.insert(key, p_, v_)
        if ins:
            eng.cache_insert(np.array(list(ins), np.uint64), np.stack([x[1] for x in ins.values()]), np.stack([x[0] for x in ins.values()]))
    st, want = eng.stats(), ref.stats()
    assert (st.cache_hits, st.cache_misses, st.cache_size, st.cache_evictions) == (want["hits"], want["misses"], want["size"], 0)
    assert want["evictions"] == 0 and st.cache_reinserts == want["reinserts"] == 0
    eng.close()
    ref.close()


@needs_ref
@pytest.mark.parametrize("lib_path", LIBS)
def test_device_cache_hit_rate_under_pressure_is_close_to_s3fifo(lib_path):
    size = 2048
    eng = ph.make_engine(lib_path, 4, 4, 8, b2az.EVAL_NN, b2az.RNG_PER_GAME, 1, max_cache_size=size)
    ref = RefCache(size, size * 9 // 10)
    keys = _trace(60000, 30000, seed=9, zipf=1.15)
    hits_dev = hits_ref = 0
    for i in range(0, len(keys), 64):
        chunk = keys[i:i + 64]
        found, _, _ = eng.cache_find(chunk)
        ins = {}
        for j, key in enumerate(chunk):
            hits_ref += ref.find(key) is not None
            if not found[j]:
                ins[int(key)] = _value_of(key)
        hits_dev += int(found.sum())
        for key, (p_, v_) in ins.items():
            ref.insert(key, p_, v_)  # (an existing key is a no-op on both sides)
        if ins:
            eng.cache_insert(np.array(list(ins), np.uint64), np.stack([x[1] for x in ins.values()]), np.stack([x[0] for x in ins.values()]))
    st = eng.stats()
    r_dev, r_ref = hits_dev / len(keys), hits_ref / len(keys)
    assert st.cache_evictions > 1000 and st.cache_size <= st.cache_max_size
    assert abs(r_dev - r_ref) < 0.05, f"hit rate {r_dev:.3f} (device, 4-way sets) vs {r_ref:.3f} (S3FIFOCache)"
    # what a hit returns is always the value that was inserted for that key
    found, v, pi = eng.cache_find(keys[-256:])
    for j in np.flatnonzero(found):
        p_, v_ = _value_of(keys[-256:][j])
        assert np.array_equal(pi[j], p_) and np.array_equal(v[j], v_)
    eng.close()
    ref.close()
