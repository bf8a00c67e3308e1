"""Golden self-play traces generated from the UNMODIFIED reference (tools/make_golden.py) vs the oracle
port and vs the host build of the engine logic. These run anywhere (no /root/reference needed)."""
import numpy as np
import pytest

import b2az
import parity_harness as ph


def _load(name):
    return dict(np.load(ph.golden_path(name)))


@pytest.mark.parametrize("name", sorted(ph.GOLDEN_CASES))
def test_port_reproduces_golden(name):
    G, games, visits, level, seed, et = ph.GOLDEN_CASES[name]
    pm = ph.PortPM(G=G, games_to_play=games, visits=visits, eval_type=et, rng_mode=1, seed=seed, **ph.level_params(level))
    got = ph.trace_run(pm, et)
    pm.close()
    ph.compare_trace(got, _load(name), f"port vs golden {name}")


@pytest.mark.parametrize("name", sorted(ph.GOLDEN_CASES))
def test_engine_logic_reproduces_golden(name):
    G, games, visits, level, seed, et = ph.GOLDEN_CASES[name]
    pm = ph.EnginePM(ph.HOSTEMU_LIB, G=G, games_to_play=games, visits=visits, eval_type=et, rng_mode=b2az.RNG_GLOBAL,
                     seed=seed, **ph.level_params(level))
    got = ph.trace_run(pm, et)
    pm.close()
    ph.compare_trace(got, _load(name), f"engine (host build) vs golden {name}")


def test_fake_net_is_batch_independent():
    rng = np.random.default_rng(0)
    x = (rng.random((64, 4, 6, 7)) < 0.3).astype(np.float32)
    v, pi = ph.fake_net(x)
    for i in (0, 17, 63):
        v1, p1 = ph.fake_net(x[i:i + 1])
        assert np.array_equal(v1[0], v[i]) and np.array_equal(p1[0], pi[i])
    assert np.allclose(v.sum(1), 1, atol=1e-6) and np.allclose(pi.sum(1), 1, atol=1e-6)
