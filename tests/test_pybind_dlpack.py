"""The zero-copy evaluator feed of the drop-in module (py_alphazero.cc leaf_batch_dlpack / update_inferences_dlpack;
SURVEY.md 8b "additive exports"): the leaf batch and its legal-move masks reach torch as DLPack capsules over the
engine's own device buffers, the evaluations go back as CUDA tensors — no host copy in either direction. With an
evaluator that answers dumb_eval's numbers (computed on the device from the mask capsule) the run equals the
EvalType.RANDOM run of the same seed, sample for sample."""
import numpy as np
import pytest

from conftest import has_cuda
from test_pybind_module import module

gpu = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def c4_params(az, G, games, visits, seed, random_eval, **kw):
    p = az.PlayParams()
    p.games_to_play, p.concurrent_games, p.max_batch_size = games, G, G
    p.mcts_visits = [visits, visits]
    p.model_groups = [0, 0]
    p.history_enabled = p.self_play = p.tree_reuse = True
    p.cpuct, p.fpu_reduction = 1.25, 0.25
    p.seed = seed
    for k, v in kw.items():
        setattr(p, k, v)
    if random_eval:
        p.eval_type = [az.EvalType.RANDOM, az.EvalType.RANDOM]
    return p


def history_rows(pm, cap, shape, A):
    canon, v, pi = np.zeros((cap,) + shape, np.float32), np.zeros((cap, 3), np.float32), np.zeros((cap, A), np.float32)
    n = pm.build_history_batch(canon, v, pi)
    return sorted(canon[i].tobytes() + v[i].tobytes() + pi[i].tobytes() for i in range(n))


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
def test_connect4_dlpack_feed_equals_the_random_eval_run(kind):
    import torch

    az = module(kind)
    G, games, visits, seed = 96, 192, 40, 777
    kw = dict(epsilon=0.25, mcts_root_temp=1.25, policy_target_pruning=True, start_temp=1.0, final_temp=0.2, temp_decay_half_life=10.0)
    pm0 = az.PlayManager(az.Connect4GS(), c4_params(az, G, games, visits, seed, True, **kw))
    pm0.play()
    want = history_rows(pm0, games * 42, (4, 6, 7), 7)
    pm = az.PlayManager(az.Connect4GS(), c4_params(az, G, games, visits, seed, False, **kw))
    batches = rows = 0
    while pm.remaining_games() > 0:
        canon, valid, n = pm.leaf_batch_dlpack(0)
        if n == 0:
            break
        x, m = torch.from_dlpack(canon), torch.from_dlpack(valid)
        assert x.is_cuda and m.is_cuda and tuple(x.shape) == (n, 4, 6, 7) and tuple(m.shape) == (n, 7)
        assert x.dtype == torch.float32 and m.dtype == torch.uint8
        mf = m.to(torch.float32)
        pi = mf / mf.sum(1, keepdim=True)  # dumb_eval (game_state.h:160-173), on the device
        v = torch.full((n, 3), 1.0 / 3.0, dtype=torch.float32, device=x.device)
        pm.update_inferences_dlpack(0, v, pi)
        batches += 1
        rows += n
    assert pm.games_completed() == games and batches > 40 and rows > games
    assert history_rows(pm, games * 42, (4, 6, 7), 7) == want
    assert np.array_equal(pm.scores(), pm0.scores())
    with pytest.raises(RuntimeError, match="no leaf batch"):
        pm.update_inferences_dlpack(0, torch.zeros(1, 3, device="cuda"), torch.zeros(1, 7, device="cuda"))


@pytest.mark.parametrize("kind", [pytest.param("cuda", marks=gpu)])
def test_wide_tree_dlpack_feed_runs_a_brandubh_and_a_star_gambit_game(kind):
    import torch

    az = module(kind)
    for gs, shape in ((az.BrandubhGS(24), (7, 7, 7)), (az.StarGambitUnifiedGS(0), (36, 13, 13))):
        G, visits = 6, 16
        A = gs.num_moves()
        p = c4_params(az, G, G, visits, 5, False)
        pm = az.PlayManager(gs, p)
        steps = 0
        while pm.remaining_games() > 0 and steps < 40000:
            canon, valid, n = pm.leaf_batch_dlpack(0)
            x = torch.from_dlpack(canon)
            assert valid is None and n == G and tuple(x.shape) == (G,) + shape and x.is_cuda
            s = x.reshape(G, -1).sum(1, keepdim=True)  # any device-side function of the planes will do as a "network"
            v = torch.softmax(torch.cat([s.sin(), s.cos(), s * 0], 1), 1).contiguous()
            pi = torch.full((G, A), 1.0 / A, dtype=torch.float32, device=x.device)
            pm.update_inferences_dlpack(0, v, pi)
            steps += 1
        assert pm.games_completed() == G
        rows = history_rows(pm, G * 600, shape, A)
        assert len(rows) > G


def test_dlpack_feed_argument_errors():
    az = module("emu")
    p = c4_params(az, 4, 4, 8, 1, True)
    if has_cuda():
        pm = az.PlayManager(az.Connect4GS(), p)
        with pytest.raises(RuntimeError, match="EvalType.NN"):
            pm.leaf_batch_dlpack(0)
