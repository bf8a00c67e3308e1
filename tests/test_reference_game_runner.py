"""The UNMODIFIED reference Python (/root/reference/src/game_runner.py: GameRunner thread pipeline + self_play) driving this
repo's `alphazero` module — the drop-in claim of SURVEY.md 8b, tested the way a user would meet it.

game_runner.self_play builds PlayParams exactly as the reference does (base_params + the self-play flags,
set_model_groups -> model_groups = [0, 0], set_eval_types), constructs alphazero.PlayManager, starts its mcts_workers /
hist_saver / monitor threads and writes the training samples as .ptz files. The test then replays the same parameters
straight through the C ABI and requires the samples on disk to be the engine's samples (fp16 storage), and the returned
SelfPlayResult to carry the engine's statistics.

The module under test is the host-emulation build on a CPU-only box (tests/cpp/emu), the CUDA build on a GPU box (where
the reference's Python comes from baseline/_ref/src, staged unmodified by __graft_entry__.build())."""
import glob
import importlib
import os
import sys
import tempfile

import numpy as np
import pytest

import b2az
import parity_harness as ph
from conftest import ROOT, has_cuda

# the reference checkout where it exists (this container), else the UNMODIFIED copy that __graft_entry__.build() staged under
# baseline/_ref/src (git-ignored, shipped to the GPU box)
REF_SRC = "/root/reference/src" if os.path.isdir("/root/reference/src") else os.path.join(ROOT, "baseline", "_ref", "src")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_SRC, "game_runner.py")),
                                reason="needs the reference's Python (/root/reference or baseline/_ref/src)")


KIND = {"emu": None}


@pytest.fixture(scope="module", params=[
    pytest.param("emu", id="host-emulation"),
    pytest.param("cuda", id="cuda", marks=[pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")])])
def gr(request):
    """import the reference's game_runner with `alphazero` = this repo's module and a zstandard stand-in"""
    cuda = request.param == "cuda"
    KIND["cuda"] = cuda
    mod_dir = os.path.join(ROOT, "alphazero-pybind11_b200") if cuda else os.path.join(ROOT, "tests", "cpp", "emu")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for name in ("alphazero", "game_runner", "config", "neural_net", "tracy_utils", "frozen_eval", "zstandard"):
        sys.modules.pop(name, None)
    sys.path[:0] = [mod_dir, os.path.join(ROOT, "tests", "stubs"), REF_SRC]
    try:
        game_runner = importlib.import_module("game_runner")
        assert os.path.dirname(game_runner.__file__) == REF_SRC
        assert os.path.dirname(sys.modules["alphazero"].__file__) == mod_dir
        yield game_runner
    finally:
        sys.path[:] = saved_path
        for name in ("alphazero", "game_runner", "config", "neural_net", "tracy_utils", "frozen_eval", "zstandard"):
            sys.modules.pop(name, None)
        sys.modules.update({k: v for k, v in saved_mods.items() if k in ("alphazero", "zstandard")})


def _load_samples(game_runner, folder):
    out = {}
    for kind in ("canonical", "v", "pi"):
        files = sorted(glob.glob(os.path.join(folder, f"*-{kind}-*.ptz")))
        assert files, f"no {kind} files written"
        out[kind] = np.concatenate([game_runner.load_compressed(f).float().numpy() for f in files])
    return out["canonical"], out["v"], out["pi"]


def _rows16(canon, v, pi):
    """the multiset of samples as they are stored on disk (fp16)"""
    rows = np.concatenate([np.asarray(canon, np.float16).reshape(len(canon), -1), np.asarray(v, np.float16),
                           np.asarray(pi, np.float16)], axis=1).view(np.uint16)
    return rows[np.lexsort(rows.T[::-1])]


def test_self_play_connect4_unmodified_game_runner(gr):
    import config as ref_config

    cfg = ref_config.TrainConfig(game="connect4", self_play_batch_size=3, self_play_concurrent_batch_mult=1, self_play_chunks=2,
                                 mcts_workers=3, playout_cap_percent=0.75, resign_percent=0.0)
    depth, fast_depth = 40, 10
    with tempfile.TemporaryDirectory() as tmp:
        paths = {"tmp_history": os.path.join(tmp, "hist"), "checkpoint": os.path.join(tmp, "ckpt")}
        res = gr.self_play(cfg, paths, "exp", best=0, iteration=7, depth=depth, fast_depth=fast_depth)
        canon, v, pi = _load_samples(gr, paths["tmp_history"])
        assert all(os.path.basename(f).startswith("0007-") for f in glob.glob(os.path.join(paths["tmp_history"], "*.ptz")))
    G, n = 3 * 2 * 1, 3 * 2 * 1 * 2
    assert abs(sum(res.win_rates) - 1.0) < 1e-6 and res.game_length > 6 and res.avg_depth > 0 and res.fast_avg_depth > 0
    assert res.hit_rate == 0 and res.variant_game_counts == {}
    # the same run straight through the C ABI (what self_play() asked the module for: game_runner.py:790-816, 2022-2041)
    lib = None if KIND["cuda"] else ph.HOSTEMU_LIB
    eng = ph.make_engine(lib, G, n, depth, b2az.EVAL_RANDOM, b2az.RNG_PER_GAME, 0, cpuct=cfg.cpuct, start_temp=cfg.self_play_temp,
                         final_temp=cfg.final_temp, temp_decay_half_life=float(cfg.temp_decay_half_life),
                         fpu_reduction=cfg.fpu_reduction, epsilon=0.25, playout_cap_randomization=1, playout_cap_depth=fast_depth,
                         playout_cap_percent=0.75, mcts_root_temp=cfg.mcts_root_temp, root_fpu_zero=int(cfg.root_fpu_zero),
                         shaped_dirichlet=int(cfg.shaped_dirichlet), policy_target_pruning=int(cfg.policy_target_pruning),
                         history_capacity=n * 42)
    chunk = max(1, min(depth, 512))  # the module's RANDOM-eval driver fuses one search's worth of generations per launch
    while eng.stats().active_games:
        eng.step(chunk)
    st = eng.stats()
    he = eng.drain_history(n * 42)
    eng.close()
    assert len(canon) == len(he[0]) > 30
    assert np.array_equal(_rows16(canon, v, pi), _rows16(*he)), "samples written by game_runner != the engine's samples"
    assert res.game_length == pytest.approx(st.avg_game_length) and res.avg_depth == pytest.approx(st.avg_leaf_depth)
    wins = np.array(st.scores[:], np.float64)
    assert np.allclose(res.win_rates, wins / wins.sum())


def test_game_runner_nn_pipeline_two_threads_connect4(gr):
    """GameRunner's batcher / gpu_loop / result_worker threads (game_runner.py:651-727) with an NN-shaped player object:
    build_batch -> process() -> update_inferences against the module, model_groups = [0, 0] as set_model_groups builds it."""
    import torch

    az = sys.modules["alphazero"]

    class FakeNet:  # the I/O contract of NNWrapper.process (neural_net.py:801-823): canonical batch -> (v, pi) probabilities
        def warmup_graphs(self, max_batch):
            pass

        def process(self, batch):
            v, pi = ph.fake_net(batch.cpu().numpy())
            return torch.from_numpy(v), torch.from_numpy(pi)

    net = FakeNet()
    p = az.PlayParams()
    p.games_to_play, p.concurrent_games, p.max_batch_size = 10, 5, 5
    p.mcts_visits = [24, 24]
    p.history_enabled = p.self_play = True
    p.cpuct, p.fpu_reduction, p.epsilon, p.mcts_root_temp = 1.25, 0.25, 0.25, 1.25
    players = [net, net]
    gr.set_model_groups(p, players)
    gr.set_eval_types(p, players)
    assert list(p.model_groups) == [0, 0]
    pm = az.PlayManager(az.Connect4GS(), p)
    with tempfile.TemporaryDirectory() as tmp:
        args = gr.GRArgs(title="t", game=az.Connect4GS, iteration=0, max_batch_size=5, mcts_workers=2, data_folder=tmp)
        gr.GameRunner(players, pm, args).run()
        canon, v, pi = _load_samples(gr, tmp)
    assert pm.games_completed() == 10 and pm.remaining_games() == 0
    assert len(canon) == int(round(pm.avg_game_length() * 10)) and np.allclose(pi.sum(1), 1.0, atol=2e-3)


def test_self_play_brandubh_unmodified_game_runner(gr):
    if not KIND["cuda"]:
        pytest.skip("the tafl self-play engine has no host-emulation build (device only)")
    import config as ref_config

    cfg = ref_config.TrainConfig(game="brandubh", self_play_batch_size=4, self_play_concurrent_batch_mult=1, self_play_chunks=1,
                                 mcts_workers=2, playout_cap_percent=0.0, resign_percent=0.0, max_turns=40)
    with tempfile.TemporaryDirectory() as tmp:
        paths = {"tmp_history": os.path.join(tmp, "hist"), "checkpoint": os.path.join(tmp, "ckpt")}
        res = gr.self_play(cfg, paths, "exp", best=0, iteration=1, depth=32, fast_depth=8)
        canon, v, pi = _load_samples(gr, paths["tmp_history"])
    assert canon.shape[1:] == (7, 7, 7) and pi.shape[1] == 686 and len(canon) > 30
    assert abs(sum(res.win_rates) - 1.0) < 1e-6 and res.game_length > 4


def test_self_play_with_the_reference_brandubh_yaml(gr):
    """configs/brandubh.yaml as shipped by the reference — Gumbel root search, Gumbel fast searches, playout-cap
    randomisation (75 % fast searches of 30 simulations), resign_percent 0.02 with playthrough, temperature decay, a
    200 k-entry cache request — through the unmodified self_play() on the tafl self-play engine. Only the sizes are
    shrunk (batch size, chunks, max_turns)."""
    if not KIND["cuda"]:
        pytest.skip("the tafl self-play engine has no host-emulation build (device only)")
    import config as ref_config

    yaml_path = os.path.join(os.path.dirname(REF_SRC), "configs", "brandubh.yaml")
    cfg = ref_config.load_config(yaml_path, {"self_play_batch_size": "8", "self_play_concurrent_batch_mult": "1",
                                             "self_play_chunks": "1", "mcts_workers": "2", "max_turns": "60"}, warn=False)
    assert cfg.gumbel_enabled and cfg.playout_cap_percent > 0 and cfg.resign_percent > 0 and cfg.max_cache_size > 0
    depth, fast_depth = cfg.selfplay_mcts_visits, cfg.fast_mcts_visits
    with tempfile.TemporaryDirectory() as tmp:
        paths = {"tmp_history": os.path.join(tmp, "hist"), "checkpoint": os.path.join(tmp, "ckpt")}
        res = gr.self_play(cfg, paths, "exp", best=0, iteration=3, depth=depth, fast_depth=fast_depth)
        canon, v, pi = _load_samples(gr, paths["tmp_history"])
    assert canon.shape[1:] == (7, 7, 7) and pi.shape[1] == 686
    assert abs(sum(res.win_rates) - 1.0) < 1e-6 and res.game_length > 4
    assert res.avg_depth > 0 and res.fast_avg_depth > 0, "both full and fast searches happened"
    # samples come from full searches only: about a quarter of the moves of the 16 games
    moves = res.game_length * 16
    assert 0.08 * moves < len(canon) < 0.5 * moves
    assert np.allclose(pi.sum(1), 1.0, atol=4e-3) and np.allclose(v.sum(1), 1.0)
