"""Workload for ncu: the bench configuration (65,536 Connect4 games, 400 sims/move, RANDOM eval), brought to
steady state, then a few short step launches to capture. Usage (GPU box):
  ncu --set full --clock-control none --import-source on -k regex:k_step -s <preroll+2> -c 2 -o gpurun_out/prof \
      python tools/profile_step.py --preroll 24 --gens 50 --launches 4
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
import b2az  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=65536)
ap.add_argument("--preroll", type=int, default=24)
ap.add_argument("--gens", type=int, default=50)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--lanes", type=int, default=0)
a = ap.parse_args()
p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=a.games, mcts_visits=(400, 400), cpuct=1.25,
                        fpu_reduction=0.25, eval_type=b2az.EVAL_RANDOM, rng_mode=b2az.RNG_PER_GAME, seed=1000,
                        tree_reuse=1, history_enabled=0, self_play=1, lanes_per_game=0)
e = b2az.Engine(p, lib=b2az.load(os.environ["B2AZ_LIB_PATH"]) if os.environ.get("B2AZ_LIB_PATH") else None)
for _ in range(a.preroll):
    e.step(400)
for _ in range(a.launches):
    e.step(a.gens)
st = e.stats()
print("sims", st.simulations, "moves", st.moves, "err", st.device_error, "depth", st.avg_leaf_depth)
e.close()
