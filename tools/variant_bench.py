"""Throughput of experiment builds of libb2az.so on the bench workload (65,536 Connect4 games, 400 sims/move,
RANDOM eval): one subprocess per library, pre-roll to steady state, then `--launches` fused launches of 400
generations timed with CUDA events. Usage (GPU box):
  python tools/variant_bench.py build/variants/*.so            # the in-tree library is always measured first
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import os, sys, json
sys.path.insert(0, os.path.join(%(root)r, "alphazero-pybind11_b200"))
import torch, b2az
lib = b2az.load(os.environ["B2AZ_LIB_PATH"]) if os.environ.get("B2AZ_LIB_PATH") else None
G, S, L = 65536, 400, %(launches)d
p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=G, mcts_visits=(S, S), cpuct=1.25,
                        fpu_reduction=0.25, eval_type=b2az.EVAL_RANDOM, rng_mode=b2az.RNG_PER_GAME, seed=1000,
                        tree_reuse=1, history_enabled=0, self_play=1, lanes_per_game=0)
e = b2az.Engine(p, lib=lib)
st = torch.cuda.current_stream().cuda_stream
for _ in range(16):
    e.step(S, st)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
ev[0].record()
for i in range(L):
    e.step(S, st)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
s = e.stats()
print(json.dumps({"lib": os.environ.get("B2AZ_LIB_PATH", "in-tree"), "Msims_per_s": round(G * S / (sum(ms) / L) / 1e3, 1),
                  "ms": [round(x, 2) for x in ms], "err": s.device_error, "depth": round(s.avg_leaf_depth, 4),
                  "sims": s.simulations, "moves": s.moves}))
e.close()
"""


def main():
    libs = [None] + sys.argv[1:]
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["B2AZ_LIB_PATH"] = os.path.abspath(lib)
        else:
            env.pop("B2AZ_LIB_PATH", None)
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "launches": 6}], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=600)
        out = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(out[-1] if out else json.dumps({"lib": lib, "failed": r.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
