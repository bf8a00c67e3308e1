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
import ctypes
prof = None
try:
    qp = e.L.b2az_debug_qprof  # only in -DB2AZ_Q_PROF experiment builds
    prof = (ctypes.c_ulonglong * 16)()
    qp(prof)  # clear
except AttributeError:
    pass
wprof = None
try:
    wp = e.L.b2az_debug_wprof  # only in -DB2AZ_W_PROF experiment builds
    wprof = (ctypes.c_ulonglong * 16)()
    wp(wprof)
except AttributeError:
    pass
ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
ev[0].record()
for i in range(L):
    e.step(S, st)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
s = e.stats()
extra = {}
if prof is not None:
    qp(prof)
    v = list(prof)
    tot = sum(v[0:5]) or 1
    extra = {"qprof": {"pct_level": round(100 * v[0] / tot, 1), "pct_leaf": round(100 * v[1] / tot, 1),
                       "pct_move": round(100 * v[2] / tot, 1), "pct_idle": round(100 * v[3] / tot, 1),
                       "pct_queue": round(100 * v[4] / tot, 1),
                       "cycles_per_batch": [round(v[i] / max(1, v[5 + i])) for i in range(3)],
                       "games_per_batch": [round(v[8 + i] / max(1, v[5 + i]), 1) for i in range(3)],
                       "batches": v[5:8], "failed_pops": v[11],
                       "per_batch_pop_decide": round(v[12] / max(1, sum(v[5:8]))),
                       "per_batch_pop_slots_fence": round(v[13] / max(1, sum(v[5:8]))),
                       "per_batch_push_fence": round(v[14] / max(1, sum(v[5:8]))),
                       "per_batch_push_fence_and_rings": round(v[15] / max(1, sum(v[5:8])))}}
if wprof is not None:
    wp(wprof)
    v = list(wprof)
    tot = sum(v[0:4]) or 1
    extra["wprof"] = {"pct_level": round(100 * v[0] / tot, 1), "pct_leaf": round(100 * v[1] / tot, 1),
                      "pct_move": round(100 * v[2] / tot, 1), "pct_barrier_wait": round(100 * v[3] / tot, 1),
                      "rounds": v[4], "cycles_per_chunk": [round(v[i] / max(1, v[5 + i])) for i in range(3)],
                      "games_per_chunk": [round(v[8 + i] / max(1, v[5 + i]), 1) for i in range(3)], "chunks": v[5:8]}
print(json.dumps({**extra, "lib": os.environ.get("B2AZ_LIB_PATH", "in-tree"), "Msims_per_s": round(G * S / (sum(ms) / L) / 1e3, 1),
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
