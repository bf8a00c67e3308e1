"""Per-generation timing of the step kernel at the bench configuration: brings the pool to steady state, then
launches `--span` single generations (b2az_step(1)) with a CUDA event after each, and prints the time of the
plain-simulation generations vs the generations in which games play a move (every game of this lock-step
RANDOM-eval workload reaches its 400-simulation budget in the same generation).
  python tools/gen_profile.py --lanes 8 --span 820
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
import torch  # noqa: E402

import b2az  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=65536)
ap.add_argument("--sims", type=int, default=400)
ap.add_argument("--preroll", type=int, default=24)
ap.add_argument("--span", type=int, default=820)
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--reuse", type=int, default=1)
a = ap.parse_args()
p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=a.games, mcts_visits=(a.sims, a.sims), cpuct=1.25,
                        fpu_reduction=0.25, eval_type=b2az.EVAL_RANDOM, rng_mode=b2az.RNG_PER_GAME, seed=1000,
                        tree_reuse=a.reuse, history_enabled=0, self_play=1, lanes_per_game=0)
e = b2az.Engine(p, lib=b2az.load(os.environ["B2AZ_LIB_PATH"]) if os.environ.get("B2AZ_LIB_PATH") else None)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(a.preroll):
    e.step(a.sims, stream)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.span + 1)]
ev[0].record()
for i in range(a.span):
    e.step(1, stream)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.span)]
order = sorted(range(a.span), key=lambda i: -ms[i])
spikes = sorted(order[: max(1, a.span // a.sims)])
plain = sorted(ms[i] for i in range(a.span) if all(abs(i - s) > 1 for s in spikes))
# fused launches (the bench shape): 3 launches of `sims` generations each
fe = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e.step(a.sims - (a.span % a.sims), stream)  # re-align to a move boundary
fe[0].record()
for i in range(3):
    e.step(a.sims, stream)
    fe[i + 1].record()
torch.cuda.synchronize()
fused = [fe[i].elapsed_time(fe[i + 1]) for i in range(3)]
st = e.stats()
out = {"fused_ms_per_%d_gens" % a.sims: [round(x, 3) for x in fused],
       "fused_Msims_per_s": round(a.games * a.sims / (sum(fused) / 3) / 1e3, 1),"lanes": a.lanes, "reuse": a.reuse, "span": a.span, "total_ms": sum(ms), "spike_gens": spikes,
       "spike_ms": [round(ms[i], 3) for i in spikes], "plain_median_ms": plain[len(plain) // 2],
       "plain_p10_ms": plain[len(plain) // 10], "plain_p90_ms": plain[9 * len(plain) // 10],
       "plain_by_phase_ms": [round(sum(ms[s + 2 + j * 50: s + 2 + (j + 1) * 50]) / 50, 4) for s in spikes[:1] for j in range(7)],
       "err": st.device_error, "depth": st.avg_leaf_depth}
print(json.dumps(out))
e.close()
