"""Generate tests/golden/c4_*.npz from the UNMODIFIED reference (oracle/_ref/libazref.so, built from
/root/reference by oracle/Makefile). Run in the build container: `python tools/make_golden.py`.

Each fixture is a full deterministic self-play trace (the reference ships no golden
transcripts, so they are generated from it): single worker thread, MCTS::seed_thread_rng(seed), lock-step
NN evaluation with tests/parity_harness.fake_net or EvalType::RANDOM. Recorded: sha256 over every
generation's leaf batch (ids + canonical planes), peeked visit counts / Q / root values, every finished
training sample (root canonical, final score, policy target), final scores and metrics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_harness as ph  # noqa: E402

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (G, games, visits, level, seed, et) in ph.GOLDEN_CASES.items():
        pm = ph.RefPM(G=G, games_to_play=games, visits=visits, eval_type=et, rng_mode=1, seed=seed, **ph.level_params(level))
        tr = ph.trace_run(pm, et)
        pm.close()
        np.savez_compressed(ph.golden_path(name), **tr)
        print(name, "generations", int(tr["generations"]), "samples", len(tr["hist_pi"]), "scores", tr["scores"],
              os.path.getsize(ph.golden_path(name)), "bytes")
