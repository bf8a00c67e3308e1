"""Generate tests/golden/forest_random_eval.npz from the UNMODIFIED reference (oracle/_ref/libazref_tafl.so): single-tree
MCTS runs (MCTS class over BrandubhGS / OpenTaflGS / TawlbwrddGS, dumb_eval, MCTS::seed_thread_rng(seed + i)); per
move the played move and CRC32s of counts() and root_q_values(). Run in the build container."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
import tafl_ref  # noqa: E402
from test_forest import GOLDEN, GOLDEN_CASES, MAX_TURNS  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name, (game, trees, n_moves, sims, seed, cpuct, fpu, rfz) in GOLDEN_CASES.items():
        moves = np.zeros((trees, n_moves), np.uint32)
        lens = np.zeros(trees, np.uint32)
        ccrc = np.zeros((trees, n_moves), np.uint32)
        qcrc = np.zeros((trees, n_moves), np.uint32)
        for i in range(trees):
            c, q, mv, _ = tafl_ref.search(game, seed + i, n_moves, sims, MAX_TURNS[game], cpuct, fpu, rfz, None)
            lens[i] = len(mv)
            moves[i, :len(mv)] = mv
            for m in range(len(mv)):
                ccrc[i, m] = zlib.crc32(c[m].tobytes())
                qcrc[i, m] = zlib.crc32(q[m].tobytes())
        out.update({f"{name}_moves": moves, f"{name}_lens": lens, f"{name}_counts_crc": ccrc, f"{name}_q_crc": qcrc})
        print(name, "moves searched per tree", lens.tolist())
    np.savez_compressed(GOLDEN, **out)
    print(os.path.getsize(GOLDEN), "bytes")
