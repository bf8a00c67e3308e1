"""Golden fixtures for tests/test_tafl_selfplay.py: the UNMODIFIED reference PlayManager (oracle/_ref/libazref_tafl.so,
built from /root/reference by oracle/Makefile) plays every case of test_tafl_selfplay.CASES slot by slot; per slot the
fixture keeps (samples, crc32 canonical, crc32 outcomes, crc32 policy targets, game length sum, crc32 scores).
Run in the build container (needs /root/reference): python tools/make_golden_tafl_selfplay.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
import tafl_ref  # noqa: E402
import test_tafl_selfplay as t  # noqa: E402

out = {}
for i, name in enumerate(sorted(t.CASES)):
    game, slots, per_slot, max_turns, visits, kw = t.CASES[name]
    seed = 1000 + 17 * i
    rows = []
    for g in range(slots):
        r = tafl_ref.selfplay(game, seed + g, max_turns, per_slot, visits, **kw)
        length = int(round(float(r["avg_game_length"]) * r["games_completed"]))
        rows.append((len(r["v"]), t.crc(r["canonical"]), t.crc(r["v"]), t.crc(r["pi"]), length, t.crc(r["scores"])))
    out[name] = np.array(rows, np.int64)
    print(name, out[name][:, 0], out[name][:, 4])
np.savez(os.path.join(ROOT, "tests", "golden", "tafl_selfplay.npz"), **out)
