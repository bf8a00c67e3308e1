"""Generate tests/golden/brandubh_transcripts.npz from the UNMODIFIED reference BrandubhGS
(oracle/_ref/libazref_tafl.so, built from /root/reference by oracle/Makefile): random legal games from the
start position and, after every move, the reference's board, side to move, turn, repetition count, scores(),
number of legal moves, a CRC of the legal-move mask and of the canonical planes (full arrays for a few games).
Also a hand-made repetition transcript (both sides shuffle back and forth until the third repetition).
Run in the build container: `python tools/make_golden_tafl.py`."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tafl_ref  # noqa: E402

N_GAMES, FULL_GAMES, MAX_LEN = 48, 4, 160


def mv(h, w, row_slide, new):
    return (h * 7 + w) * 14 + (new if row_slide else 7 + new)


def shuffle_game():
    # attacker (0,3)->(0,2)->(0,3)..., defender (2,3)->(2,2)->(2,3)...: positions repeat, third repetition ends it
    a1, a2 = mv(0, 3, True, 2), mv(0, 2, True, 3)
    d1, d2 = mv(2, 3, True, 2), mv(2, 2, True, 3)
    return np.array([a1, d1, a2, d2, a1, d1, a2, d2], np.uint32)


if __name__ == "__main__":
    games = [tafl_ref.random_game(tafl_ref.BRANDUBH, 1000 + i, max_turns=150, max_len=MAX_LEN) for i in range(N_GAMES - 2)]
    games.append(tafl_ref.random_game(tafl_ref.BRANDUBH, 77, max_turns=12, max_len=MAX_LEN))  # ends by max_turns
    games.append(shuffle_game())
    max_turns = [150] * (N_GAMES - 2) + [12, 150]
    L = max(len(g) for g in games)
    moves = np.zeros((N_GAMES, L), np.uint16)
    lens = np.zeros(N_GAMES, np.uint32)
    R = (N_GAMES, L + 1)
    out = dict(boards=np.zeros(R + (3, 7, 7), np.int8), players=np.zeros(R, np.uint8), turns=np.zeros(R, np.uint32),
               reps=np.zeros(R, np.uint8), terminal=np.zeros(R, np.uint8), n_valid=np.zeros(R, np.uint32),
               valid_crc=np.zeros(R, np.uint32), canon_crc=np.zeros(R, np.uint32),
               valid_full=np.zeros((FULL_GAMES, L + 1, 686), np.uint8),
               canon_full=np.zeros((FULL_GAMES, L + 1, 7, 7, 7), np.float32))
    for i, g in enumerate(games):
        r = tafl_ref.replay(tafl_ref.BRANDUBH, g, max_turns=max_turns[i])
        n = len(g) + 1
        moves[i, :len(g)] = g
        lens[i] = len(g)
        for k in ("boards", "players", "turns", "reps", "terminal", "n_valid"):
            out[k][i, :n] = r[k]
        for k in range(n):
            out["valid_crc"][i, k] = zlib.crc32(r["valid"][k].tobytes())
            out["canon_crc"][i, k] = zlib.crc32(r["canonical"][k].tobytes())
        if i < FULL_GAMES:
            out["valid_full"][i, :n] = r["valid"]
            out["canon_full"][i, :n] = r["canonical"]
    path = os.path.join(ROOT, "tests", "golden", "brandubh_transcripts.npz")
    np.savez_compressed(path, moves=moves, lens=lens, max_turns=np.array(max_turns, np.uint32), **out)
    ends = [int(out["terminal"][i, lens[i]]) for i in range(N_GAMES)]
    print("games", N_GAMES, "mean length", lens.mean(), "max", lens.max(), "terminal codes", np.bincount(ends, minlength=4),
          "max rep", out["reps"].max(), os.path.getsize(path), "bytes")
