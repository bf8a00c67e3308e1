"""Generate tests/golden/tafl_<game>_transcripts.npz from the UNMODIFIED reference BrandubhGS / OpenTaflGS /
TawlbwrddGS (oracle/_ref/libazref_tafl.so, built from /root/reference by oracle/Makefile): random legal games
from the start position and, after every move, the reference's board, side to move, turn, repetition count,
scores(), number of legal moves, a CRC of the legal-move mask and of the canonical planes (full arrays for two
games). Plus a short-max_turns game (draw by turn limit) and a hand-made shuffle game (third repetition).
Run in the build container: `python tools/make_golden_tafl.py`."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tafl_ref  # noqa: E402

NAMES = {0: "brandubh", 1: "opentafl", 2: "tawlbwrdd"}
N_GAMES = {0: 48, 1: 20, 2: 20}
MAX_TURNS = {0: 150, 1: 400, 2: 400}
FULL_GAMES = 2


def shuffle_game(game):
    """Both sides move a piece out and back until the start position (attackers to move) occurs a third time."""
    S = tafl_ref.dims(game)[0]

    def mv(h, w, new_w):
        return (h * S + w) * 2 * S + new_w

    if game == 0:
        a, d = ((0, 3), 2), ((2, 3), 2)
    elif game == 1:
        a, d = ((0, 3), 2), ((3, 5), 4)
    else:
        a, d = ((0, 4), 3), ((2, 5), 4)
    (ah, aw), at = a
    (dh, dw), dt = d
    cyc = [mv(ah, aw, at), mv(dh, dw, dt), mv(ah, at, aw), mv(dh, dt, dw)]
    return np.array(cyc * 2, np.uint32)


def path(game):
    return os.path.join(ROOT, "tests", "golden", f"tafl_{NAMES[game]}_transcripts.npz")


if __name__ == "__main__":
    for game in (0, 1, 2):
        S, A, P = tafl_ref.dims(game)
        n = N_GAMES[game]
        games = [tafl_ref.random_game(game, 1000 + i, max_turns=MAX_TURNS[game], max_len=MAX_TURNS[game] + 8) for i in range(n - 2)]
        games.append(tafl_ref.random_game(game, 77, max_turns=12, max_len=64))  # ends by max_turns at the latest
        games.append(shuffle_game(game))
        max_turns = [MAX_TURNS[game]] * (n - 2) + [12, MAX_TURNS[game]]
        L = max(len(g) for g in games)
        moves = np.zeros((n, L), np.uint16)
        lens = np.zeros(n, np.uint32)
        R = (n, L + 1)
        out = dict(boards=np.zeros(R + (3, S, S), np.int8), players=np.zeros(R, np.uint8), turns=np.zeros(R, np.uint32),
                   reps=np.zeros(R, np.uint8), terminal=np.zeros(R, np.uint8), n_valid=np.zeros(R, np.uint32),
                   valid_crc=np.zeros(R, np.uint32), canon_crc=np.zeros(R, np.uint32),
                   valid_full=np.zeros((FULL_GAMES, L + 1, A), np.uint8),
                   canon_full=np.zeros((FULL_GAMES, L + 1, P, S, S), np.float32))
        for i, g in enumerate(games):
            r = tafl_ref.replay(game, g, max_turns=max_turns[i])
            m = len(g) + 1
            moves[i, :len(g)] = g
            lens[i] = len(g)
            for k in ("boards", "players", "turns", "reps", "terminal", "n_valid"):
                out[k][i, :m] = r[k]
            for k in range(m):
                out["valid_crc"][i, k] = zlib.crc32(r["valid"][k].tobytes())
                out["canon_crc"][i, k] = zlib.crc32(r["canonical"][k].tobytes())
            if i < FULL_GAMES:
                out["valid_full"][i, :m] = r["valid"]
                out["canon_full"][i, :m] = r["canonical"]
        np.savez_compressed(path(game), moves=moves, lens=lens, max_turns=np.array(max_turns, np.uint32), **out)
        ends = [int(out["terminal"][i, lens[i]]) for i in range(n)]
        print(NAMES[game], "games", n, "mean length", round(float(lens.mean()), 1), "max", lens.max(), "terminal codes",
              np.bincount(ends, minlength=4), "max rep", out["reps"].max(), "mean legal moves",
              round(float(out["n_valid"].sum() / (lens + 1).sum()), 1), os.path.getsize(path(game)), "bytes")
