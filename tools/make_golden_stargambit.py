"""Golden Star Gambit transcripts from the UNMODIFIED reference (star_gambit_gs.cc in oracle/_ref/libazref_tafl.so):
for the four variants' own classes and the four pinned Unified views, random legal games (azref_tafl_random_game's
splitmix64 move choice) with, after every move, player / turn / terminal code / number of legal moves and CRC-32s of
the legal-move mask, the canonical planes (float32 bytes) and the serialised unit list — tests/golden/
stargambit_transcripts.npz. Run here (needs /root/reference built into oracle/_ref); the fixture travels to the GPU box."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tafl_ref  # noqa: E402

GAMES = [10, 11, 12, 13, 20, 21, 22, 23]
MAX_LEN = 480


def units_blob(row, n):
    """[num_units u32 | units 9 B each | reserves 8 | player | turn u32 | acted | over | winner] of the reference bytes"""
    b = row[:n].tobytes()
    nu = int.from_bytes(b[:4], "little")
    return b[: 4 + 9 * nu + 8 + 1 + 4 + 3]


def main():
    out = {}
    for g in GAMES:
        trans = [tafl_ref.sg_random_game(g, 424242 + 31 * g + i, max_len=MAX_LEN if i else 4096)[:MAX_LEN] for i in range(3)]
        # one transcript that runs into the end of a game: take the tail of a full game is not replayable; instead search seeds
        for seed in range(400):
            t = tafl_ref.sg_random_game(g, 9000 + seed, max_len=MAX_LEN)
            if len(t) < MAX_LEN:
                trans.append(t)
                break
        n = len(trans)
        moves = np.zeros((n, MAX_LEN), np.uint16)
        lens = np.zeros(n, np.uint32)
        meta = np.zeros((n, MAX_LEN + 1, 4), np.uint32)   # player, turn, terminal, n_valid
        crcs = np.zeros((n, MAX_LEN + 1, 3), np.uint32)   # valid, canonical, units
        for i, t in enumerate(trans):
            moves[i, : len(t)] = t
            lens[i] = len(t)
            ref = tafl_ref.sg_replay(g, t)
            for k in range(len(t) + 1):
                meta[i, k] = (ref["players"][k], ref["turns"][k], ref["terminal"][k], ref["n_valid"][k])
                crcs[i, k] = (zlib.crc32(ref["valid"][k].tobytes()), zlib.crc32(ref["canonical"][k].tobytes()),
                              zlib.crc32(units_blob(ref["bytes"][k], int(ref["bytes_len"][k]))))
        out[f"g{g}_moves"] = moves
        out[f"g{g}_lens"] = lens
        out[f"g{g}_meta"] = meta
        out[f"g{g}_crcs"] = crcs
        print(g, "lens", lens.tolist(), "ended", [int(meta[i, lens[i], 2]) for i in range(n)])
    path = os.path.join(ROOT, "tests", "golden", "stargambit_transcripts.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
