"""The N>1 path on CPU: two processes (gloo), each with its own pool of games on the host-emulation build of the
engine, sharded / reduced / gathered with b2az.dist exactly the way bench.py and a multi-GPU self-play run do it
with NCCL. Checks: the shards cover the budget, the reduced statistics equal a single-process run of the same
games, and rank 0 receives every training sample."""
import os
import subprocess
import sys

import numpy as np
import pytest

import b2az
import parity_harness as ph
from b2az import dist as bd

WORKER = r"""
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["B2AZ_PKG"]); sys.path.insert(0, os.environ["B2AZ_TESTS"])
import b2az, parity_harness as ph
from b2az import dist as bd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = b2az.load(ph.HOSTEMU_LIB)
base = b2az.default_params(lib, games_to_play=22, concurrent_games=10, mcts_visits=(24, 24), eval_type=b2az.EVAL_RANDOM,
                           rng_mode=b2az.RNG_GLOBAL, seed=100, history_enabled=1, self_play=1, **ph.level_params(1))
p = bd.shard_params(base, rank, world)
p.seed = 100 + rank  # the single-process check below replays exactly these two pools
e = b2az.Engine(p, lib=lib)
while e.stats().active_games:
    e.step(32)
st = e.stats()
red = bd.allreduce_stats(st)
hist = e.drain_history(1 << 16)
got = bd.gather_history(*hist)
tmax = bd.max_over_ranks(rank + 1.5)
if rank == 0:
    c, v, pi = [x.numpy() for x in got]
    np.savez(os.environ["B2AZ_OUT"], canon=c, v=v, pi=pi)
    print(json.dumps({"red": red, "tmax": tmax, "local_games": [p.games_to_play, p.concurrent_games]}))
dist.destroy_process_group()
"""


def test_shard_games_covers_the_budget():
    for total, world in [(22, 2), (65536, 8), (7, 8), (100, 3)]:
        spans = [bd.shard_games(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_shard_slots_keeps_every_slot_on_its_own_stream():
    for n, world in [(8192, 8), (10, 3), (5, 8)]:
        seen = []
        for r in range(world):
            k, seed, lo = bd.shard_slots(n, 1000, r, world)
            seen += [seed + g for g in range(k)]
            assert seed == 1000 + lo
        assert seen == list(range(1000, 1000 + n))


def test_two_rank_run_matches_single_process(tmp_path):
    out = tmp_path / "hist.npz"
    env = dict(os.environ, B2AZ_PKG=os.path.join(ph.ROOT, "alphazero-pybind11_b200"), B2AZ_TESTS=os.path.join(ph.ROOT, "tests"),
               B2AZ_OUT=str(out), MASTER_ADDR="127.0.0.1")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29731", str(script)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["tmax"] == 2.5 and res["local_games"] == [11, 5]
    # the same two pools in this process
    lib = b2az.load(ph.HOSTEMU_LIB)
    tot = dict(sims=0, moves=0, games=0, scores=np.zeros(3), hist=[])
    for rank in range(2):
        p = b2az.default_params(lib, games_to_play=11, concurrent_games=5, mcts_visits=(24, 24), eval_type=b2az.EVAL_RANDOM,
                                rng_mode=b2az.RNG_GLOBAL, seed=100 + rank, history_enabled=1, self_play=1, **ph.level_params(1))
        e = b2az.Engine(p, lib=lib)
        while e.stats().active_games:
            e.step(32)
        st = e.stats()
        tot["sims"] += st.simulations; tot["moves"] += st.moves; tot["games"] += st.games_completed
        tot["scores"] += np.array(st.scores[:])
        tot["hist"].append(e.drain_history(1 << 16))
        e.close()
    red = res["red"]
    assert (red["simulations"], red["moves"], red["games_completed"]) == (tot["sims"], tot["moves"], tot["games"]) and tot["games"] == 22
    assert red["scores"] == tot["scores"].tolist() and red["active_games"] == 0
    got = np.load(out)
    want = [np.concatenate([h[i] for h in tot["hist"]]) for i in range(3)]
    assert np.array_equal(got["canon"], want[0]) and np.array_equal(got["v"], want[1])
    assert np.array_equal(got["pi"].view(np.uint32), want[2].view(np.uint32))


TAFL_WORKER = r"""
import os, sys, json, zlib
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["B2AZ_PKG"]); sys.path.insert(0, os.environ["B2AZ_TESTS"])
import b2az, parity_harness as ph
from b2az import dist as bd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = b2az.load(ph.HOSTEMU_LIB)
g = np.load(os.path.join(ph.ROOT, "tests", "golden", "tafl_brandubh_transcripts.npz"))
idx = np.nonzero(g["max_turns"] == 150)[0]
lo, hi = bd.shard_games(len(idx), rank, world)          # independent transcripts: rank r replays its own slice
mine = idx[lo:hi]
r = b2az.tafl_replay(0, g["moves"][mine], g["lens"][mine], 150, lib=lib)
positions = int((g["lens"][mine] + 1).sum())
ends = np.array([r["terminal"][j, g["lens"][i]] for j, i in enumerate(mine)], np.int64)
t = torch.tensor([positions, (ends == 1).sum(), (ends == 2).sum(), (ends == 3).sum()], dtype=torch.float64)
dist.all_reduce(t)                                       # the only collective: additive statistics
crc = np.array([zlib.crc32(r["canonical"][j, : g["lens"][i] + 1].tobytes()) for j, i in enumerate(mine)], np.int64)
got = bd.gather_history(crc.reshape(-1, 1), ends.reshape(-1, 1), mine.reshape(-1, 1).astype(np.int64))
if rank == 0:
    print(json.dumps({"tot": t.tolist(), "crc": got[0].flatten().tolist(), "ends": got[1].flatten().tolist(),
                      "idx": got[2].flatten().tolist()}))
dist.destroy_process_group()
"""


def test_two_rank_tafl_replay_shards_match_single_process(tmp_path):
    """The tafl game kernels shard like the self-play pool: independent transcripts per rank, no data-path collective."""
    import json
    import zlib

    env = dict(os.environ, B2AZ_PKG=os.path.join(ph.ROOT, "alphazero-pybind11_b200"), B2AZ_TESTS=os.path.join(ph.ROOT, "tests"),
               MASTER_ADDR="127.0.0.1")
    script = tmp_path / "tafl_worker.py"
    script.write_text(TAFL_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29741", str(script)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    g = np.load(os.path.join(ph.ROOT, "tests", "golden", "tafl_brandubh_transcripts.npz"))
    idx = np.nonzero(g["max_turns"] == 150)[0]
    one = b2az.tafl_replay(0, g["moves"][idx], g["lens"][idx], 150, lib=b2az.load(ph.HOSTEMU_LIB))
    ends = [int(one["terminal"][j, g["lens"][i]]) for j, i in enumerate(idx)]
    crc = [zlib.crc32(one["canonical"][j, : g["lens"][i] + 1].tobytes()) for j, i in enumerate(idx)]
    assert res["idx"] == idx.tolist() and res["ends"] == ends and res["crc"] == crc
    assert res["tot"] == [float((g["lens"][idx] + 1).sum()), float(ends.count(1)), float(ends.count(2)), float(ends.count(3))]


BUDGET_WORKER = r"""
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["B2AZ_PKG"]); sys.path.insert(0, os.environ["B2AZ_TESTS"])
import b2az, parity_harness as ph
from b2az import dist as bd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = b2az.load(ph.HOSTEMU_LIB)
# weight broadcast: rank 0's "network" reaches every rank, dtypes and shapes kept
net = torch.nn.Sequential(torch.nn.Conv2d(4, 3, 3), torch.nn.BatchNorm2d(3), torch.nn.Linear(5, 2))
torch.manual_seed(rank)
for p_ in net.parameters():
    p_.data.normal_()
nbytes = bd.broadcast_weights(net)
digest = float(sum(p_.double().sum() for p_ in net.state_dict().values() if p_.dtype.is_floating_point))
# one GLOBAL games_to_play budget over both ranks (play_manager.cc:506-513)
base = b2az.default_params(lib, games_to_play=10 ** 6, concurrent_games=12, mcts_visits=(16, 16), eval_type=b2az.EVAL_RANDOM,
                           rng_mode=b2az.RNG_PER_GAME, seed=500, history_enabled=0, self_play=1, **ph.level_params(0))
p = bd.shard_params(base, rank, world)
e = b2az.Engine(p, lib=lib)
budget = bd.GlobalBudget(e, 40)
for it in range(100000):
    e.step(16 if rank == 0 else 48)   # the ranks progress at different speeds
    g = budget.sync()
    if g["active_games"] == 0:
        break
st = e.stats()
red = bd.allreduce_stats(st)
with open(os.environ["B2AZ_OUT"] + f".{rank}", "w") as f:
    json.dump({"nbytes": nbytes, "digest": digest, "red": red, "local_seed": int(p.seed), "local_games": int(st.games_completed)}, f)
dist.destroy_process_group()
"""


def test_global_budget_weight_broadcast_and_raw_stat_reduction(tmp_path):
    import json

    env = dict(os.environ, B2AZ_PKG=os.path.join(ph.ROOT, "alphazero-pybind11_b200"), B2AZ_TESTS=os.path.join(ph.ROOT, "tests"),
               MASTER_ADDR="127.0.0.1", B2AZ_OUT=str(tmp_path / "out.json"))
    script = tmp_path / "budget_worker.py"
    script.write_text(BUDGET_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29751", str(script)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r0, r1 = (json.load(open(str(tmp_path / "out.json") + f".{k}")) for k in (0, 1))
    assert r0["digest"] == r1["digest"] and r0["nbytes"] > 0, "every rank holds rank 0's weights after the broadcast"
    assert (r0["local_seed"], r1["local_seed"]) == (500, 506), "rank r owns slots [6r, 6r + 6): seed + lo"
    red = r0["red"]
    # 12 slots, a budget of 40 games: all slots retire; the overshoot is bounded by the games started between two syncs
    assert red["active_games"] == 0 and 40 <= red["games_completed"] <= 40 + 12
    assert red["games_completed"] == r0["local_games"] + r1["local_games"] and r1["local_games"] > r0["local_games"] > 0
    assert sum(red["scores"]) == red["games_completed"]
    assert red["avg_game_length"] > 7 and 0 < red["avg_leaf_depth"] < 10 and red["avg_valid_moves"] > 1
    with pytest.raises(ValueError, match="world size"):
        bd.shard_params(b2az.default_params(b2az.load(ph.HOSTEMU_LIB), concurrent_games=4), 0, 8)
