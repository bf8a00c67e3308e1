"""Throughput of PlayManager::play over a tafl game on the device (b2az_tafl_selfplay_*): `--games` concurrent game
slots (two search trees each), `--sims` simulations per move with the reference's dumb_eval evaluator fused in, the
self-play settings of configs/brandubh.yaml / open_tafl.yaml / tawlbwrdd.yaml (Gumbel root search m = 16, or --puct:
Dirichlet noise + temperature schedule + pruned policy targets), training samples captured and drained to the host
every `--drain` moves. Simulations/s and moves/s over `--moves` lock-step moves after `--warm` warm-up moves, next to
the UNMODIFIED reference PlayManager doing the same on the host (one thread per slot-sized run, `--cpu-seconds`).
    python tools/tafl_selfplay_bench.py [--game 0|1|2] [--games N] [--sims 120] [--moves 24]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import b2az  # noqa: E402
import tafl_ref  # noqa: E402

NAMES = {0: "brandubh", 1: "opentafl", 2: "tawlbwrdd"}
MAX_TURNS = {0: 150, 1: 400, 2: 400}  # configs/*.yaml max_turns
for _v, _n in enumerate(("skirmish", "showdown", "clash", "battle")):  # Star Gambit: --game 10 + variant / 20 + variant (Unified)
    NAMES[10 + _v], NAMES[20 + _v] = f"star gambit {_n}", f"star gambit unified ({_n})"
    MAX_TURNS[10 + _v] = MAX_TURNS[20 + _v] = 640  # staged training samples per game slot (a game is <= 200 turns of several actions)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", type=int, default=0)
    ap.add_argument("--games", type=int, default=8192)
    ap.add_argument("--sims", type=int, default=120)
    ap.add_argument("--moves", type=int, default=24)
    ap.add_argument("--warm", type=int, default=4)
    ap.add_argument("--drain", type=int, default=8)
    ap.add_argument("--max-turns", type=int, default=0)
    ap.add_argument("--puct", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--net", action="store_true", help="EvalType::NN: a random-init torch conv net between find_leaf and "
                                                       "process_result, zero copy (device canonical batch in, v / pi out)")
    a = ap.parse_args()
    # one process per GPU under torchrun (weak scaling: --games slots PER GPU, slot streams sharded with
    # b2az.dist.shard_slots, no data-path collective; NCCL only for the barrier and the max-over-ranks time)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        from b2az import dist as bd
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        _, seed, _ = bd.shard_slots(a.games * world, 1, rank, world)
    else:
        seed = 1
    mt = a.max_turns or MAX_TURNS[a.game]
    kw = (dict(epsilon=0.25, root_policy_temp=1.25, shaped_dirichlet=True, policy_target_pruning=True, start_temp=1.0,
               final_temp=0.2, temp_decay_half_life=10.0) if a.puct else dict(gumbel_m=16, root_policy_temp=1.25))
    words = 2 * (1 + 3 * a.sims * (1 + 8 * (48 if a.game == 0 or a.game >= 10 else 140)))
    ring = a.games * (a.drain + 2) * (1 if a.game >= 10 else 4)  # (a Star Gambit sample is 31 KB)
    sp = b2az.TaflSelfplay(a.game, a.games, mt, a.sims, games_per_slot=1 << 20, seed=seed, words_per_tree=words, device=local,
                           hist_capacity=ring,
                           lib=b2az.load(os.environ.get("B2AZ_LIB_PATH")), **kw)  # experiment builds via env
    stream = torch.cuda.current_stream().cuda_stream
    # the consumer's buffers: pinned host memory, allocated once (what a history saver would hold)
    cap = ring
    pinned = (torch.empty((cap, sp.P, sp.S, sp.S), dtype=torch.float32).pin_memory(), torch.empty((cap, 3), dtype=torch.float32).pin_memory(),
              torch.empty((cap, sp.A), dtype=torch.float32).pin_memory(), torch.empty(cap, dtype=torch.int32).pin_memory())
    host_out = (pinned[0].numpy(), pinned[1].numpy(), pinned[2].numpy(), pinned[3].numpy().view("uint32"))
    play = lambda n: sp.play(n, stream, want_active=False)
    if a.net:
        import ctypes as C
        S, P, A = sp.S, sp.P, sp.A
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(P, 64, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, padding=1),
                                  torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(64 * S * S, A + 3)).cuda().eval()
        canon_ptr = sp.find_leaf(stream)

        class _Dev:
            __cuda_array_interface__ = {"shape": (a.games, P, S, S), "typestr": "<f4", "data": (canon_ptr, False), "version": 3}
        x = torch.as_tensor(_Dev(), device="cuda")
        v_buf = torch.empty((a.games, 3), dtype=torch.float32, device="cuda")
        pi_buf = torch.empty((a.games, A), dtype=torch.float32, device="cuda")

        def evaluate():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x)
            torch.softmax(out[:, :A].float(), dim=1, out=pi_buf)
            torch.softmax(out[:, A:].float(), dim=1, out=v_buf)
            sp.process_result(C.c_void_p(v_buf.data_ptr()), C.c_void_p(pi_buf.data_ptr()), host=False, stream=stream)

        def play(n):
            for _ in range(n * a.sims):
                sp.find_leaf(stream)
                evaluate()
        evaluate()  # answers the leaf found above: the first simulation of the warm-up
        for _ in range(a.sims - 1):
            sp.find_leaf(stream)
            evaluate()
        a.warm -= 1
    play(a.warm)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    st0, _ = sp.slots()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples, t0 = 0, time.perf_counter()
    e0.record()
    done, seg = 0, []
    while done < a.moves:
        n = min(a.drain, a.moves - done)
        seg.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        seg[-1][0].record()
        play(n)
        seg[-1][1].record()
        done += n
        samples += len(sp.drain_history(stream, out=host_out)[1])  # finished games' samples to pinned host memory (synchronises)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    dev_ms = sum(x.elapsed_time(y) for x, y in seg)  # the search + move launches alone (samples stay in HBM)
    st1, err = sp.slots()
    sp.close()
    assert (err == 0).all() and (st1["error"] == 0).all(), (set(err.tolist()), set(st1["error"].tolist()))
    sims = int(st1["simulations"].sum() - st0["simulations"].sum())
    if world > 1:  # whole-job figures: units of all ranks over the slowest rank's time
        t = torch.tensor([ms, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, dev_ms = float(t[0]), float(t[1])
        c = torch.tensor([sims, samples], dtype=torch.int64, device="cuda")
        dist.all_reduce(c)
        sims, samples = int(c[0]), int(c[1])
        dist.destroy_process_group()
        if rank != 0:
            sys.exit(0)
    moves = sims // a.sims
    games = int(st1["games_completed"].sum() - st0["games_completed"].sum())
    full = max(1, int(st1["total_full_move_count"].sum()))
    # the unmodified reference PlayManager on one host thread, same settings
    t0 = time.perf_counter()
    ref_moves, n_ref = 0, 0
    while time.perf_counter() - t0 < a.cpu_seconds:
        r = tafl_ref.selfplay(a.game, 900 + n_ref, mt, 1, a.sims, **kw)
        ref_moves += len(r["v"])
        n_ref += 1
    cpu_s = time.perf_counter() - t0
    # algorithmic bytes per simulation (RANDOM-eval form: no canonical write, no evaluation read), with the
    # depth D and children k measured by the engine in this run; peak = MEASURED_PEAKS.json hbm_gbs (else 6650)
    D = float(st1["leaf_depth"].sum() / full)
    k = float(st1["valid_moves"].sum() / max(1, int(st1["total_move_count"].sum())))
    bps = D * (12 * k + 8) + D * 28 + (16 * k + 16) + (200 if a.game >= 10 else 3 * 16 + 8)  # state: SGState / three bitboards
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    ach = bps * sims / world / (dev_ms * 1e-3) / 1e9  # per GPU
    roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "kernel": "k_sp_search", "bytes_per_sim": bps, "note": "instruction / latency bound (profiles/*k_sp_search*): "
            "the fraction of the HBM roofline is reported for completeness"}
    print(json.dumps({
        "kernel": "k_sp_find_leaf + torch net + k_sp_process_result + k_sp_move" if a.net else "k_sp_search + k_sp_move",
        "evaluator": "torch conv net (2x64 conv + linear heads, bf16 autocast), zero-copy" if a.net else "dumb_eval", "workload": f"{NAMES[a.game]} self-play (PlayManager::play on the device), "
        f"{a.games} concurrent games, {a.sims} sims/move, " + ("PUCT + Dirichlet + pruned targets" if a.puct else "Gumbel m=16") +
        ", dumb_eval, tree reuse, history on", "n_gpus": world, "scaling": "weak", "ms": round(ms, 2), "moves_timed": a.moves,
        "simulations_per_second": sims / (ms * 1e-3), "moves_per_second": moves / (ms * 1e-3),
        "device_ms": round(dev_ms, 2), "simulations_per_second_device": sims / (dev_ms * 1e-3), "roofline": roof,
        "games_finished_in_window": games, "samples_drained": samples, "wall_s": round(wall, 3),
        "mean_leaf_depth_finished_games": float(st1["leaf_depth"].sum() / full),
        "mean_legal_moves_finished_games": float(st1["valid_moves"].sum() / max(1, int(st1["total_move_count"].sum()))),
        "cpu_baseline": {"value": ref_moves * a.sims / cpu_s, "unit": "sims/s", "moves_per_second": ref_moves / cpu_s, "cores": 1,
                         "kind": "reference", "sample": f"{n_ref} games of the unmodified reference PlayManager (concurrent_games=1, "
                                                       f"EvalType::RANDOM), same settings, one host thread"}}))
