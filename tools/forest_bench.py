"""Throughput of the batched wide-tree MCTS (b2az_forest_*) in greedy self-play with the reference's dumb_eval
evaluator: `--trees` trees each run `--sims` simulations per move (one fused launch), then play their most visited
move on the device; simulations/s over `--moves` moves, next to the unmodified reference's MCTS class doing the same
on one host core.   python tools/forest_bench.py [--game 0|1|2] [--trees N] [--sims 120] [--moves 12]"""
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
MAX_TURNS = {0: 150, 1: 400, 2: 400}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", type=int, default=0)
    ap.add_argument("--trees", type=int, default=16384)
    ap.add_argument("--sims", type=int, default=120)  # configs/brandubh.yaml mcts_visits
    ap.add_argument("--moves", type=int, default=12)
    ap.add_argument("--gumbel-m", type=int, default=0, help="Gumbel root search with m candidates (configs/brandubh.yaml: 16)")
    ap.add_argument("--net", action="store_true", help="evaluate the leaves with a random-init torch conv net on the device "
                                                       "(find_leaf -> net -> process_result, zero copy, no host sync)")
    a = ap.parse_args()
    # each half of a tree's slab: the kept subtree + one move's new nodes (1 + 7k words each), with head room
    words = 2 * (1 + 3 * a.sims * (1 + 8 * (48 if a.game == 0 else 140)))
    f = b2az.Forest(a.game, a.trees, MAX_TURNS[a.game], cpuct=1.25, fpu_reduction=0.25, seed=1, words_per_tree=words,
                    gumbel_m=a.gumbel_m, lib=b2az.load(os.environ.get("B2AZ_LIB_PATH")))  # experiment builds via env
    stream = torch.cuda.current_stream().cuda_stream
    if a.net:
        # the reference's evaluator shape (neural_net.py: conv trunk, policy + value heads), random init, bf16 autocast;
        # the canonical batch is read in place from the forest's device buffer and v / pi are handed back as pointers
        import ctypes as C
        S, P, A = f.S, f.P, f.A
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(P, 64, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, padding=1),
                                  torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(64 * S * S, A + 3)).cuda().eval()
        canon_ptr = f.find_leaf(stream)  # (also the first simulation of the warm-up move)

        class _Dev:
            __cuda_array_interface__ = {"shape": (a.trees, P, S, S), "typestr": "<f4", "data": (canon_ptr, False), "version": 3}
        x = torch.as_tensor(_Dev(), device="cuda")
        v_buf = torch.empty((a.trees, 3), dtype=torch.float32, device="cuda")
        pi_buf = torch.empty((a.trees, A), dtype=torch.float32, device="cuda")

        def evaluate():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x)
            torch.softmax(out[:, :A].float(), dim=1, out=pi_buf)
            torch.softmax(out[:, A:].float(), dim=1, out=v_buf)
            f.process_result_device(C.c_void_p(v_buf.data_ptr()), C.c_void_p(pi_buf.data_ptr()), stream=stream)

        def search():
            for _ in range(a.sims):
                f.find_leaf(stream)
                evaluate()
        evaluate()
        for _ in range(a.sims - 1):
            f.find_leaf(stream)
            evaluate()
    else:
        def search():
            f.simulate(a.sims, stream)
        if a.gumbel_m:
            f.set_gumbel_num_sims(a.sims, stream)
        search()  # warm-up move
    f.advance(stream)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.moves + 1)]
    ev[0].record()
    for m in range(a.moves):
        if a.gumbel_m:
            f.set_gumbel_num_sims(a.sims, stream)
        search()
        f.advance(stream)  # (greedy by visit count in both modes: the throughput does not depend on the move rule)
        ev[m + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.moves)]
    if a.gumbel_m:
        f.set_gumbel_num_sims(a.sims, stream)
    search()  # one more (untimed) search, kept un-advanced so that the depth statistics can be read
    torch.cuda.synchronize()
    _, _, info = f.counts(want_q=False)
    live = info["depth"] > 0
    leaf_depth = float(info["total_leaf_depth"][live].sum() / max(1, info["depth"][live].sum()))
    root_k = float(info["root_k"][live].mean()) if live.any() else 0.0
    f.close()
    assert (info["error"] == 0).all(), set(info["error"].tolist())
    sims_total = a.trees * a.sims * a.moves
    t0 = time.perf_counter()
    n_ref, ref_sims = 0, 0
    while time.perf_counter() - t0 < 10:
        r = tafl_ref.search(a.game, 900 + n_ref, a.moves + 1, a.sims, MAX_TURNS[a.game], 1.25, 0.25, False, None, a.gumbel_m)
        ref_sims += len(r[2]) * a.sims
        n_ref += 1
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"kernel": "k_forest_find_leaf + torch net + k_forest_process_result" if a.net else "k_forest_simulate",
                      "evaluator": "torch conv net (2x64 conv + linear heads, bf16 autocast), zero-copy" if a.net else "dumb_eval",
                      "game": NAMES[a.game], "trees": a.trees, "gumbel_m": a.gumbel_m, "sims_per_move": a.sims,
                      "moves": a.moves, "ms_per_move": [round(x, 2) for x in ms],
                      "simulations_per_second": sims_total / (sum(ms) * 1e-3), "moves_per_second": a.trees * a.moves / (sum(ms) * 1e-3),
                      "mean_slab_words_used": float(info["words_used"].mean()), "mean_leaf_depth": leaf_depth, "mean_root_children": root_k, "games_over": int((info["root_term"] != 0).sum()),
                      "cpu_baseline": {"value": ref_sims / cpu_s, "unit": "sims/s", "cores": 1, "kind": "reference",
                                       "sample": f"{n_ref} single-tree runs of the unmodified reference MCTS class, dumb_eval, same settings"}}))
