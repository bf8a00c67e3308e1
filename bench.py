#!/usr/bin/env python
"""bench.py — MCTS simulations/s of the B200 self-play engine (BASELINE.json metric), one JSON line.

Workload (config.workload, BASELINE.json configs[1]): Connect4, 65,536 concurrent games PER GPU, 400
simulations per move, device-resident trees, EvalType::RANDOM evaluator on the device (uniform priors over
the legal moves, value 1/3 — src/game_state.h:160-173), cache off: "Throughput A" of SURVEY.md §8(d), the
same workload the reference's own benchmark uses (src/play_manager_bench.cc:39-48, RANDOM eval).

A "step" is one launch of the step kernel = `--gens` (default 400) iterations of PlayManager::play()'s
loop body (process_result -> [move] -> find_leaf) for every game slot, i.e. G*gens simulations.

  value        whole-job simulations/s, tree pools resident in HBM, CUDA events on the launching stream,
               barrier + synchronize on both sides, max over ranks
  e2e          same metric through the public API with HOST result buffers: every step additionally drains
               the finished training samples (canonical planes, value and policy targets: what self-play
               produces) into pinned host memory and reads the statistics struct back. This workload has no
               per-step host inputs (the evaluator is the reference's RANDOM backend), hence h2d bytes = 0.
  e2e_nn_host  the legacy per-leaf host round trip of the reference API (build_batch -> update_inferences
               with host buffers, py_wrapper.cc:449-504 / play_manager.cc:619-642) for comparison
  roofline     algorithmic HBM bytes of the step kernel / its measured duration vs MEASURED_PEAKS.json
  cpu_baseline the unmodified reference PlayManager (oracle/_ref) on this box's host cores, same workload

`--impl reference` times the reference's own CPU implementation instead (all host threads).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIMS = 400


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="concurrent games per GPU")
    ap.add_argument("--gens", type=int, default=400, help="loop iterations per game slot per step (one launch)")
    ap.add_argument("--preroll", type=int, default=24, help="untimed moves per game before warm-up (steady state)")
    ap.add_argument("--lanes", type=int, default=0, help="threads per game slot (0/1: one thread per game)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extra", action="store_true", help="skip the tafl legs (BASELINE.json configs[2..3], secondary)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ CPU side
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_cpu(seconds_warm, seconds_timed, threads, windows=1):
    """The reference PlayManager (unmodified, oracle/_ref/libazref.so) on `threads` worker threads, RANDOM eval,
    400 visits, concurrent_games = 64*threads (the StreamPool shape of src/play_manager_bench.cc:170-181),
    queue_shards = threads (src/config.py:428-433). Returns (kind, sims/s per window list)."""
    import refdriver

    if refdriver.available():
        L = refdriver.lib()
        G = 64 * threads
        cfg = refdriver.play_cfg(games_to_play=2 ** 31 - 1, concurrent_games=G, max_batch_size=G,
                                 queue_shards=min(threads, 255), cache_shards=1, mcts_visits=(SIMS, SIMS), cpuct=1.25,
                                 fpu_reduction=0.25, self_play=1, tree_reuse=1, eval_type=1, history_enabled=0)
        pm = refdriver.RefPlayManager(cfg)
        pm.start_workers(threads, 12345, True)
        time.sleep(seconds_warm)
        rates = []
        for _ in range(windows):
            c0, t0 = L.azref_pm_progress_sims(pm.h, SIMS), time.perf_counter()
            time.sleep(seconds_timed)
            c1, t1 = L.azref_pm_progress_sims(pm.h, SIMS), time.perf_counter()
            rates.append((c1 - c0) / (t1 - t0))
        L.azref_pm_stop(pm.h)
        pm.join()
        pm.close()
        return "reference", rates
    # fallback: the oracle port, one independent PlayManager per thread (no shared queue => an upper bound)
    import parity_harness as ph

    done = []

    def worker(k):
        kw = dict(G=64, visits=SIMS, eval_type=1, rng_mode=1, history=False, **ph.level_params(0))
        cal = ph.PortPM(games_to_play=128, seed=100 + k, **kw)  # calibrate, then size a bounded run
        t0 = time.perf_counter()
        cal.advance()
        dt = time.perf_counter() - t0
        cal.close()
        run = ph.PortPM(games_to_play=max(128, int(128 * seconds_timed * windows / dt)), seed=200 + k, **kw)
        t0 = time.perf_counter()
        run.advance()
        done.append(run.simulations() / (time.perf_counter() - t0))
        run.close()

    th = [threading.Thread(target=worker, args=(k,)) for k in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    return "port", [sum(done)] * windows


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = f"/tmp/b2az_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


# ------------------------------------------------------------------------------------ main arms
def reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    workers = max(1, threads - 1)  # reference default: mcts_workers = cpu_count - 1 (src/config.py:229, 440-441)
    per = max(1.0, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    kind, rates = run_reference_cpu(per * args.warmup, per, workers, windows=args.steps)
    v = sum(rates) / len(rates)
    sample = (f"{kind} PlayManager, Connect4, EvalType::RANDOM, {SIMS} sims/move, {workers} worker threads, "
              f"concurrent_games={64 * workers}, {args.steps} windows of {per:.1f} s after {per * args.warmup:.1f} s warm-up")
    line = {"impl": "reference", "metric": "mcts_simulations_per_second", "value": v, "unit": "sims/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": "sims/s", "cores": workers, "kind": kind, "sample": sample,
                             "cpu_model": cpu_model(), "host_threads": threads},
            "e2e": {"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": f"connect4 self-play, {args.games} concurrent games/GPU, {SIMS} sims/move, EvalType::RANDOM "
                        f"evaluator, tree reuse on, cache off (BASELINE.json configs[1], SURVEY.md 8d Throughput A)",
            "concurrent_games_per_gpu": args.games, "sims_per_move": SIMS, "gens_per_step": args.gens,
            "parallelism": f"games sharded over {n} GPU(s), no data-path collective",
            "l2_policy": "working set (tree pools, tens of GB) far larger than the 126 MB L2; no flush needed"}


def roofline_bytes_per_sim(avg_leaf_depth, avg_children):
    """DESIGN.md 'Algorithmic bytes per simulation' (RANDOM-eval variant: no canonical write,
    no evaluation read)."""
    D, k = avg_leaf_depth, avg_children
    select = D * (12.0 * k + 8.0)
    backprop = D * 28.0
    expand = 16.0 * k + 16.0
    state = 24.0
    return select + backprop + expand + state


def nn_device_leg(args, torch, b2az, local, stream, barrier, max_over_ranks, world, rank):
    """SURVEY.md 8d "Throughput B": the PyTorch net stays the evaluator, fed zero-copy from the engine's device
    leaf batch (b2az_leaf_batch_device) and answered in place (b2az_submit_eval_all): no host synchronisation and
    no host copy per generation. Random-init dense conv net of the connect4 default shape (depth 4, 12 channels,
    5x5 kernels: src/config.py:44-47), bf16 autocast, self-play settings of connect4.yaml, position cache of
    200,000 entries (src/config.py:197), history on."""
    import ctypes as C

    nn = torch.nn
    G = args.games

    class C4Net(nn.Module):
        def __init__(self, depth=4, ch=12, k=5):
            super().__init__()
            self.convs = nn.ModuleList()
            c_in = 4
            for _ in range(depth):  # dense connectivity: every layer sees all earlier feature maps
                self.convs.append(nn.Conv2d(c_in, ch, k, padding=k // 2))
                c_in += ch
            self.v_head = nn.Sequential(nn.Conv2d(c_in, 4, 1), nn.ReLU(), nn.Flatten(), nn.Linear(4 * 42, 3))
            self.pi_head = nn.Sequential(nn.Conv2d(c_in, 4, 1), nn.ReLU(), nn.Flatten(), nn.Linear(4 * 42, 7))

        def forward(self, x):
            for conv in self.convs:
                x = torch.cat([x, torch.relu(conv(x))], 1)
            return torch.softmax(self.v_head(x).float(), 1), torch.softmax(self.pi_head(x).float(), 1)

    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True  # the 6x7 boards with 4..52 channels are far from cuDNN's default heuristics
    net = C4Net().cuda().eval().to(memory_format=torch.channels_last)
    # Two half-populations, each on its own stream: while the net evaluates the leaves of one half, the step kernel of the
    # other half runs (the reference overlaps the same way: its MCTS threads keep going while the GPU thread evaluates).
    H = 2
    Gh = G // H
    streams = [torch.cuda.Stream() for _ in range(H)]
    engs = []
    for h in range(H):
        p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=Gh, mcts_visits=(SIMS, SIMS), cpuct=1.25,
                                fpu_reduction=0.25, epsilon=0.25, mcts_root_temp=1.25, start_temp=1.0, final_temp=0.2,
                                temp_decay_half_life=10.0, root_fpu_zero=1, shaped_dirichlet=1, policy_target_pruning=1,
                                eval_type=b2az.EVAL_NN, rng_mode=b2az.RNG_PER_GAME, seed=3000 + 7 * rank + h, tree_reuse=1,
                                history_enabled=1, self_play=1, max_cache_size=200000 // H, history_capacity=8 * Gh)
        engs.append(b2az.Engine(p, device=local))

    class _View:  # zero-copy: torch wraps the engine's device batch through __cuda_array_interface__
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 3}

    # The net runs on the rows of the generation only, in a power-of-two bucket; every bucket is a captured CUDA graph
    # over static buffers — the reference's own scheme (neural_net.py:513-561). Cache hits never become rows: a game
    # whose leaf is in the cache goes on with its next simulation inside the step kernel.
    buckets = [dict() for _ in range(H)]
    x_all = [None] * H
    rows_seen = []

    def bucket_for(h, n):
        b = 64
        while b < n:
            b *= 2
        b = min(b, Gh)
        if b not in buckets[h]:
            with torch.cuda.stream(streams[h]):
                xin = torch.zeros((b, 4, 6, 7), dtype=torch.float32, device="cuda").contiguous(memory_format=torch.channels_last)
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    for _ in range(3):  # warm-up outside capture (cudnn autotune, lazy init)
                        net(xin)
                streams[h].synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=streams[h]):
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                        v, pi = net(xin)
                        v_s, pi_s = v.float().contiguous(), pi.float().contiguous()
            buckets[h][b] = (gr, xin, v_s, pi_s)
        return buckets[h][b]

    def evaluate(h):  # leaves of half h -> net -> evaluations, everything on stream h
        st = streams[h].cuda_stream
        n, cptr, iptr = engs[h].leaf_batch(st)  # synchronises stream h to read the row count (8 bytes)
        rows_seen.append(n)
        if n == 0:
            return
        with torch.cuda.stream(streams[h]):
            if x_all[h] is None:
                x_all[h] = torch.as_tensor(_View(cptr, (Gh, 4, 6, 7)), device="cuda")
            gr, xin, v_s, pi_s = bucket_for(h, n)
            xin[:n].copy_(x_all[h][:n])
            gr.replay()
        engs[h].submit_eval(v_s.data_ptr(), pi_s.data_ptr(), n)

    def generation():  # one generation of BOTH halves, software-pipelined
        for h in range(H):
            evaluate(h)                              # waits for half h's step, then queues its net
            engs[h].step(1, streams[h].cuda_stream)  # queued behind the net on stream h; overlaps the other half's net

    for h in range(H):
        for b in (64, 256, 1024, 4096, 16384, Gh):  # no capture inside the timed region
            bucket_for(h, min(b, Gh))
        engs[h].step(1, streams[h].cuda_stream)
    warm, timed = 80, 400
    for _ in range(warm):
        generation()
    torch.cuda.synchronize()
    barrier()
    s0 = [e.stats(streams[h].cuda_stream) for h, e in enumerate(engs)]
    rows_seen.clear()
    t0 = time.perf_counter()
    for _ in range(timed):
        generation()
    torch.cuda.synchronize()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    s1 = [e.stats(streams[h].cuda_stream) for h, e in enumerate(engs)]
    sims = sum(b_.simulations - a_.simulations for a_, b_ in zip(s0, s1))
    moves = sum(b_.moves - a_.moves for a_, b_ in zip(s0, s1))
    hits = sum(b_.cache_hits - a_.cache_hits for a_, b_ in zip(s0, s1))
    misses = sum(b_.cache_misses - a_.cache_misses for a_, b_ in zip(s0, s1))
    out = {"value": world * sims / (ms * 1e-3), "unit": "sims/s", "moves_per_second": world * moves / (ms * 1e-3),
           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * H,
           "generations": timed, "ms_per_generation": ms / timed,
           "mean_batch_rows": sum(rows_seen) / max(1, len(rows_seen)), "max_batch_rows": max(rows_seen or [0]),
           "net_rows_per_second": sum(rows_seen) / (ms * 1e-3),
           "cache_hit_rate": hits / max(1, hits + misses), "cache_entries": 200000,
           "device_error": int(max(x_.device_error for x_ in s1)),
           "net": "connect4 default arch (dense, depth 4, 12 channels, 5x5), random init, bf16 autocast, CUDA-graph buckets",
           "note": "two half-populations of concurrent_games/2 on two streams; per half and generation: b2az_step(1) [cache "
                   "hits are answered inside the step kernel and the game goes on to its next simulation] + k_canonicalize "
                   "of the missed leaves + 8-byte row-count read + torch net on exactly those rows + b2az_submit_eval; "
                   "leaf batch and evaluations never leave the device; wall-clock timed (two streams)"}
    for e in engs:
        e.close()
    return out


def c4_net(torch):
    """Random-init dense conv net of the connect4 default shape (depth 4, 12 channels, 5x5 kernels: src/config.py:44-47),
    shared by every NN leg (torch.manual_seed(0))."""
    nn = torch.nn

    class C4Net(nn.Module):
        def __init__(self, depth=4, ch=12, k=5):
            super().__init__()
            self.convs = nn.ModuleList()
            c_in = 4
            for _ in range(depth):
                self.convs.append(nn.Conv2d(c_in, ch, k, padding=k // 2))
                c_in += ch
            self.v_head = nn.Sequential(nn.Conv2d(c_in, 4, 1), nn.ReLU(), nn.Flatten(), nn.Linear(4 * 42, 3))
            self.pi_head = nn.Sequential(nn.Conv2d(c_in, 4, 1), nn.ReLU(), nn.Flatten(), nn.Linear(4 * 42, 7))

        def forward(self, x):
            for conv in self.convs:
                x = torch.cat([x, torch.relu(conv(x))], 1)
            return torch.softmax(self.v_head(x).float(), 1), torch.softmax(self.pi_head(x).float(), 1)

    torch.manual_seed(0)
    return C4Net().cuda().eval()


def reference_nn_leg(args, torch, seconds=8.0, warm=3.0):
    """SURVEY.md 8d "Throughput B", REFERENCE side: the UNMODIFIED reference PlayManager
    (oracle/_ref/libazref.so) with its MCTS worker threads on the box's host cores, its S3-FIFO cache (200,000 entries),
    the self-play flags of connect4.yaml, and the same random-init torch net on this B200 answering its leaf batches
    (bf16 autocast) — batches taken with build_batch and answered with update_inferences the way game_runner.py's
    batcher and result worker do (src/game_runner.py:648-727), one Python thread."""
    import numpy as np
    import refdriver

    if not refdriver.available():
        return {"unavailable": "oracle/_ref/libazref.so not built"}
    L = refdriver.lib()
    threads = host_threads()
    workers = max(1, threads - 2)  # one core for this batcher thread, as in the reference's thread budget
    G, B = 4096, 4096
    cfg = refdriver.play_cfg(games_to_play=2 ** 31 - 1, concurrent_games=G, max_batch_size=B, max_cache_size=200000,
                             queue_shards=min(workers, 255), cache_shards=min(threads, 255), mcts_visits=(SIMS, SIMS), cpuct=1.25,
                             fpu_reduction=0.25, epsilon=0.25, mcts_root_temp=1.25, start_temp=1.0, final_temp=0.2,
                             temp_decay_half_life=10.0, root_fpu_zero=1, shaped_dirichlet=1, policy_target_pruning=1,
                             self_play=1, tree_reuse=1, eval_type=0, history_enabled=1)
    pm = refdriver.RefPlayManager(cfg)
    net = c4_net(torch)
    pm.start_workers(workers, 4242, True)
    xin = torch.empty((B, 4, 6, 7), dtype=torch.float32).pin_memory()
    rows, batches = 0, 0
    t_start = time.perf_counter()
    c0 = t0 = None
    hist_drained = 0
    while True:
        now = time.perf_counter()
        if c0 is None and now - t_start >= warm:
            c0, t0 = L.azref_pm_progress_sims(pm.h, SIMS), now
            rows = batches = 0
        if c0 is not None and now - t0 >= seconds:
            break
        ids, canon = pm.build_batch(max_rows=B)
        n = len(ids)
        if n == 0:
            continue
        xin[:n].copy_(torch.from_numpy(canon))
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            v, pi = net(xin[:n].cuda(non_blocking=True))
        pm.update_inferences(ids, v.float().cpu().numpy(), pi.float().cpu().numpy())
        rows += n
        batches += 1
        if batches % 64 == 0:  # the history saver: keep history_ from growing without bound
            hist_drained += len(pm.drain_history(65536)[1])
    c1, t1 = L.azref_pm_progress_sims(pm.h, SIMS), time.perf_counter()
    cs = np.zeros(6, np.uint64)
    L.azref_pm_cache_stats(pm.h, refdriver.P(cs))
    L.azref_pm_stop(pm.h)
    pm.join()
    pm.close()
    dt = t1 - t0
    return {"value": (c1 - c0) / dt, "unit": "sims/s", "cores": workers, "kind": "reference", "host_threads": threads,
            "concurrent_games": G, "max_batch_size": B, "mean_batch_rows": rows / max(1, batches),
            "net_rows_per_second": rows / dt, "cache_hit_rate": float(cs[0]) / max(1.0, float(cs[0] + cs[1])),
            "cache_entries": 200000, "seconds": dt,
            "note": "the unmodified reference PlayManager + its cache on the host cores, the same torch net on this B200 "
                    "(bf16 autocast), build_batch / update_inferences from one Python thread"}


def pybind_dlpack_leg(args, torch, local):
    """SURVEY.md 8d "Throughput B", this repo's side THROUGH THE DROP-IN MODULE: alphazero.PlayManager (csrc/py_alphazero.cc)
    with the zero-copy DLPack feed — leaf_batch_dlpack hands torch the canonical batch and the legal-move masks in the
    engine's device buffers, update_inferences_dlpack takes the CUDA tensors back; no host copy per leaf."""
    import importlib.util
    import sysconfig

    path = os.path.join(ROOT, "alphazero-pybind11_b200", "alphazero" + sysconfig.get_config_var("EXT_SUFFIX"))
    spec = importlib.util.spec_from_file_location("alphazero", path)
    az = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(az)
    G = args.games
    p = az.PlayParams()
    p.games_to_play, p.concurrent_games, p.max_batch_size = 2 ** 31 - 1 - (2 ** 31 - 1) % G, G, G
    p.mcts_visits = [SIMS, SIMS]
    p.model_groups = [0, 0]
    p.history_enabled = p.self_play = p.tree_reuse = True
    p.cpuct, p.fpu_reduction, p.epsilon, p.mcts_root_temp = 1.25, 0.25, 0.25, 1.25
    p.start_temp, p.final_temp, p.temp_decay_half_life = 1.0, 0.2, 10.0
    p.root_fpu_zero = p.shaped_dirichlet = p.policy_target_pruning = True
    p.max_cache_size = 200000
    p.seed, p.device = 777, local
    pm = az.PlayManager(az.Connect4GS(), p)
    net = c4_net(torch).to(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True

    def generation():
        canon, valid, n = pm.leaf_batch_dlpack(0)
        if n == 0:
            return 0
        x = torch.from_dlpack(canon)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            v, pi = net(x)
        pi = pi * torch.from_dlpack(valid)  # the reference masks the policy head with valid_moves (neural_net.py)
        pm.update_inferences_dlpack(0, v.float().contiguous(), pi.float().contiguous())
        return n

    for _ in range(40):
        generation()
    torch.cuda.synchronize()
    s0 = pm.simulations()
    h0, m0 = pm.cache_hits(), pm.cache_misses()
    rows, gens = 0, 200
    t0 = time.perf_counter()
    for _ in range(gens):
        rows += generation()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sims = pm.simulations() - s0
    hits, misses = pm.cache_hits() - h0, pm.cache_misses() - m0
    pm.stop()
    return {"value": sims / dt, "unit": "sims/s", "generations": gens, "ms_per_generation": dt / gens * 1e3,
            "mean_batch_rows": rows / gens, "net_rows_per_second": rows / dt, "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": 4, "cache_hit_rate": hits / max(1, hits + misses), "cache_entries": 200000,
            "note": "alphazero.PlayManager (the drop-in pybind module) + leaf_batch_dlpack / update_inferences_dlpack + "
                    "torch net (eager, bf16 autocast, channels_last); one 4-byte row-count read per generation"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import b2az

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream
    G, gens, K, W = args.games, args.gens, args.steps, args.warmup

    def make(history):
        p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=G, mcts_visits=(SIMS, SIMS), cpuct=1.25,
                                fpu_reduction=0.25, eval_type=b2az.EVAL_RANDOM, rng_mode=b2az.RNG_PER_GAME,
                                seed=1000 + rank, tree_reuse=1, history_enabled=int(history), self_play=1,
                                lanes_per_game=args.lanes, history_capacity=(8 * G if history else 0))
        return b2az.Engine(p, device=local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- value: device-resident throughput
    eng = make(history=False)
    for _ in range(args.preroll):
        eng.step(SIMS, stream)
    for _ in range(W):
        eng.step(gens, stream)
    barrier()
    s0 = eng.stats(stream)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    evs[0].record()
    for i in range(K):
        eng.step(gens, stream)
        evs[i + 1].record()
    torch.cuda.synchronize()
    total_ms = evs[0].elapsed_time(evs[K])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    s1 = eng.stats(stream)
    if s1.device_error:
        raise SystemExit(f"bench.py: device error bits {s1.device_error}")
    sims_rank = s1.simulations - s0.simulations
    assert sims_rank == G * gens * K, (sims_rank, G * gens * K)
    moves_rank = s1.moves - s0.moves
    kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    total_ms = max_over_ranks(total_ms)
    value = world * sims_rank / (total_ms * 1e-3)
    moves_per_s = world * moves_rank / (total_ms * 1e-3)
    depth = float(s1.avg_leaf_depth) if s1.avg_leaf_depth > 0 else 5.0
    kids = float(s1.avg_valid_moves) if s1.avg_valid_moves > 0 else 6.5
    pool = (int(s1.pool_pages_total), int(s1.pool_pages_free))
    eng.close()

    # ---------------------------------------------------------------- e2e: public API, host result buffers
    e2e = None
    e2e_nn = None
    if not args.no_e2e:
        eng = make(history=True)
        cap = 4 * G
        h_canon = torch.empty((cap, 4, 6, 7), dtype=torch.float32).pin_memory()
        h_v = torch.empty((cap, 3), dtype=torch.float32).pin_memory()
        h_pi = torch.empty((cap, 7), dtype=torch.float32).pin_memory()

        # Software pipeline, as the reference's own threads do it (hist_saver next to play(), game_runner.py:729-745):
        # step k + 1 is enqueued, then the samples of step k are drained on a second stream while it runs, then the
        # statistics are read (which waits for step k + 1). Every sample reaches pinned host memory inside the timed region.
        drain_stream = torch.cuda.Stream()
        ds = drain_stream.cuda_stream
        marked = [False]

        def e2e_step():
            eng.step(gens, stream)
            n = 0
            if marked[0]:
                n = eng.drain_history_marked_into(h_canon.data_ptr(), h_v.data_ptr(), h_pi.data_ptr(), cap, ds)
            eng.history_mark(stream)
            marked[0] = True
            st = eng.stats(stream)
            return n, st

        for _ in range(args.preroll):
            eng.step(SIMS, stream)
            eng.history_mark(stream)
            eng.drain_history_marked_into(h_canon.data_ptr(), h_v.data_ptr(), h_pi.data_ptr(), cap, ds)
        marked[0] = False
        ctl_net = None

        def control_plane(st, n):
            """the multi-GPU control plane of a self-play run, once per K steps INSIDE the timed region:
            updated weights from rank 0, the additive statistics of all ranks, the newest samples gathered on rank 0"""
            nonlocal ctl_net
            from b2az import dist as bd

            if ctl_net is None:  # a stand-in with the parameter count of the connect4 default net (config.py:44-47)
                ctl_net = torch.nn.Sequential(torch.nn.Conv2d(4, 12, 5, padding=2), torch.nn.Conv2d(16, 12, 5, padding=2),
                                              torch.nn.Conv2d(28, 12, 5, padding=2), torch.nn.Conv2d(40, 12, 5, padding=2),
                                              torch.nn.Linear(168, 3), torch.nn.Linear(168, 7)).cuda()
            b_w = bd.broadcast_weights(ctl_net)
            red = bd.allreduce_stats(st)
            m = max(1, min(int(n), 8192))  # the newest samples of this rank, device resident
            got = bd.gather_history(h_canon[:m].cuda(non_blocking=True), h_v[:m].cuda(non_blocking=True), h_pi[:m].cuda(non_blocking=True))
            return {"broadcast_bytes": int(b_w), "allreduce_bytes": int(red["nccl_bytes"]),
                    "gather_bytes": int(bd.gather_history.last_nccl_bytes), "gathered_samples": int(got[0].shape[0]) if got else None,
                    "global_simulations": red["simulations"]}

        for _ in range(W):
            n_w, st_w = e2e_step()
        if dist is not None:
            control_plane(st_w, n_w)  # warm-up: NCCL sets its channels up on first use
        barrier()
        st0 = eng.stats(stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        a.record()
        samples = 0
        for _ in range(K):
            n, st = e2e_step()
            samples += n
        samples += eng.drain_history_marked_into(h_canon.data_ptr(), h_v.data_ptr(), h_pi.data_ptr(), cap, ds)  # the last step's
        nccl = control_plane(st, n) if dist is not None else None
        b.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ev_ms = a.elapsed_time(b)
        barrier()
        e_ms = max_over_ranks(max(wall_ms, ev_ms))
        e_sims = st.simulations - st0.simulations
        e2e = {"value": world * e_sims / (e_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(samples / K * (168 + 3 + 7) * 4 + C.sizeof(b2az.Stats) + 16),
               "samples_per_step": samples / K, "ms_per_step": e_ms / K, "nccl_per_k_steps": nccl,
               "note": "step kernel + drain of finished training samples to pinned host buffers (overlapped with the next "
                       "step on a second stream: b2az_history_mark / b2az_drain_history_marked) + stats read, every step; "
                       "with N > 1 one weight broadcast + one stats all-reduce + one sample gather per K steps inside the "
                       "timed region; the RANDOM-eval workload has no per-step host inputs"}
        eng.close()

        # legacy per-leaf host round trip (EVAL_NN + host buffers), a few generations
        p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=G, mcts_visits=(SIMS, SIMS), cpuct=1.25,
                                fpu_reduction=0.25, eval_type=b2az.EVAL_NN, rng_mode=b2az.RNG_PER_GAME,
                                seed=2000 + rank, tree_reuse=1, history_enabled=0, self_play=1, lanes_per_game=args.lanes)
        eng = b2az.Engine(p, device=local)
        hb = torch.empty((G, 4, 6, 7), dtype=torch.float32).pin_memory()
        hid = torch.empty((G,), dtype=torch.int32).pin_memory()
        hv = torch.full((G, 3), 1.0 / 3.0, dtype=torch.float32).pin_memory()
        hp = torch.full((G, 7), 1.0 / 7.0, dtype=torch.float32).pin_memory()

        def nn_gen():
            eng.step(1, stream)
            n = eng.leaf_batch_host_into(hb.data_ptr(), hid.data_ptr(), G, stream)
            eng.submit_eval_host_from(hid.data_ptr(), hv.data_ptr(), hp.data_ptr(), n, stream)
            return n

        ngen = 40
        for _ in range(10):
            nn_gen()
        barrier()
        t0 = time.perf_counter()
        leaves = 0
        for _ in range(ngen):
            leaves += nn_gen()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_nn = {"value": world * leaves / dt, "unit": "sims/s", "h2d_bytes_per_step": G * (3 + 7) * 4,
                  "d2h_bytes_per_step": G * (168 * 4 + 4), "generations": ngen,
                  "note": "reference-API compatibility path: one generation = b2az_step(1) + b2az_leaf_batch_host "
                          "(fp32 canonical planes D2H) + b2az_submit_eval_host (v, pi H2D); pinned buffers"}
        eng.close()

    # ---------------------------------------------------------------- NN evaluator on the device, zero host sync
    e2e_nn_dev = None
    if not args.no_e2e:
        e2e_nn_dev = nn_device_leg(args, torch, b2az, local, stream, barrier, max_over_ranks, world, rank)

    # ---------------------------------------------------------------- NN through the drop-in module (DLPack) and the
    # reference's own PlayManager with the same net on the same GPU (Throughput B, both arms)
    e2e_nn_pybind = e2e_nn_reference = None
    if not args.no_e2e and world == 1:
        for name, fn in (("pybind", lambda: pybind_dlpack_leg(args, torch, local)), ("reference", lambda: reference_nn_leg(args, torch))):
            try:
                r = fn()
            except Exception as ex:  # a secondary leg never breaks the headline line
                r = {"failed": repr(ex)[:300]}
            if name == "pybind":
                e2e_nn_pybind = r
            else:
                e2e_nn_reference = r

    # ---------------------------------------------------------------- roofline + cpu baseline (rank 0)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bps = roofline_bytes_per_sim(depth, kids)
    launch_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = (bps * G * gens) / (launch_ms * 1e-3) / 1e9
    # measured DRAM traffic of the step kernel: ncu cannot run inside a timed bench, so the per-simulation figure comes from
    # the committed ncu capture — tied to the kernel and to the hash of the kernel source it was taken on, and flagged
    # stale (not silently reused) when either no longer matches what just ran
    traffic, traffic_src, traffic_stale = None, None, None
    kernel_ran = {"f": "k_step", "w": "k_step_w", "q": "k_step_q"}.get(os.environ.get("B2AZ_STEP_KERNEL", "s")[:1], "k_step_sync")
    try:
        import hashlib

        prof = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))
        traffic = prof["dram_bytes_per_simulation"] * G * gens  # per launch, like `achieved`
        sha = hashlib.sha256(open(os.path.join(ROOT, "alphazero-pybind11_b200", "csrc", "az_engine_logic.h"), "rb").read()).hexdigest()
        traffic_stale = prof.get("kernel") != kernel_ran or prof.get("logic_header_sha256") != sha
        traffic_src = prof.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "traffic_stale": traffic_stale,
                "kernel": kernel_ran, "bytes_per_sim": bps, "avg_leaf_depth": depth,
                "avg_children": kids, "launch_ms": launch_ms, "algorithmic_bytes_per_launch": bps * G * gens,
                "note": "latency bound, not bandwidth bound: one dependent 160 B block load per tree level per simulation, "
                        "1.6 us unloaded (profiles/r2o_tlb_probe2.jsonl); the 32 games of a warp run in lock step "
                        "(DESIGN.md 3); tensor cores unused by design",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        workers = max(1, threads - 1)
        kind, rates = run_reference_cpu(min(4.0, args.cpu_seconds / 3), args.cpu_seconds * 2 / 3, workers)
        _, r1 = run_reference_cpu(1.5, 4.0, 1)
        cpu = {"value": rates[0], "unit": "sims/s", "cores": workers, "kind": kind, "value_1_thread": r1[0],
               "cpu_model": cpu_model(), "host_threads": threads,
               "sample": f"{kind} PlayManager, Connect4, EvalType::RANDOM, {SIMS} sims/move, {workers} worker threads, "
                         f"concurrent_games={64 * workers}, {args.cpu_seconds * 2 / 3:.0f} s window after warm-up"}
    # ---------------------------------------------------------------- secondary: the tafl configs (parity-test
    # cases of BASELINE.json, not the headline): wide-tree search and game kernels, each in its own process
    extra = None
    if world == 1 and not args.no_extra:
        extra = {}
        for key, cmd in (("brandubh_gumbel_search", ["tools/forest_bench.py", "--game", "0", "--trees", "8192", "--moves", "6",
                                                     "--gumbel-m", "16"]),
                         ("brandubh_selfplay", ["tools/tafl_selfplay_bench.py", "--game", "0", "--games", "8192", "--moves", "16",
                                                "--cpu-seconds", "5"]),
                         ("opentafl_game_kernels", ["tools/tafl_bench.py", "--game", "1", "--games", "8192", "--reps", "3"]),
                         # BASELINE.json configs[4]: Star Gambit Unified with the Gumbel root search (configs/star_gambit_unified.yaml);
                         # one leg per end of the variant mix (11x11 Skirmish, 13x13 Battle), 120 simulations per move
                         ("star_gambit_unified_battle_selfplay", ["tools/tafl_selfplay_bench.py", "--game", "23", "--games", "8192",
                                                                  "--moves", "16", "--cpu-seconds", "5"]),
                         ("star_gambit_unified_skirmish_selfplay", ["tools/tafl_selfplay_bench.py", "--game", "20", "--games", "8192",
                                                                    "--moves", "16", "--cpu-seconds", "5"]),
                         ("star_gambit_game_kernels", ["tools/sg_bench.py", "--game", "23", "--games", "4096", "--moves", "64"])):
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, cmd[0])] + cmd[1:], stdout=subprocess.PIPE,
                                   stderr=subprocess.PIPE, text=True, timeout=300)
                rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
                extra[key] = json.loads(rows[-1]) if rows else {"failed": r.stderr[-300:]}
            except Exception as ex:  # never let a secondary leg break the headline line
                extra[key] = {"failed": repr(ex)[:300]}
    line = {"metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "moves_per_second": moves_per_s, "clocks": clocks, "e2e": e2e, "e2e_nn_host": e2e_nn,
            "e2e_nn_device": e2e_nn_dev, "e2e_nn_pybind_dlpack": e2e_nn_pybind, "e2e_nn_reference": e2e_nn_reference,
            "gpu_launches": K * world, "roofline": roofline, "cpu_baseline": cpu,
            "pool_pages": {"total": pool[0], "free": pool[1]}, "other_configs": extra}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
