"""Multi-GPU plumbing for the self-play pool: games are independent units, so rank r simply owns
its own pool of games — no collective on the data path. torch.distributed (NCCL on the GPU box, gloo in the CPU
test-suite) is used only for what the reference's single process does with shared memory: the global
`games_to_play` budget, the additive score / metric vectors, and gathering the finished training samples.

Reference counterparts: games_started_/games_completed_ accounting (play_manager.cc:506-513), scores_ and the
metric accumulators (play_manager.h:288-366), history_ (play_manager.cc:448-460)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_games(total_games, rank, world):
    """Contiguous split of a global game budget: rank r plays games [lo, hi). Sizes differ by at most one."""
    base, rem = divmod(int(total_games), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_params(params, rank, world, total_games=None, total_concurrent=None):
    """Per-rank copy of a b2az.Params: games_to_play / concurrent_games split over the ranks, a distinct RNG
    seed per rank so that per-game streams (seed, game_index) never collide across ranks."""
    import copy

    p = copy.copy(params)
    tg = params.games_to_play if total_games is None else total_games
    tc = params.concurrent_games if total_concurrent is None else total_concurrent
    lo, hi = shard_games(tg, rank, world)
    clo, chi = shard_games(tc, rank, world)
    p.games_to_play = max(hi - lo, chi - clo)
    p.concurrent_games = chi - clo
    p.seed = params.seed + 0x9E3779B97F4A7C15 * rank % (1 << 63)
    return p


def shard_slots(n_games, seed, rank, world):
    """Tafl self-play (b2az.TaflSelfplay): slot g draws from pcg32(seed + g), so rank r takes the contiguous slot range
    [lo, hi) and passes seed + lo — the union over the ranks is the single-process run, slot for slot
    (tests/test_tafl_selfplay.py::test_sharded_slots_equal_the_unsharded_run). Returns (n_games_local, seed_local, lo)."""
    lo, hi = shard_games(n_games, rank, world)
    return hi - lo, int(seed) + lo, lo


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_stats(stats):
    """Sum the additive counters of b2az.Stats over the ranks and recompute the derived means the way
    PlayManager's getters do (play_manager.h:288-315). Returns a dict."""
    s = stats
    n_full = float(s.avg_leaf_depth != 0)  # the engine exposes means; weight them by what they were means of
    vec = torch.tensor([s.simulations, s.moves, s.games_completed, s.scores[0], s.scores[1], s.scores[2],
                        s.avg_game_length * s.games_completed if s.games_completed else 0.0,
                        s.avg_leaf_depth * s.moves, s.avg_search_entropy * s.moves, s.avg_valid_moves * s.moves,
                        s.hist_count, s.active_games, n_full], dtype=torch.float64, device=_device())
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    v = vec.tolist()
    games, moves = v[2], v[1]
    return {"simulations": int(v[0]), "moves": int(moves), "games_completed": int(games), "scores": v[3:6],
            "avg_game_length": v[6] / games if games else 0.0, "avg_leaf_depth": v[7] / moves if moves else 0.0,
            "avg_search_entropy": v[8] / moves if moves else 0.0, "avg_valid_moves": v[9] / moves if moves else 0.0,
            "hist_count": int(v[10]), "active_games": int(v[11])}


def gather_history(canon, v, pi, dst=0):
    """Gather every rank's finished training samples (numpy or torch arrays, first dim = samples) on rank `dst`
    (others get None). Variable counts per rank: sizes are exchanged first, then padded all_gather."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device()
    t = [torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(dev) for a in (canon, v, pi)]
    n = torch.tensor([t[0].shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    cap = max(sizes) if sizes else 0
    out = []
    for a in t:
        pad = torch.zeros((cap,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev)
        pad[: a.shape[0]] = a
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out.append(torch.cat([p[:k] for p, k in zip(parts, sizes)], 0) if rank == dst else None)
    return tuple(out) if rank == dst else None


def max_over_ranks(value):
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
