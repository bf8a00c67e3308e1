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
    """Per-rank copy of a b2az.Params: the slot range [lo, hi) of the global pool goes to rank r together with
    seed + lo, because slot g draws from pcg32(seed + g) — the union over the ranks is then the single-process run
    slot for slot. games_to_play is split the same way (a static share; GlobalBudget below keeps the budget global)."""
    import copy

    p = copy.copy(params)
    tg = params.games_to_play if total_games is None else total_games
    tc = params.concurrent_games if total_concurrent is None else total_concurrent
    if world > tc:
        raise ValueError(f"world size {world} > concurrent_games {tc}: a rank would own no game slot")
    lo, hi = shard_games(tg, rank, world)
    clo, chi = shard_games(tc, rank, world)
    p.games_to_play = max(hi - lo, chi - clo)
    p.concurrent_games = chi - clo
    p.seed = (params.seed + clo) % (1 << 64)
    return p


def shard_slots(n_games, seed, rank, world):
    """Tafl self-play (b2az.TaflSelfplay): slot g draws from pcg32(seed + g), so rank r takes the contiguous slot range
    [lo, hi) and passes seed + lo — the union over the ranks is the single-process run, slot for slot
    (tests/test_tafl_selfplay.py::test_sharded_slots_equal_the_unsharded_run). Returns (n_games_local, seed_local, lo)."""
    lo, hi = shard_games(n_games, rank, world)
    return hi - lo, int(seed) + lo, lo


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


_SUMS = ("simulations", "moves", "games_completed", "games_started", "active_games", "hist_count", "sum_game_length",
         "total_move_count", "full_move_count", "fast_move_count", "sum_leaf_depth", "sum_search_entropy",
         "fast_sum_leaf_depth", "fast_sum_search_entropy", "sum_valid_moves", "cache_hits", "cache_misses")


def allreduce_stats(stats, async_op=False):
    """Sum the RAW accumulators of b2az.Stats over the ranks (all of them are sums over completed games, like the
    reference's members behind play_manager.h:288-315) and recompute the means exactly as PlayManager's getters do.
    Returns a dict (or, with async_op, a (work, finish) pair: call finish() after work.wait())."""
    s = stats
    vals = [float(getattr(s, k)) for k in _SUMS] + [float(x) for x in s.scores] + [float(x) for x in s.resign_scores]
    vec = torch.tensor(vals, dtype=torch.float64, device=_device())
    work = dist.all_reduce(vec, op=dist.ReduceOp.SUM, async_op=async_op)

    def finish():
        v = dict(zip(_SUMS, vec.tolist()))
        sc = vec.tolist()[len(_SUMS):]
        games, full, fast, total, length = (v["games_completed"], v["full_move_count"], v["fast_move_count"],
                                            v["total_move_count"], v["sum_game_length"])
        out = {k: int(v[k]) for k in ("simulations", "moves", "games_completed", "games_started", "active_games", "hist_count",
                                      "cache_hits", "cache_misses")}
        out.update(scores=sc[:3], resign_scores=sc[3:6],
                   avg_game_length=length / games if games else 0.0,
                   avg_leaf_depth=v["sum_leaf_depth"] / full if full else 0.0,
                   avg_search_entropy=v["sum_search_entropy"] / full if full else 0.0,
                   fast_avg_leaf_depth=v["fast_sum_leaf_depth"] / fast if fast else 0.0,
                   fast_avg_search_entropy=v["fast_sum_search_entropy"] / fast if fast else 0.0,
                   avg_moves_per_turn=total / length if length else 0.0,
                   avg_valid_moves=v["sum_valid_moves"] / total if total else 0.0,
                   nccl_bytes=vec.numel() * 8)
        return out

    return (work, finish) if async_op else finish()


def broadcast_weights(module_or_tensors, src=0, async_op=False):
    """Weight broadcast after a training step: every parameter and buffer of the torch net — or a plain
    list of tensors — goes from rank `src` to all ranks in ONE flat bucket (one NCCL launch; the nets here are a few
    hundred KB, so launch latency, not bandwidth, is what matters over NVLink). Returns the bytes broadcast (or
    (work, finish) with async_op: finish() copies the bucket back into the tensors)."""
    if hasattr(module_or_tensors, "state_dict"):
        tensors = [t for t in module_or_tensors.state_dict().values() if torch.is_tensor(t) and t.numel() > 0]
    else:
        tensors = list(module_or_tensors)
    if not tensors:
        return 0
    dev = _device()
    flat = torch.cat([t.detach().reshape(-1).to(device=dev, dtype=torch.float32) for t in tensors])
    work = dist.broadcast(flat, src=src, async_op=async_op)

    def finish():
        off = 0
        with torch.no_grad():
            for t in tensors:
                n = t.numel()
                t.copy_(flat[off:off + n].reshape(t.shape).to(dtype=t.dtype, device=t.device))
                off += n
        return flat.numel() * 4

    return (work, finish) if async_op else finish()


class GlobalBudget:
    """ONE games_to_play budget for all ranks, like the reference's games_started_ counter (play_manager.cc:506-513): every
    rank starts with the whole budget as its own target (no slot retires early), and every `sync()` all-reduces the
    number of games started so far; once the global count reaches the budget each rank lowers its engine's target to what
    it has started itself, so its slots retire as their current games end. The overshoot is at most the games started
    between two sync() calls, as with the reference's check-then-increment under a mutex it is zero."""

    def __init__(self, engine, total_games):
        self.engine, self.total, self.closed = engine, int(total_games), False
        engine.set_games_to_play(min(self.total, 0xFFFFFFFF))

    def sync(self, stats=None):
        s = stats or self.engine.stats()
        t = torch.tensor([float(s.games_started), float(s.games_completed), float(s.active_games)], dtype=torch.float64,
                         device=_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        started, completed, active = (int(x) for x in t.tolist())
        if not self.closed and started >= self.total:
            self.engine.set_games_to_play(max(int(s.games_started), 1))  # no slot of this rank starts another game
            self.closed = True
        return {"games_started": started, "games_completed": completed, "active_games": active, "closed": self.closed}


def gather_history(canon, v, pi, dst=0, async_op=False):
    """Gather every rank's finished training samples (numpy or torch arrays, first dim = samples) on rank `dst`
    (others get None). Variable counts per rank: sizes are exchanged first, then a padded all_gather per array.
    With async_op the three gathers are only LAUNCHED: returns (works, finish) and finish() assembles the result
    after the works have been waited for — a run gathers the samples of generation k while generation k + 1 plays."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device()
    t = [torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(dev) for a in (canon, v, pi)]
    n = torch.tensor([t[0].shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    cap = max(sizes) if sizes else 0
    works, parts_all, nbytes = [], [], 0
    for a in t:
        pad = torch.zeros((cap,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev)
        pad[: a.shape[0]] = a
        parts = [torch.empty_like(pad) for _ in range(world)]
        works.append(dist.all_gather(parts, pad, async_op=async_op))
        parts_all.append(parts)
        nbytes += pad.numel() * pad.element_size() * world

    def finish():
        if rank != dst:
            return None
        return tuple(torch.cat([p[:k] for p, k in zip(parts, sizes)], 0) for parts in parts_all)

    gather_history.last_nccl_bytes = nbytes
    return (works, finish) if async_op else finish()


def max_over_ranks(value):
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
