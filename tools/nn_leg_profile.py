"""Where a generation of the NN-evaluator loop spends its time: step kernel vs canonicalisation vs the torch net, for
several kernels / cache-hit caps. Usage (GPU box): python tools/nn_leg_profile.py  (env B2AZ_STEP_KERNEL, B2AZ_HIT_CAP)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-pybind11_b200"))
import torch  # noqa: E402
import b2az  # noqa: E402

G, SIMS = 65536, 400
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
net = C4Net().cuda().eval().to(memory_format=torch.channels_last)
p = b2az.default_params(games_to_play=2 ** 31 - 1, concurrent_games=G, mcts_visits=(SIMS, SIMS), cpuct=1.25,
                        fpu_reduction=0.25, epsilon=0.25, mcts_root_temp=1.25, start_temp=1.0, final_temp=0.2,
                        temp_decay_half_life=10.0, root_fpu_zero=1, shaped_dirichlet=1, policy_target_pruning=1,
                        eval_type=b2az.EVAL_NN, rng_mode=b2az.RNG_PER_GAME, seed=3000, tree_reuse=1,
                        history_enabled=1, self_play=1, max_cache_size=int(os.environ.get("CACHE", "200000")),
                        history_capacity=8 * G)
eng = b2az.Engine(p)
stream = torch.cuda.current_stream().cuda_stream


class _View:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 3}


xin = torch.zeros((G, 4, 6, 7), device="cuda").contiguous(memory_format=torch.channels_last)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        net(xin)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        v, pi = net(xin)
        v_s, pi_s = v.float().contiguous(), pi.float().contiguous()
x_all = None
T = {"step": 0.0, "canon": 0.0, "net": 0.0}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
gens = 0
for it in range(300):
    ev[0].record()
    eng.step(1, stream)
    ev[1].record()
    n, cptr, iptr = eng.leaf_batch(stream)
    if x_all is None:
        x_all = torch.as_tensor(_View(cptr, (G, 4, 6, 7)), device="cuda")
    ev[2].record()
    xin[:n].copy_(x_all[:n])
    gr.replay()
    ev[3].record()
    eng.submit_eval(v_s.data_ptr(), pi_s.data_ptr(), n)
    torch.cuda.synchronize()
    if it == 99:
        s0 = eng.stats(stream)
    if it >= 100:
        T["step"] += ev[0].elapsed_time(ev[1])
        T["canon"] += ev[1].elapsed_time(ev[2])
        T["net"] += ev[2].elapsed_time(ev[3])
        gens += 1
s1 = eng.stats(stream)
sims = s1.simulations - s0.simulations
tot = sum(T.values())
print(json.dumps({"kernel": os.environ.get("B2AZ_STEP_KERNEL", "default"), "hit_cap": os.environ.get("B2AZ_HIT_CAP", "64"),
                  "cache": os.environ.get("CACHE", "200000"),
                  "ms_per_gen": {k: round(v / gens, 3) for k, v in T.items()}, "sims_per_gen": sims / gens,
                  "Msims_per_s": round(sims / tot / 1e3, 1),
                  "hit_rate": round((s1.cache_hits - s0.cache_hits) / max(1, s1.cache_hits - s0.cache_hits + s1.cache_misses - s0.cache_misses), 3)}))
eng.close()
