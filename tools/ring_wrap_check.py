"""Soak check of the per-device launch-counter ring (sched_slot): more dynamically scheduled launches than the ring has
slots (4096), results compared with the first round and with the LSU-path kernels.  Diagnostics only."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracefly_b200 import _lib as L, batched as B
P, W = 700001, 64
pop = B.init_walks(3, P, W, start=(512, 512, 512), lo=(0, 0, 0), hi=(1023, 1023, 1023))
env = B.Environment.empty()
prm = L.default_params()
fit, _ = B.fitness_batch(pop, None, env, prm)
tk = B.TopK(P, 2)
idx = torch.empty(2, dtype=torch.int32, device="cuda")
ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=4, mutation_rate=0.2)
ref = None
for rnd in range(3):
    for i in range(2200):            # 2200 ranking launches + 2200 generation launches per round: the 4096-slot ring wraps
        tk(fit, out=idx)
        if i % 50 == 0:
            out = B.generation(env, pop, fit, idx, ga, prm, seed=9, generation=4)
    out = B.generation(env, pop, fit, idx, ga, prm, seed=9, generation=4)
    torch.cuda.synchronize()
    cur = (idx.clone(), out[0].clone(), out[1].clone())
    if ref is None:
        L.set_option(L.OPT_DISABLE_STREAM, 1)
        o2 = B.generation(env, pop, fit, B.rank_topk(fit, 2), ga, prm, seed=9, generation=4)
        L.set_option(L.OPT_DISABLE_STREAM, 0)
        assert torch.equal(o2[0], out[0]) and torch.equal(o2[1], out[1])
        ref = cur
    assert all(torch.equal(a, b) for a, b in zip(ref, cur)), rnd
    print("round", rnd, "ok; launches so far", L.launch_count())
