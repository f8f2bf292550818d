import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tracefly_b200 import _lib as L, batched as B
from tools.microbench import timeit
from tests.helpers import random_table
rng = np.random.default_rng(1)
for sem, name in ((L.SEM_HR, "HR spheres N=10"), (L.SEM_PF, "PF shapes N=150")):
    table = random_table(rng, 10 if sem == L.SEM_HR else 150, 64, sem_pf=(sem == L.SEM_PF))
    env = B.Environment(sem, table, 2.0 if sem == L.SEM_HR else 0.0)
    for W in (20, 48):
        P = (1 << 25) // W
        pop = B.init_walks(7, P, W, start=(32, 32, 32), lo=(0, 0, 0), hi=(63, 63, 63))
        prm = L.default_params(dist_zero_guard=1 if sem == L.SEM_HR else 0)
        ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=4, mutation_rate=0.1)
        fit, _ = B.fitness_batch(pop, None, env, prm, want_collisions=False)
        top = B.rank_topk(fit, 2)
        nxt, nfit = torch.empty_like(pop), torch.empty_like(fit)
        out = []
        for dis in (0, 1):
            L.set_option(L.OPT_DISABLE_STREAM, dis)
            out.append(timeit(torch, lambda: B.generation(env, pop, fit, top, ga, prm, seed=1, generation=1, out_pop=nxt, out_fit=nfit), n=8)[0])
        L.set_option(L.OPT_DISABLE_STREAM, 0)
        print(f"{name} W={W} P={P}: TMA-staged {out[0]:.3f} ms, LSU {out[1]:.3f} ms")
