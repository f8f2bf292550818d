#!/usr/bin/env python
"""Small invocation of every hot kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tracefly_b200 import _lib as L, batched as B

rng = np.random.default_rng(0)
dims = (64, 64, 32)
table = np.zeros((300, 8)); table[:, :3] = rng.integers(0, 64, size=(300, 3)); table[:, 2] %= 32
table[:, 3] = rng.integers(1, 4, size=300); table[:, 7] = rng.integers(0, 4, size=300); table[:, 4:7] = 2
for W in (64, 32, 20):
    P = 1000
    pop = B.init_walks(1, P, W, start=(32, 32, 16), lo=(0, 0, 0), hi=(63, 63, 31))
    envs = [B.Environment(L.SEM_PF, table, 0.0), B.Environment(L.SEM_HR, table, 2.0),
            B.Environment.empty().set_voxels(B.voxelize(table, dims, L.SEM_PF), dims),
            B.Environment(L.SEM_PF, table, 0.0).build_static_index((0, 0, 0), dims, 8.0, 96, 1.0, 32)]
    dyn = np.zeros((50, 8), np.float32); dyn[:, :3] = rng.uniform(0, 32, size=(50, 3)); dyn[:, 3] = 2; dyn[:, 4:7] = 0.3
    e = B.Environment.empty().set_dynamic(dyn, 1.0).build_dynamic_index(W, (0, 0, 0), dims, 16.0, cap=64, band=16)
    e.set_hazard(torch.rand((8, 16, 16), device="cuda"), 0.25, 0.25, 0, 0, 7.0 / max(W - 1, 1))
    envs.append(e)
    for env in envs:
        prm = L.default_params(hazard_weight=0.2)
        fit, coll = B.fitness_batch(pop, None, env, prm)
        top = B.rank_topk(fit, 10)
        for sel in (L.SEL_TOURNAMENT, L.SEL_TRUNCATION):
            ga = L.default_ga_params(selection=sel, mutation_rate=0.5)
            B.generation(env, pop, fit, top, ga, prm, seed=1, generation=2, want_collisions=True, want_parity=True)
            B.ga_step(pop, fit, top, ga, seed=1, generation=2)
    B.SelectSet(P, 100)(fit, worst=True); B.rank_topk(fit, 200)
torch.cuda.synchronize()
print("sanitize case done")
