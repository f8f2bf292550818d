#!/usr/bin/env python
"""cfg4 in the converged regime bench.py times: where the moving-sphere time goes (diagnostics only).
Evolves the cfg4 population for N generations, then reports how many waypoints fall outside the candidate
index, and the fitness-sweep time / equality for several index geometries against the brute-force loop."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from tracefly_b200 import _lib as L, batched as B
from tracefly_b200.islands import IslandGA
from tools.microbench import timeit

cfg = dict(bench.WORKLOADS["cfg4"]); P, W = cfg["P"], cfg["W"]
env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
isl = IslandGA(pop, ga, seed=bench.SEED, env=env, params=params)
for _ in range(int(os.environ.get("GENS", 60))):
    isl.step()
torch.cuda.synchronize()
pop = isl.pop
wx, wy, wz = cfg["world"]
xyz = pop[..., :3]
outside = ((xyz < 0) | (xyz >= torch.tensor([wx, wy, wz], device="cuda"))).any(-1)
print("waypoints outside the world box: %.4f %%" % (100.0 * outside.float().mean().item()))
print("chromosomes with an outside waypoint: %.4f %%" % (100.0 * outside.any(-1).float().mean().item()))
uniq = torch.unique(pop.view(P, -1), dim=0).shape[0] if P <= (1 << 20) else -1
print("distinct chromosomes:", uniq, "of", P)
fit = torch.empty(P, device="cuda")
ref = None
for cell, cap, band in ((32, 48, 16), (24, 40, 16), (16, 32, 16), (0, 0, 0)):
    if cell:
        env.build_dynamic_index(W, (0, 0, 0), cfg["world"], cell=float(cell), cap=cap, band=band)
        st = env.dynamic_index_stats()
    else:
        env.set_dynamic(bench.dyn_table(cfg, cfg["n_dyn"], bench.SEED), 2.0)     # drops the index: brute force
        st = "brute force"
    t = timeit(torch, lambda: B.fitness_batch(pop, None, env, params, out=fit, want_collisions=False), n=3, warm=1)
    f, c = B.fitness_batch(pop, None, env, params)
    if ref is None:
        ref = (f.clone(), c.clone())
    same = bool(torch.equal(f, ref[0]) and torch.equal(c, ref[1]))
    print(f"{t[0]:8.3f} ms  cell {cell} cap {cap} band {band}  {st}  identical to first: {same}")
