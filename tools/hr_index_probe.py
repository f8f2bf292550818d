#!/usr/bin/env python
"""BASELINE configs[1] shape with HR sphere rules: brute force vs candidate index."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tracefly_b200 import _lib as L, batched as B
from tools.microbench import timeit
rng = np.random.default_rng(0)
P, W, n = 65536, 32, 1024
world = (512, 512, 128)
obstacles = [{"position": [int(rng.integers(0, world[a])) for a in range(3)], "size": int(rng.integers(1, 5))} for _ in range(n)]
pop = B.init_walks(1, P, W, start=(256, 256, 64), lo=(0, 0, 0), hi=(511, 511, 127))
fit = torch.empty(P, device="cuda")
plain = B.Environment.from_hr_obstacles(obstacles, 2.0)
print("plain HR 1024 spheres  %.3f ms" % timeit(torch, lambda: B.fitness_batch(pop, None, plain, out=fit, want_collisions=False))[0])
for cell in (16.0, 24.0, 32.0):
    fast = B.Environment.from_hr_obstacles_indexed(obstacles, (0, 0, 0), world, cell=cell, cap=96)
    t = timeit(torch, lambda: B.fitness_batch(pop, None, fast, out=fit, want_collisions=False))[0]
    print("indexed cell %4.0f        %.3f ms" % (cell, t), fast.dynamic_index_stats())
# spread-out population (paths all over the world)
starts = torch.rand((P, 4), device="cuda") * torch.tensor([512, 512, 128, 1.0], device="cuda")
pop2 = B.init_walks(2, P, W, lo=(0, 0, 0), hi=(511, 511, 127), start_per_path=starts.floor())
print("spread population: plain %.3f ms" % timeit(torch, lambda: B.fitness_batch(pop2, None, plain, out=fit, want_collisions=False))[0])
fast = B.Environment.from_hr_obstacles_indexed(obstacles, (0, 0, 0), world, cell=16.0, cap=96)
print("spread population: indexed %.3f ms" % timeit(torch, lambda: B.fitness_batch(pop2, None, fast, out=fit, want_collisions=False))[0])
