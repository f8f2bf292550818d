#!/usr/bin/env python
"""cfg4 component timings: which of {voxels, hazard, moving spheres} costs what."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from tracefly_b200 import _lib as L, batched as B
from tools.microbench import timeit
cfg = dict(bench.WORKLOADS["cfg4"]); P, W = cfg["P"], cfg["W"]
env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
print("index", env.dynamic_index_stats())
fit = torch.empty(P, device="cuda")
out = {}
out["all"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, params, out=fit, want_collisions=False), n=5)
env.set_hazard(None)
p2 = L.default_params()
out["vox+dyn"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, p2, out=fit, want_collisions=False), n=5)
for cell, cap, band in ((64, 96, 16), (128, 256, 64), (32, 48, 16), (48, 64, 16)):
    env.build_dynamic_index(W, (0, 0, 0), cfg["world"], cell=float(cell), cap=cap, band=band)
    st = env.dynamic_index_stats()
    out[f"vox+dyn cell{cell} cap{cap} band{band} {st}"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, p2, out=fit, want_collisions=False), n=5)
env.set_dynamic(np.zeros((0, 8), np.float32))
out["vox only"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, p2, out=fit, want_collisions=False), n=5)
env.set_hazard(torch.from_numpy(bench.hazard_field()).cuda(), 0.25, 0.25, 0.0, 0.0, 23.0 / (W - 1))
p3 = L.default_params(hazard_weight=0.5)
out["vox+haz"] = timeit(torch, lambda: B.fitness_batch(pop, None, env, p3, out=fit, want_collisions=False), n=5)
for k, v in out.items(): print(f"{v[0]:8.3f} ms  {k}")
