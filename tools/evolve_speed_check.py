import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tracefly_b200 import _lib as L, batched as B
cfg = dict(bench.WORKLOADS["cfg3"])
env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
for gens in (5, 200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = B.evolve_batch(pop.clone(), env, generations=gens, ga=ga, params=params, seed=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(gens, "generations:", round(dt * 1e3, 2), "ms ->", round(dt * 1e3 / gens, 4), "ms/gen", type(out).__name__)
