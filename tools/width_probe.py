#!/usr/bin/env python
"""Fused generation / fitness sweep at other chromosome lengths (1 GiB populations, cfg3 voxel world): W = 32 / 64 run
the TMA-staged kernels, every other W the LSU-path kernels.  Diagnostics only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tracefly_b200 import _lib as L, batched as B
from tools.microbench import timeit

cfg = dict(bench.WORKLOADS["cfg3"])
env, params, _, ga = bench.build_gpu_workload(dict(cfg, P=1024), 0, torch, B, L)
for W in (20, 32, 48, 64, 128, 256):
    P = (1 << 26) // W
    starts = (torch.rand((P, 4), device="cuda") * 1024).floor()
    pop = B.init_walks(7, P, W, lo=(0, 0, 0), hi=(1023, 1023, 1023), start_per_path=starts)
    fit, _ = B.fitness_batch(pop, None, env, params, want_collisions=False)
    top = B.rank_topk(fit, 2)
    nxt, nfit = torch.empty_like(pop), torch.empty_like(fit)
    tf = timeit(torch, lambda: B.fitness_batch(pop, None, env, params, out=fit, want_collisions=False), n=8)
    tg = timeit(torch, lambda: B.generation(env, pop, fit, top, ga, params, seed=1, generation=1, out_pop=nxt, out_fit=nfit), n=8)
    gb = 2 * P * W * 16 / 1e9
    print(f"W {W:4d} P {P:8d}: fitness {tf[0]:.3f} ms ({gb / 2 / tf[0] * 1e3:.0f} GB/s read), generation {tg[0]:.3f} ms ({gb / tg[0] * 1e3:.0f} GB/s read+write)")
    del pop, nxt, fit, nfit
