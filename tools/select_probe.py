#!/usr/bin/env python
"""Times tf_select_set (the elite / victim selection of an island exchange) and tf_rank_topk on the fitness vector of a
cfg3 population after 20 generations (clustered values: the case that matters).  Diagnostics only."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch
from tracefly_b200 import _lib as L, batched as B
from tracefly_b200.islands import IslandGA

torch.cuda.set_device(0)
cfg = dict(bench.WORKLOADS["cfg3"])
env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
isl = IslandGA(pop, ga, seed=1, elites=1024, interval=10, env=env, params=params)
for _ in range(20):
    isl.step()
torch.cuda.synchronize()
out = {}
for name, k, worst in (("best_1024", 1024, False), ("worst_1024", 1024, True), ("worst_7168", 7168, True)):
    sel = B.SelectSet(isl.P, k)
    idx = torch.empty(k, dtype=torch.int32, device="cuda")
    ts = []
    for rep in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sel(isl.fit, worst=worst, out=idx); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    out[name] = {"median_us": float(np.median(ts[5:])), "min_us": float(np.min(ts[5:]))}
top = B.TopK(isl.P, 2)
t2 = torch.empty(2, dtype=torch.int32, device="cuda")
ts = []
for rep in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); top(isl.fit, out=t2); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
out["rank_topk_2"] = {"median_us": float(np.median(ts[5:])), "min_us": float(np.min(ts[5:]))}
out["distinct_top11_bins"] = int(torch.unique((isl.fit.view(torch.int32) >> 21)).numel())
print(json.dumps(out))
