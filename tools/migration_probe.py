#!/usr/bin/env python
"""Times the pieces of one island migration on N GPUs (torchrun).  Diagnostics only."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch, torch.distributed as dist
from tracefly_b200 import _lib as L, batched as B
from tracefly_b200.islands import IslandGA

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = dict(bench.WORKLOADS["cfg3"])
env, params, pop, ga = bench.build_gpu_workload(cfg, rank, torch, B, L)
isl = IslandGA(pop, ga, seed=1, elites=1024, interval=10, env=env, params=params)
for _ in range(10): isl.step()
torch.cuda.synchronize()
def ev(): return torch.cuda.Event(enable_timing=True)
res = {}
for rep in range(5):
    marks = []
    def mark(name):
        e = ev(); e.record(); marks.append((name, e, time.perf_counter()))
    mark("start")
    E, W = isl.E, isl.W
    if world > 1:
        isl.ops.elites(isl.fit, isl.elite_idx); mark("topk_elites")
        isl.ops.pack(isl.pop, isl.fit, isl.elite_idx, isl.send_rows, isl.send_fit); mark("pack")
        dist.all_gather_into_tensor(isl.recv_rows, isl.send_rows); dist.all_gather_into_tensor(isl.recv_fit, isl.send_fit); mark("all_gather")
        isl.ops.worst(isl.fit, isl.worst_idx); mark("select_worst_set")
        isl.ops.replace(isl.pop, isl.fit, isl.worst_idx, isl.recv_rows, isl.recv_fit, world, rank, E); mark("replace")
    torch.cuda.synchronize()
    for (n0, e0, t0), (n1, e1, t1) in zip(marks[:-1], marks[1:]):
        res.setdefault(n1, []).append((e0.elapsed_time(e1) * 1e3, (t1 - t0) * 1e6))
t0 = time.perf_counter(); isl.migrate() if world > 1 else None; torch.cuda.synchronize(); t_mig = (time.perf_counter() - t0) * 1e6
if rank == 0:
    print(json.dumps({k: {"gpu_us": float(np.median([a for a, b in v])), "host_us": float(np.median([b for a, b in v]))} for k, v in res.items()}, indent=1))
    print("migrate() wall us", t_mig)
if world > 1: dist.destroy_process_group()
