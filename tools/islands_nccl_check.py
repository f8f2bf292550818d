#!/usr/bin/env python
"""Island model on real GPUs over NCCL against the single-process CPU restatement (tests/test_islands_gloo.py's
`_simulate`): every rank evolves its island with the product path (CudaOps + all_gather_into_tensor) and compares
its final population / fitness, bit for bit, with the restatement of the whole world.  Diagnostics; run as
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/islands_nccl_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tests.test_islands_gloo as T  # noqa: E402  (the restatement lives with the gloo test)
from tracefly_b200 import _lib as L, batched as B  # noqa: E402
from tracefly_b200.islands import IslandGA  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
T.P, T.W, T.E, T.INTERVAL, T.GENS = 4096, 32, 64, 3, 11
rng = np.random.default_rng(5)                       # the same draws as T._world
pos, size, kind = rng.integers(0, 32, size=(12, 3)), rng.integers(1, 4, size=12), rng.integers(0, 3, size=12)
table = np.zeros((12, 8))
table[:, :3], table[:, 3], table[:, 7] = pos, size, kind
ok = True
for pipelined in (True, False):
    O, oenv, pop_h = T._world(rank)
    env = B.Environment(L.SEM_PF, table, 0.0)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.3)
    isl = IslandGA(torch.from_numpy(pop_h).cuda(), ga, seed=T.SEED, elites=T.E, interval=T.INTERVAL, env=env,
                   params=L.default_params(dist_zero_guard=1), pipelined=pipelined)
    isl.warm_up()               # throw-away pipelined exchange beside a generation kernel: state must not change
    for g in range(T.GENS):
        isl.step()
        if g == 1:
            isl.warm_up()
    isl.flush()
    torch.cuda.synchronize()
    ref = T._simulate(world, pipelined)[rank]
    same = np.array_equal(isl.pop.cpu().numpy(), ref["pop"]) and np.array_equal(isl.fit.cpu().numpy(), ref["fit"])
    print(f"rank {rank}/{world} pipelined={pipelined}: migrations {isl.migrations}, identical to the restatement: {same}", flush=True)
    ok = ok and same and isl.migrations == T.GENS // T.INTERVAL
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
