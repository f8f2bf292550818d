#!/usr/bin/env python
"""cfg4 (voxels + hazard map + moving spheres): fused-generation time against the candidate-index geometry, in the
regime the driver's bench run times (GENS generations after the start, default 6), plus the component split.
    python tools/cfg4_sweep.py            (TRACEFLY_B200_LIB=<variant .so> for A/B builds)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from tracefly_b200 import _lib as L, batched as B  # noqa: E402
from tracefly_b200.islands import IslandGA  # noqa: E402
from tools.microbench import timeit  # noqa: E402

cfg = dict(bench.WORKLOADS["cfg4"])
P, W = cfg["P"], cfg["W"]
wx, wy, wz = cfg["world"]
env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
isl = IslandGA(pop, ga, seed=bench.SEED, env=env, params=params)
for _ in range(int(os.environ.get("GENS", 6))):
    isl.step()
torch.cuda.synchronize()
pop, fit = isl.pop, isl.fit
top = B.rank_topk(fit, 2)
nxt, nfit = isl.nxt, isl.nfit
out = {"lib": os.environ.get("TRACEFLY_B200_LIB", "product"), "gens_before": isl.gen}
bytes_gen = (32 * W + 12) * P
peak = bench.measured_peak_gbs()[0]


def gen_ms(e, prm):
    return timeit(torch, lambda: B.generation(e, pop, fit, top, ga, prm, seed=1, generation=99, out_pop=nxt, out_fit=nfit), n=10)[0]


m = 64.0
ref = None
GEOMS = [tuple(int(v) for v in g.split(",")) for g in os.environ.get("GEOMS", "32,48,16;16,23,16;12,15,16;16,23,8;8,15,16").split(";")]
for cell, cap, band in GEOMS:
    try:
        env.build_dynamic_index(W, (-m, -m, -m), (wx + m, wy + m, wz + m), cell=float(cell), cap=cap, band=band)
    except Exception as ex:   # too many lists for this geometry
        out[f"cell{cell}_cap{cap}_band{band}"] = str(ex)
        continue
    st = env.dynamic_index_stats()
    t = gen_ms(env, params)
    f2, c2 = B.fitness_batch(pop, None, env, params)
    if ref is None:
        ref = (f2.clone(), c2.clone())
    same = bool(torch.equal(f2, ref[0]) and torch.equal(c2, ref[1]))
    out[f"cell{cell}_cap{cap}_band{band}"] = dict(ms=t, frac=bytes_gen / (t * 1e-3) / 1e9 / peak, stats=st, identical_to_first=same)
# components (with the default index geometry of bench.py)
env.build_dynamic_index(W, (-m, -m, -m), (wx + m, wy + m, wz + m), cell=float(os.environ.get("TF_DYN_CELL", 8)),
                        cap=int(os.environ.get("TF_DYN_CAP", 15)), band=16)
out["all"] = gen_ms(env, params)
env.set_hazard(None)
out["vox+dyn"] = gen_ms(env, L.default_params())
env.set_dynamic(np.zeros((0, 8), np.float32))
out["vox"] = gen_ms(env, L.default_params())
env.set_hazard(torch.from_numpy(bench.hazard_field()).cuda(), 0.25, 0.25, 0.0, 0.0, 23.0 / (W - 1))
out["vox+haz"] = gen_ms(env, L.default_params(hazard_weight=0.5))
print(json.dumps(out, indent=1))
