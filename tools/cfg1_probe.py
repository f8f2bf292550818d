#!/usr/bin/env python
"""BASELINE configs[0] through the drop-in HRMAGAGenerator: wall time vs the reference's 48 s."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tracefly_b200 import path_planner_hrmaga as hr, _lib as L
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "evolve_hr_cfg1.npz"))
meta = json.loads(str(g["meta"]))
obs = [{"position": [int(v) for v in p], "size": int(s)} for p, s in zip(g["obs_pos"], g["obs_size"])]
for rep in range(3):
    np.random.seed(meta["seed"])
    n0 = L.launch_count()
    t0 = time.perf_counter()
    best = hr.HRMAGAGenerator(meta["P"], meta["rate"], meta["G"]).evolve(tuple(meta["start"]), tuple(meta["end"]), obs)
    dt = time.perf_counter() - t0
    ok = [tuple(map(float, w)) for w in best] == [tuple(w) for w in g["best"].tolist()]
    units = (meta["G"] + 1) * meta["P"] * (meta["W"] - 1)
    print(json.dumps({"rep": rep, "seconds": dt, "same_path_as_reference": ok, "units_per_s": units / dt,
                      "generations_per_s": meta["G"] / dt, "launches": L.launch_count() - n0,
                      "reference_seconds": meta["ref_seconds"], "speedup": meta["ref_seconds"] / dt}))
