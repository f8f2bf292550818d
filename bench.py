#!/usr/bin/env python
"""bench.py — waypoint-segment fitness evaluations/s (+ GA generations/s) of the TraceFly GA hot
path on N B200s, next to the reference's own Python GA and the C port timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg4|cfg5] [--impl reference]

A "step" is one full GA generation over one island: rank (tf_rank_topk) -> selection + crossover
+ mutation + fitness sweep (fused tf_generation).  With N > 1 (torchrun, one process per GPU) each
rank owns one island of the same size (weak scaling) and the islands exchange elites with an NCCL
all-gather every 10 generations.  Unit: one waypoint-segment evaluation: a
generation over P chromosomes of W waypoints is P*(W-1) units.

Workloads (BASELINE.json configs; synthetic, seeded, generated on the device):
  cfg3 (default, configs[2], "the 1M-path config" the north_star target is quoted on):
        P = 1,048,576 x W = 64 vs a 1024^3 bit-packed voxel grid (128 MiB, rasterised from 16,384
        seeded spheres / cubes / cylinders), paths start uniformly over the world.
  cfg2  (configs[1]) P = 65,536 x W = 32 vs 1,024 analytic sphere / cube / cylinder obstacles.
  cfg4  (configs[3]) cfg3 + hazard heat-map f32[24,256,256] + 1,024 moving spheres.
  cfg5  (configs[4]) cfg3's world with 8,388,608 chromosomes per GPU (64M over 8 islands).
  cfg1  (configs[0]) the reference's CPU scenario — HRMAGAGenerator(100, 0.1, 200) on 20 waypoints —
        through the drop-in class, beside the unmodified reference run in full.
Inputs are far larger than L2 (population 1 GiB, ping-ponged between two buffers; grid 128 MiB).

The ONE JSON line is the selected workload (default cfg3).  On top of the contract keys it carries
  "workloads": short timed runs of the other BASELINE configs (N = 1: cfg1, cfg2, cfg4, cfg5's per-GPU size; N > 1:
               cfg5), each with its own roofline and cpu_baseline (cfg5 shares cfg3's world: no CPU legs of its own),
  "parity_gate": the correctness gate run with every benchmark (GPU vs oracle on a sample of the timed run's own
               population: fitness and collision flags bit-exact in the FP32 model, parents / crossover points /
               children bit-exact; N > 1: a small island world over NCCL vs the CPU restatement),
  "cpu_baseline": the UNMODIFIED reference (oracle/_ref, copied there by build()) on all host cores, and
  "cpu_baseline_port": the C port of the same generation (oracle/tf_oracle.c), a second, labelled figure.

--impl reference times the reference's own Python code (all host cores, multiprocessing) on a bounded sample of
the same workload; when oracle/_ref is absent it falls back to the C port and says so (kind "port").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "waypoint_segment_fitness_evals_per_sec"
UNIT = "segment-evals/s"

WORKLOADS = {
    # name: P per GPU, W, world dims, analytic obstacles, voxel obstacles, hazard, dynamic
    "cfg3": dict(P=1 << 20, W=64, world=(1024, 1024, 1024), n_analytic=0, n_voxel_obs=16384, hazard=False, n_dyn=0),
    "cfg2": dict(P=1 << 16, W=32, world=(512, 512, 128), n_analytic=1024, n_voxel_obs=0, hazard=False, n_dyn=0),
    "cfg4": dict(P=1 << 20, W=64, world=(1024, 1024, 1024), n_analytic=0, n_voxel_obs=16384, hazard=True, n_dyn=1024),
    "cfg5": dict(P=1 << 23, W=64, world=(1024, 1024, 1024), n_analytic=0, n_voxel_obs=16384, hazard=False, n_dyn=0),
}
SEED = 0x5EED
MIGRATE_EVERY = 10
ELITES = 1024
GA_KW = dict(selection=1, tournament_k=4, crossover=0, mutation=0, mutation_rate=0.1, n_elite=2)


# ----------------------------------------------------------------------------------------- workload
def obstacle_table(cfg, n, max_size, seed):
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 8))
    for a in range(3):
        t[:, a] = rng.integers(0, cfg["world"][a], size=n)
    t[:, 3] = rng.integers(1, max_size + 1, size=n)
    t[:, 7] = rng.integers(0, 3, size=n)   # ~1/3 each sphere / cube / cylinder (PF:20)
    return t


def hazard_field(nt=24, ny=256, nx=256, seed=SEED):
    """Seeded smooth field in [0,1]: sum of 8 Gaussians drifting over t (SURVEY §8d cfg 4)."""
    rng = np.random.default_rng(seed + 4)
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    out = np.zeros((nt, ny, nx), np.float32)
    c0 = rng.uniform(0, 256, size=(8, 2)); v = rng.uniform(-3, 3, size=(8, 2)); s = rng.uniform(12, 40, size=8)
    for t in range(nt):
        for k in range(8):
            cy, cx = c0[k] + v[k] * t
            out[t] += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s[k] ** 2)).astype(np.float32)
    return out / out.max()


def dyn_table(cfg, n, seed):
    rng = np.random.default_rng(seed + 9)
    d = np.zeros((n, 8), np.float32)
    for a in range(3):
        d[:, a] = rng.uniform(0, cfg["world"][a], size=n)
    d[:, 3] = rng.integers(1, 5, size=n)
    d[:, 4:7] = rng.uniform(-1, 1, size=(n, 3)) / np.sqrt(3)
    return d


def algorithmic_bytes_per_unit(W):
    """fused generation per child = read 16W (parent halves) + write 16W + 4 (fitness)
    + 8 (fitness / rank reads)."""
    return (32 * W + 12) / (W - 1)


def workload_label(name, cfg, P):
    return (f"{name}: pop {P:,} x {cfg['W']} waypoints per GPU"
            + (f" vs {cfg['world'][0]}^3 bit-packed voxel grid" if cfg["n_voxel_obs"] else "")
            + (f" vs {cfg['n_analytic']} analytic sphere/cube/cylinder obstacles" if cfg["n_analytic"] else "")
            + (" + hazard heat-map 24x256x256 + 1024 moving spheres" if cfg["hazard"] else ""))


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region through NVML (the same counters
    nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.* prints), every millisecond."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.power = index, [], set(), []
        self.stop_flag = False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.001)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def static_traffic(workload, P, first_gen, last_gen):
    """DRAM bytes per launch of the generation kernel from the committed `ncu --set full` captures
    (profiles/r2_traffic.json).  A STATIC figure: the capture whose generation number is closest to the middle of
    the timed region is reported, labelled with the generation it was taken at — it is not measured in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            caps = json.load(f).get(workload, [])
        caps = [c for c in caps if c.get("P") == P]
        if not caps:
            return None, None
        mid = 0.5 * (first_gen + last_gen)
        c = min(caps, key=lambda c: abs(c["generation"] - mid))
        return c["dram_bytes_per_launch"], (f"static ncu --set full capture at generation {c['generation']} "
                                            f"({c['file']}); timed generations {first_gen}..{last_gen}")
    except Exception:
        return None, None


# ----------------------------------------------------------------------------------------- GPU arm
def build_gpu_workload(cfg, rank, torch, B, L):
    P, W = cfg["P"], cfg["W"]
    wx, wy, wz = cfg["world"]
    params = L.default_params(dist_zero_guard=1)
    if cfg["n_analytic"]:
        env = B.Environment(L.SEM_PF, obstacle_table(cfg, cfg["n_analytic"], 4, SEED + 1), 0.0)
        if os.environ.get("TF_STATIC_INDEX", "1") != "0":   # candidate index: same results, ~20x fewer tests
            env.build_static_index((0, 0, 0), cfg["world"], cell=float(os.environ.get("TF_STATIC_CELL", 8)), cap=96,
                                   z_cell=1.0, z_cap=48)
        params.dist_zero_guard = 0
    else:
        env = B.Environment.empty()
    if cfg["n_voxel_obs"]:
        bricks = B.voxelize(obstacle_table(cfg, cfg["n_voxel_obs"], 10, SEED + 2), cfg["world"], L.SEM_PF)
        env.set_voxels(bricks, cfg["world"])
    if cfg["hazard"]:
        env.set_hazard(torch.from_numpy(hazard_field()).cuda(), 0.25, 0.25, 0.0, 0.0, 23.0 / (W - 1))
        params.hazard_weight = 0.5
    if cfg["n_dyn"]:
        env.set_dynamic(dyn_table(cfg, cfg["n_dyn"], SEED), 2.0)
        # the index box is the world plus a margin: lattice mutations let waypoints drift a few cells past the
        # walls, and a waypoint outside the index falls back to the loop over all 1024 spheres (after 60
        # generations 0.16 % of the waypoints — but 9 % of the chromosomes, and the whole warp waits for them)
        m = float(os.environ.get("TF_DYN_MARGIN", 64))
        # cell 8: a waypoint visits 1.8 spheres on average (cell 32: 5.4); 12 M lists of 32 bytes
        env.build_dynamic_index(W, (-m, -m, -m), (wx + m, wy + m, wz + m), cell=float(os.environ.get("TF_DYN_CELL", 8)),
                                cap=int(os.environ.get("TF_DYN_CAP", 15)), band=int(os.environ.get("TF_DYN_BAND", 16)))
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + 100 + rank)
    if cfg["n_analytic"]:
        pop = B.init_walks(SEED + rank, P, W, start=(wx // 2, wy // 2, wz // 2), lo=(0, 0, 0), hi=(wx - 1, wy - 1, wz - 1))
    else:
        starts = torch.rand((P, 4), device="cuda", generator=g) * torch.tensor([wx, wy, wz, 1.0], device="cuda")
        pop = B.init_walks(SEED + rank, P, W, lo=(0, 0, 0), hi=(wx - 1, wy - 1, wz - 1), start_per_path=starts.floor())
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=4, crossover=L.CX_SINGLE_POINT,
                             mutation=L.MUT_LATTICE, mutation_rate=0.1, n_elite=2)
    return env, params, pop, ga


def _claim_stdout():
    """NCCL / driver libraries may printf to fd 1; keep fd 1 for the ONE JSON line by pointing the
    process's stdout at stderr and returning a private handle to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


class Ctx:
    """torch / distributed handles of this rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from tracefly_b200 import _lib as L
        from tracefly_b200 import batched as B
        from tracefly_b200.islands import IslandGA
        self.torch, self.dist, self.L, self.B, self.IslandGA = torch, dist, L, B, IslandGA
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def time_generations(ctx, name, cfg, steps, warmup):
    """W untimed + exactly K timed generations of one island per rank; CUDA events on the launching stream (the
    generation kernel is bracketed through IslandGA's hook), barrier + synchronize on both sides, max over ranks."""
    torch, L, B = ctx.torch, ctx.L, ctx.B
    P, W = cfg["P"], cfg["W"]
    env, params, pop, ga = build_gpu_workload(cfg, ctx.rank, torch, B, L)
    island = ctx.IslandGA(pop, ga, seed=SEED, elites=ELITES, interval=MIGRATE_EVERY, env=env, params=params)
    del pop
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        island.step()
    if ctx.world > 1:
        island.migrate()  # untimed: NCCL communicator set-up and first-collective connect happen here
        for _ in range(int(os.environ.get("TF_BENCH_WARM_EXCHANGES", 1))):
            island.warm_up()  # ... and the first all-gather that runs BESIDE a generation kernel (throw-away, state unchanged)
    ctx.sync_all()

    sampler = ClockSampler(ctx.local_rank)
    gen_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    counter = {"i": 0}

    def hook(what, isl):
        gen_ev[counter["i"]][0 if what == "generation_begin" else 1].record(stream)

    island.hook = hook
    first_gen = island.gen
    launches0 = L.launch_count()
    ctx.sync_all()
    sampler.start()
    step_ev[0].record(stream)
    for i in range(steps):
        counter["i"] = i
        island.step()
        if i + 1 < steps:
            step_ev[i + 1].record(stream)
    island.flush()
    step_ev[steps].record(stream)
    ctx.sync_all()
    island.hook = None
    launches = L.launch_count() - launches0
    clocks = sampler.finish()
    ms_total = ctx.max_over_ranks(step_ev[0].elapsed_time(step_ev[steps]))
    gen_times = [a.elapsed_time(b) for a, b in gen_ev]
    step_times = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(steps)]
    gen_ms = float(np.mean(gen_times))
    gen_ms_slowest = ctx.max_over_ranks(gen_ms)    # the islands meet at every exchange: the slowest GPU sets the pace
    q = max(1, len(gen_times) // 4)
    units_per_step = ctx.world * P * (W - 1)
    peak, peak_src = measured_peak_gbs()
    bpu = algorithmic_bytes_per_unit(W)
    achieved = bpu * P * (W - 1) / (gen_ms * 1e-3) / 1e9
    traffic, traffic_src = static_traffic(name, P, first_gen, first_gen + steps - 1)
    res = {
        "value": units_per_step * steps / (ms_total * 1e-3), "unit": UNIT, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "generations_per_sec": steps / (ms_total * 1e-3),
        "config": {"workload": workload_label(name, cfg, P),
                   "step": "one GA generation = tf_rank_topk + fused tf_generation (tournament k=4, single-point crossover, "
                           "lattice mutation 0.1, 2 elites, Philox4x32-10)",
                   "population_per_gpu": P, "waypoints": W, "islands": ctx.world,
                   "timed_generations": [first_gen, first_gen + steps - 1],
                   "migration": (f"{ELITES} elites NCCL all-gather every {MIGRATE_EVERY} generations, pipelined with the next "
                                 f"generation on a side stream ({island.migrations} exchanges applied so far)") if ctx.world > 1 else "none (1 island)",
                   "l2": "inputs larger than L2: population 2 x %.0f MiB ping-pong, voxel grid %d MiB" % (
                       P * W * 16 / 2 ** 20, (cfg["world"][0] * cfg["world"][1] * cfg["world"][2] // 8) >> 20 if cfg["n_voxel_obs"] else 0)},
        "roofline": {"bound": "hbm", "kernel": "tf_generation (TMA-staged fused generation kernel: duo_kernel<GA> without an obstacle table, stream_kernel<GA> with one)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bpu * P * (W - 1), "algorithmic_bytes_per_unit": bpu, "kernel_ms": gen_ms,
                     "kernel_ms_first_quarter": float(np.mean(gen_times[:q])), "kernel_ms_last_quarter": float(np.mean(gen_times[-q:])),
                     "kernel_share_of_step": gen_ms / (ms_total / steps), "kernel_ms_slowest_rank": gen_ms_slowest},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if os.environ.get("TF_BENCH_TRACE"):
        res["trace"] = {"step_ms": step_times, "kernel_ms": gen_times}
    return res, island, env, params, ga


# ----------------------------------------------------------------------------------------- parity gate
def oracle_world(workload):
    """The oracle's view of a workload's environment (voxel grid rasterised by the oracle itself)."""
    from oracle import c_oracle as O
    cfg = WORKLOADS[workload]
    W = cfg["W"]
    wx, wy, wz = cfg["world"]
    kw = {}
    sem, obs32 = O.SEM_NONE, None
    if cfg["n_analytic"]:
        t = obstacle_table(cfg, cfg["n_analytic"], 4, SEED + 1)
        obs32 = O.obstacle_table(t[:, :3], t[:, 3], t[:, 7].astype(np.int32))
        sem = O.SEM_PF
    if cfg["n_voxel_obs"]:
        t = obstacle_table(cfg, cfg["n_voxel_obs"], 10, SEED + 2)
        venv = O.OracleEnv(O.SEM_PF, O.obstacle_table(t[:, :3], t[:, 3], t[:, 7].astype(np.int32)), 0.0)
        kw.update(bricks=O.voxelize(venv, wx, wy, wz), dims=cfg["world"])
    if cfg["hazard"]:
        kw.update(hazard=hazard_field(), hazard_map=(0.25, 0.25, 0.0, 0.0, 23.0 / (W - 1)))
    if cfg["n_dyn"]:
        kw.update(dyn=dyn_table(cfg, cfg["n_dyn"], SEED), dyn_safety=2.0)
    env = O.OracleEnv(sem, obs32, 0.0, **kw)
    prm = O.default_params(guard=0 if cfg["n_analytic"] else 1, hazard_weight=0.5 if cfg["hazard"] else 0.0)
    return O, env, prm


def parity_gate(ctx, workload, island, env, params, ga, oracle=None, n=2048):
    """Correctness gate (last row), run on the population the timed run ended with: a strided sample
    of n chromosomes goes through the oracle's FP32 model (fitness + collision flags must match the GPU's bit for
    bit), then the sample evolves one generation as a population of its own on both sides (parents, crossover
    points, children and their fitness bit-exact).  Untimed; the oracle is only the checker."""
    torch, L, B = ctx.torch, ctx.L, ctx.B
    O, oenv, oprm = oracle or oracle_world(workload)
    P = island.P
    island.flush()
    idx = torch.arange(0, P, max(1, P // n), device="cuda")[:n]
    sample = island.pop[idx].contiguous()
    fit_run = island.fit[idx].cpu().numpy()
    fit_g, coll_g = B.fitness_batch(sample, None, env, params)
    top_g = B.rank_topk(fit_g, 2)
    g2 = L.GaParams.from_buffer_copy(ga)
    nxt_g, nfit_g, ncoll_g, sel_g, cp_g = B.generation(env, sample, fit_g, top_g, g2, params, seed=SEED, generation=7,
                                                         want_collisions=True, want_parity=True)
    torch.cuda.synchronize()
    pop_h = sample.cpu().numpy()
    fit_o, coll_o = O.fitness_f32(oenv, pop_h, None, oprm)
    top_o = O.topk(fit_o, 2)
    oga = O.default_ga(stream_id=int(ga.stream_id), **GA_KW)
    nxt_o, _, sel_o, cp_o = O.ga_step(pop_h, fit_o, top_o, oga, seed=SEED, gen=7)
    nfit_o, ncoll_o = O.fitness_f32(oenv, nxt_o, None, oprm)
    ref64 = np.maximum(np.abs(fit_o.astype(np.float64)), 1e-30)
    out = {
        "sample": int(len(idx)), "of": f"the population the timed run ended with (generation {island.gen}), every {max(1, P // n)}-th chromosome",
        "fitness_bit_exact_vs_oracle_fp32": bool(np.array_equal(fit_o, fit_g.cpu().numpy())),
        "fitness_of_timed_run_bit_exact": bool(np.array_equal(fit_o, fit_run)),
        "fitness_max_rel_err": float(np.max(np.abs(fit_g.cpu().numpy().astype(np.float64) - fit_o) / ref64)),
        "collision_flags_bit_exact": bool(np.array_equal(coll_o, coll_g.cpu().numpy().view(np.uint32))),
        "selected_parents_bit_exact": bool(np.array_equal(sel_o, sel_g.cpu().numpy())),
        "crossover_points_bit_exact": bool(np.array_equal(cp_o, cp_g.cpu().numpy())),
        "children_bit_exact": bool(np.array_equal(nxt_o, nxt_g.cpu().numpy())),
        "children_fitness_bit_exact": bool(np.array_equal(nfit_o, nfit_g.cpu().numpy())
                                           and np.array_equal(ncoll_o, ncoll_g.cpu().numpy().view(np.uint32))),
    }
    out["ok"] = all(v for k, v in out.items() if k.endswith("bit_exact"))
    return out


def island_gate(ctx):
    """N > 1: a small island world (4096 x 32 per rank, 11 generations, exchange every 3) evolved by the product
    path over NCCL, compared bit for bit with the single-process CPU restatement of the WHOLE world
    (oracle/islands_sim.py, rank 0 computes it).  Untimed."""
    torch, L, B, dist = ctx.torch, ctx.L, ctx.B, ctx.dist
    from oracle import c_oracle as O
    from oracle import islands_sim
    P, W, E, interval, gens, seed = 4096, 32, 64, 3, 11, 42
    rng = np.random.default_rng(5)
    pos, size, kind = rng.integers(0, 32, size=(12, 3)), rng.integers(1, 4, size=12), rng.integers(0, 3, size=12)
    table = np.zeros((12, 8))
    table[:, :3], table[:, 3], table[:, 7] = pos, size, kind

    def make(r):
        oenv = O.OracleEnv(O.SEM_PF, O.obstacle_table(pos, size, kind), 0.0)
        return oenv, None, O.init_walks(seed + r, P, W, (16, 16, 16), (0, 0, 0), (31, 31, 31))

    env = B.Environment(L.SEM_PF, table, 0.0)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.3)
    isl = ctx.IslandGA(torch.from_numpy(make(ctx.rank)[2]).cuda(), ga, seed=seed, elites=E, interval=interval, env=env,
                       params=L.default_params(dist_zero_guard=1), pipelined=True)
    isl.warm_up()               # (a throw-away exchange leaves the island unchanged: the gate covers that too)
    for _ in range(gens):
        isl.step()
    isl.flush()
    torch.cuda.synchronize()
    pops = [torch.empty_like(isl.pop) for _ in range(ctx.world)]
    fits = [torch.empty_like(isl.fit) for _ in range(ctx.world)]
    dist.all_gather(pops, isl.pop)
    dist.all_gather(fits, isl.fit)
    if ctx.rank != 0:
        return None
    O.set_threads(os.cpu_count() or 1)
    ref = islands_sim.simulate(ctx.world, make, dict(selection=1, tournament_k=3, mutation_rate=0.3), seed, gens, interval, E, True)
    O.set_threads(1)
    same = [bool(np.array_equal(pops[r].cpu().numpy(), ref[r]["pop"]) and np.array_equal(fits[r].cpu().numpy(), ref[r]["fit"]))
            for r in range(ctx.world)]
    checksum = int(sum(int(np.frombuffer(ref[r]["pop"].tobytes(), np.uint32).astype(np.uint64).sum()) for r in range(ctx.world)) & 0xFFFFFFFFFFFF)
    return {"world": ctx.world, "islands": f"{P} x {W} per rank, {gens} generations, {E} elites every {interval}, pipelined",
            "migrations_applied": int(isl.migrations), "populations_bit_exact_vs_cpu_restatement": same,
            "population_checksum_u48": checksum, "ok": all(same) and isl.migrations == gens // interval}


# ----------------------------------------------------------------------------------------- e2e
def pcie_ceiling(ctx, h2d_bytes, d2h_bytes, reps=5):
    """Plain pinned cudaMemcpyAsync of one step's bytes in both directions at once (two streams), all ranks at the
    same time: what the host link gives with no kernel in between — the ceiling of the e2e figure."""
    torch = ctx.torch
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def go():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)

    go()
    ctx.sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3 / reps)
    del hin, hout, din, dout
    return {"ms_per_step": ms, "h2d_GBs_per_gpu": h2d_bytes / ms / 1e6, "d2h_GBs_per_gpu": d2h_bytes / ms / 1e6,
            "what": "pinned cudaMemcpyAsync H2D + D2H of one step's bytes, both directions at once, every rank at the same time"}


def run_e2e(ctx, island, env, params, ga, fmt, steps):
    """The public host-array API (batched.HostStepper): every step copies that step's population from pinned host
    memory to the GPU, runs tf_fitness + tf_rank_topk + tf_generation, and copies the fitness vector and the next
    population back; `depth` requests are in flight so the D2H leg of step n overlaps the H2D leg of step n+1."""
    torch, B = ctx.torch, ctx.B
    P, W = island.P, island.W
    depth = int(os.environ.get("TF_E2E_DEPTH", 3))   # 3 requests in flight: 2.5 % over 2 (18.1 -> 17.7 ms per step), 4 adds nothing
    stepper = B.HostStepper(env, P, W, ga, params, fmt=fmt, depth=depth)
    hdt, hc = B.HostStepper._HOST[fmt]
    hosts = [torch.empty((P, W, hc), dtype=hdt).pin_memory() for _ in range(depth)]
    for h in hosts:
        h.copy_(island.pop[..., :hc].to(hdt))
    ctx.sync_all()
    tickets, checksum, t_start = [], 0.0, None
    for it in range(steps + depth):      # `depth` untimed warm-up requests
        if it == depth:
            while tickets:
                stepper.result(tickets.pop(0))
            ctx.sync_all()
            t_start = time.perf_counter()
        if len(tickets) == depth:
            fit_h, next_h = stepper.result(tickets.pop(0))
            checksum += float(fit_h[0])      # the caller really reads the result
        tickets.append(stepper.submit(hosts[it % depth], seed=SEED, generation=1000 + it))
    while tickets:
        fit_h, next_h = stepper.result(tickets.pop(0))
        checksum += float(fit_h[0])
    torch.cuda.synchronize()
    ms = ctx.max_over_ranks((time.perf_counter() - t_start) * 1e3)
    h2d, d2h = stepper.bytes_per_step()
    del stepper, hosts
    torch.cuda.empty_cache()
    return {"value": ctx.world * P * (W - 1) * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": ms / steps, "steps": steps,
            "host_format": {"f32": "float32[P,W,4] (the device layout)", "xyz": "float32[P,W,3] (the reference's 3-tuples as an array)",
                            "i16": "int16[P,W,4] (lattice chromosomes)"}[fmt],
            "what": "pinned host population in -> tf_fitness + tf_rank_topk + tf_generation -> host fitness "
                    f"vector + next population out; {depth} requests in flight (D2H of step n overlaps H2D of step n+1); "
                    "wall clock around the whole loop incl. the final drain; per-rank bytes"}


# ----------------------------------------------------------------------------------------- CPU legs
def _ref_worker_init(workload):
    """multiprocessing worker: the unmodified PF module + the workload's obstacle set as a ComplexEnvironment."""
    global _REF
    from oracle import ref_py
    pf = ref_py.load_pf()
    cfg = WORKLOADS[workload]
    t = (obstacle_table(cfg, cfg["n_analytic"], 4, SEED + 1) if cfg["n_analytic"]
         else obstacle_table(cfg, cfg["n_voxel_obs"], 10, SEED + 2))
    kinds = ("sphere", "cube", "cylinder")
    env = pf.ComplexEnvironment.__new__(pf.ComplexEnvironment)      # PF:10-16 without the random generator
    env.x_bounds, env.y_bounds, env.z_bounds = [(0, w) for w in cfg["world"]]
    env.obstacles = [{"shape": kinds[int(r[7])], "position": (int(r[0]), int(r[1]), int(r[2])), "size": int(r[3])} for r in t]
    _REF = (pf, env)


def _ref_fitness(rows):
    """Path.calculate_fitness (PF:144-170) of the unmodified reference for a list of chromosomes."""
    pf, env = _REF
    out = []
    for wps in rows:
        p = pf.Path([tuple(int(v) for v in w) for w in wps])
        p.calculate_fitness(env)
        out.append(p.fitness)
    return out


def reference_sample_steps(workload, steps, warmup, per_worker=None, target_step_s=5.0):
    """`steps` generations of the UNMODIFIED reference (PF: "GA_complex _PF.py") on a sample population of the
    workload, all host cores through multiprocessing (fitness is embarrassingly parallel; one Python process = one
    core).  A generation = PF's own select_parents / crossover / mutate (PF:192-208) in the parent process + one
    Path.calculate_fitness (PF:144-170) per chromosome in the workers — the work PF:210-220 does per generation.
    The environment is the workload's analytic obstacle set (cfg3: the 16,384 obstacles its voxel grid is
    rasterised from — the reference has no voxel grid, its cost per waypoint grows with the obstacle count).
    Returns (seconds, units, cores, sample description) over the timed steps."""
    import multiprocessing as mp
    from oracle import c_oracle as O
    from oracle import ref_py
    cfg = WORKLOADS[workload]
    W = cfg["W"]
    wx, wy, wz = cfg["world"]
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")     # the parent may hold a CUDA context: never fork it
    pool = ctx.Pool(cores, initializer=_ref_worker_init, initargs=(workload,))
    rng = np.random.default_rng(SEED)

    def sample_pop(n):
        if cfg["n_analytic"]:
            return O.init_walks(SEED, n, W, (wx // 2, wy // 2, wz // 2), (0, 0, 0), (wx - 1, wy - 1, wz - 1))
        starts = np.zeros((n, 4), np.float32)
        starts[:, :3] = np.floor(rng.uniform(0, 1, size=(n, 3)) * np.array([wx, wy, wz]))
        return O.init_walks(SEED, n, W, (0, 0, 0), (0, 0, 0), (wx - 1, wy - 1, wz - 1), start_pp=starts)

    if per_worker is None:   # calibrate: one chromosome per worker
        one = sample_pop(cores)[..., :3].astype(np.int64)
        t0 = time.perf_counter()
        pool.map(_ref_fitness, [[r] for r in one.tolist()])
        per_worker = int(max(1, min(256, target_step_s / max(time.perf_counter() - t0, 1e-3))))
    S = cores * per_worker
    pop = sample_pop(S)[..., :3].astype(np.int64).tolist()
    _ref_worker_init(workload)
    pf, env = _REF
    ga = pf.GeneticAlgorithm.__new__(pf.GeneticAlgorithm)     # PF:175-182 without the PathPlanner initial population
    ga.population_size, ga.mutation_rate, ga.generations, ga.environment = S, 0.1, steps, env
    import random
    random.seed(SEED)
    np.random.seed(SEED & 0x7FFFFFFF)
    fits = [0.0] * S
    elapsed = 0.0
    for g in range(warmup + steps):
        t0 = time.perf_counter()
        ga.population = [pf.Path([tuple(w) for w in row]) for row in pop]
        for p, f in zip(ga.population, fits):
            p.fitness = f
        nxt = list(pop)
        for j in range(0, S - 1, 2):                       # PF:192-208 per child pair, children replace the pair
            p1, p2 = ga.population[j], ga.population[j + 1]
            c1, c2 = ga.crossover(p1, p2)
            ga.mutate(c1)
            ga.mutate(c2)
            nxt[j], nxt[j + 1] = [tuple(int(v) for v in w) for w in c1.waypoints][:W], [tuple(int(v) for v in w) for w in c2.waypoints][:W]
        chunks = [nxt[i::cores] for i in range(cores)]
        res = pool.map(_ref_fitness, chunks)
        for i in range(cores):
            fits[i::cores] = res[i]
        pop = nxt
        if g >= warmup:
            elapsed += time.perf_counter() - t0
    pool.close()
    pool.join()
    desc = (f"{steps} generations of {S} chromosomes x {W} waypoints sampled from {workload}: unmodified reference "
            f"\"GA_complex _PF.py\" (oracle/_ref) — crossover + mutate (PF:196-208) and Path.calculate_fitness (PF:144-170) "
            f"against the workload's {len(env.obstacles)} analytic obstacles, {cores} processes (multiprocessing), {elapsed:.1f} s")
    return elapsed, S * (W - 1) * steps, cores, desc, S


def cpu_generation_sample(workload, P_sample, threads, steps, warmup):
    """One GA generation (rank + GA step + fitness sweep) of the C port on P_sample chromosomes of
    the same workload (same environment, same operators, same Philox draws)."""
    O, env, prm = oracle_world(workload)
    cfg = WORKLOADS[workload]
    W = cfg["W"]
    wx, wy, wz = cfg["world"]
    O.set_threads(threads)
    rng = np.random.default_rng(SEED)
    if cfg["n_analytic"]:
        pop = O.init_walks(SEED, P_sample, W, (wx // 2, wy // 2, wz // 2), (0, 0, 0), (wx - 1, wy - 1, wz - 1))
    else:
        starts = np.zeros((P_sample, 4), np.float32)
        starts[:, :3] = np.floor(rng.uniform(0, 1, size=(P_sample, 3)) * np.array([wx, wy, wz]))
        pop = O.init_walks(SEED, P_sample, W, (0, 0, 0), (0, 0, 0), (wx - 1, wy - 1, wz - 1), start_pp=starts)
    ga = O.default_ga(**GA_KW)
    fit, _ = O.fitness_f32(env, pop, None, prm)
    times = []
    for g in range(warmup + steps):
        t0 = time.perf_counter()
        top = O.topk(fit, 2)
        pop, _, _, _ = O.ga_step(pop, fit, top, ga, seed=SEED, gen=g)
        fit, _ = O.fitness_f32(env, pop, None, prm)
        dt = time.perf_counter() - t0
        if g >= warmup:
            times.append(dt)
    O.set_threads(1)
    return float(np.sum(times)), P_sample * (W - 1) * steps


def cpu_baseline_port(workload, seconds=6.0):
    """Bounded sample (~`seconds` of CPU work) of the same workload through the C port."""
    threads = os.cpu_count() or 1
    cfg = WORKLOADS[workload]
    P_sample = 4096 if cfg["n_analytic"] else 65536
    dt, units = cpu_generation_sample(workload, P_sample, threads, steps=1, warmup=1)   # calibrate
    steps = int(max(1, min(200, seconds / max(dt, 1e-6))))
    dt, units = cpu_generation_sample(workload, P_sample, threads, steps=steps, warmup=0)
    return {"value": units / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} generations of {P_sample} chromosomes x {cfg['W']} waypoints of {workload} "
                      f"(oracle/tf_oracle.c FP32 model: rank + GA step + fitness sweep), {dt:.1f} s",
            "seconds": dt}


def cpu_baseline_reference(workload, steps=2):
    """The unmodified reference on a bounded sample (kind "reference"), or None when the workload has no
    counterpart in the reference (hazard map / moving spheres) or oracle/_ref is absent."""
    from oracle import ref_py
    if WORKLOADS[workload]["hazard"] or WORKLOADS[workload]["n_dyn"] or not ref_py.available():
        return None
    dt, units, cores, desc, _ = reference_sample_steps(workload, steps=steps, warmup=0, target_step_s=5.0)
    return {"value": units / dt, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc, "seconds": dt}


def run_cfg1(with_reference=True):
    """configs[0] through the drop-in class (GPU, host in / host out, wall clock) and, beside it, the unmodified
    reference HRMAGAGenerator(100, 0.1, 200).evolve((0,0,0), (19,19,8), obstacles) run in full on one core.
    Scenario, obstacles, seed and the path the reference returned come from the committed fixture
    tests/golden/evolve_hr_cfg1.npz (recorded from the unmodified reference by tests/golden/gen_golden.py)."""
    import torch
    from tracefly_b200 import _lib as L
    from tracefly_b200 import path_planner_hrmaga as HRB
    g = np.load(os.path.join(ROOT, "tests", "golden", "evolve_hr_cfg1.npz"))
    meta = json.loads(str(g["meta"]))
    obs = [{"position": [int(v) for v in p], "size": int(s)} for p, s in zip(g["obs_pos"], g["obs_size"])]
    P, G, W, rate = meta["P"], meta["G"], meta["W"], meta["rate"]
    start, end = tuple(meta["start"]), tuple(meta["end"])
    units = (G + 1) * P * (W - 1)
    times, launches, path = [], 0, None
    for rep in range(4):          # the first call pays module / kernel loading
        np.random.seed(meta["seed"])
        n0 = L.launch_count()
        t0 = time.perf_counter()
        path = HRB.HRMAGAGenerator(P, rate, G).evolve(start, end, obs)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = L.launch_count() - n0
    dt = float(np.median(times[1:]))
    same = [tuple(map(float, w)) for w in path] == [tuple(w) for w in g["best"].tolist()]
    out = {"workload": f"cfg1: HRMAGAGenerator({P}, {rate}, {G}).evolve({start}, {end}, 10 obstacles): {G + 1} sweeps x {P} paths x {W - 1} segments",
           "value": units / dt, "unit": UNIT, "seconds_per_evolve": dt, "generations_per_sec": G / dt, "dtype": "f64",
           "what": "drop-in tracefly_b200.path_planner_hrmaga.HRMAGAGenerator.evolve, wall clock, host lists in / host list out "
                   "(FP64 reference-order kernels; median of 3 calls after one warm-up call)",
           "gpu_launches_per_evolve": int(launches), "same_path_as_recorded_reference_run": bool(same),
           "roofline": {"bound": "latency", "achieved": None, "peak": None, "frac": None,
                        "note": "100 x 20 x 32 B = 64 KB per sweep: launch- / latency-bound by construction, no HBM roofline applies"},
           "recorded_reference": {"seconds": meta["ref_seconds"], "value": meta["units_per_s"],
                                  "note": "timed in the build container when the golden was recorded (1 core)"}}
    if with_reference:
        from oracle import ref_py
        if ref_py.available():
            hr = ref_py.load_hr()
            np.random.seed(meta["seed"])
            t0 = time.perf_counter()
            rpath = hr.HRMAGAGenerator(P, rate, G).evolve(start, end, obs)
            rdt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": units / rdt, "unit": UNIT, "cores": 1, "kind": "reference", "seconds": rdt,
                                   "sample": "the whole scenario, unmodified path_planner_hrmaga.py (oracle/_ref), one Python process"}
            out["same_path_as_reference_run_here"] = bool([tuple(map(float, w)) for w in rpath] == [tuple(map(float, w)) for w in path])
            out["speedup_vs_reference"] = rdt / dt
    return out


def summarize_workload(res, cpu_ref, cpu_port, gate):
    out = {"workload": res["config"]["workload"], "value": res["value"], "unit": UNIT, "steps": res["steps"], "warmup": res["warmup"],
           "ms_per_step": res["ms_per_step"], "generations_per_sec": res["generations_per_sec"], "dtype": "f32",
           "roofline": res["roofline"], "gpu_launches": res["gpu_launches"], "config": res["config"], "parity_gate": gate}
    if cpu_ref:
        out["cpu_baseline"] = cpu_ref
        out["cpu_baseline_port"] = cpu_port
    elif cpu_port:
        out["cpu_baseline"] = cpu_port
    else:
        out["cpu_baseline"] = {"see": "the cfg3 line: same world, same operators, only the population size differs"}
    return out


def run_gpu(args):
    out = _claim_stdout()
    ctx = Ctx()
    torch = ctx.torch
    name = args.workload
    cfg = dict(WORKLOADS[name])
    if args.pop_per_gpu:
        cfg["P"] = args.pop_per_gpu
    P, W = cfg["P"], cfg["W"]
    res, island, env, params, ga = time_generations(ctx, name, cfg, args.steps, args.warmup)

    # ---- correctness gate on the population the timed run ended with (rank 0's island) + N > 1: island world
    gate = None
    oracle = None
    if not args.no_gate:
        if ctx.rank == 0:
            oracle = oracle_world(name)
            gate = parity_gate(ctx, name, island, env, params, ga, oracle)
        if ctx.world > 1:
            ig = island_gate(ctx)
            if ctx.rank == 0:
                gate["islands"] = ig
                gate["ok"] = gate["ok"] and ig["ok"]
        ctx.sync_all()

    # ---- e2e through the host-array API
    e2e = e2e_padded = e2e_compact = None
    if not args.no_e2e:
        del island.nxt, island.nfit          # make room: the stepper owns its own device buffers
        torch.cuda.empty_cache()
        e2e_steps = max(4, min(args.steps, 10))
        e2e = run_e2e(ctx, island, env, params, ga, "xyz", e2e_steps)
        ceiling = pcie_ceiling(ctx, e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"])
        e2e["pcie_ceiling"] = ceiling
        e2e["fraction_of_pcie_ceiling"] = ceiling["ms_per_step"] / e2e["ms_per_step"]
        e2e_padded = run_e2e(ctx, island, env, params, ga, "f32", e2e_steps)
        e2e_compact = run_e2e(ctx, island, env, params, ga, "i16", e2e_steps)
    del island
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, short timed runs
    extras = {}
    if not args.no_extras:
        k, w = min(args.steps, 20), min(args.warmup, 5)
        # cfg5's per-GPU size also at N = 1: the driver's 1/2/4/8 runs then hold a weak-scaling curve of cfg5 itself
        names = [n for n in (("cfg2", "cfg4", "cfg5") if ctx.world == 1 else ("cfg5",)) if n != name]
        for n2 in names:
            env = params = ga = None
            torch.cuda.empty_cache()
            r2, isl2, env2, prm2, ga2 = time_generations(ctx, n2, dict(WORKLOADS[n2]), k if n2 != "cfg5" else min(k, 10), w)
            g2 = None
            if ctx.rank == 0 and not args.no_gate and n2 != "cfg5":
                g2 = parity_gate(ctx, n2, isl2, env2, prm2, ga2)
            if ctx.rank == 0 and n2 == "cfg5" and not args.no_gate and oracle is not None:
                g2 = parity_gate(ctx, "cfg3", isl2, env2, prm2, ga2, oracle)    # same world as cfg3
            del isl2, env2
            torch.cuda.empty_cache()
            if ctx.rank == 0:
                cref = cport = None
                if ctx.world == 1 and not args.no_cpu and n2 != "cfg5":   # cfg5 = cfg3's world and operators: its CPU legs are cfg3's
                    cport = cpu_baseline_port(n2, seconds=3.0)
                    cref = cpu_baseline_reference(n2, steps=1)
                extras[n2] = summarize_workload(r2, cref, cport, g2)
            ctx.sync_all()
        if ctx.world == 1 and ctx.rank == 0 and name != "cfg1":
            extras["cfg1"] = run_cfg1(with_reference=not args.no_cpu)

    if ctx.rank != 0:
        if ctx.world > 1:
            ctx.dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded lattice random walks, rasterised seeded obstacles; generated on device)",
        "generations_per_sec": res["generations_per_sec"], "config": res["config"], "roofline": res["roofline"],
        "clocks": res["clocks"], "gpu_launches": res["gpu_launches"],
    }
    if "trace" in res:
        line["trace"] = res["trace"]
    if gate is not None:
        line["parity_gate"] = gate
    if e2e:
        line["e2e"] = e2e
        line["e2e_padded_host_format"] = e2e_padded
        line["e2e_compact_host_format"] = e2e_compact
    if ctx.world == 1 and not args.no_cpu:
        port = cpu_baseline_port(name, seconds=args.cpu_seconds)
        ref = cpu_baseline_reference(name, steps=2)
        if ref:
            line["cpu_baseline"] = ref
            line["cpu_baseline_port"] = port
        else:
            line["cpu_baseline"] = port
    if extras:
        line["workloads"] = extras
    out.write(json.dumps(line) + "\n")
    out.flush()
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import ref_py
    threads = os.cpu_count() or 1
    name = args.workload
    cfg = WORKLOADS[name]
    use_ref = ref_py.available() and not (cfg["hazard"] or cfg["n_dyn"]) and not args.port
    P_sample = 4096 if cfg["n_analytic"] else 65536
    pdt, punits = cpu_generation_sample(name, P_sample, threads, steps=max(1, min(args.steps, 20)), warmup=1)
    port = {"value": punits / pdt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"generations of {P_sample} chromosomes x {cfg['W']} waypoints of {name}, C port oracle/tf_oracle.c, {threads} pthreads"}
    if use_ref:
        # one chromosome per worker process and step: a cfg3 chromosome costs the reference ~5 s (64 waypoints x 16,384
        # obstacles x 2 passes), so K + W steps stay within a few minutes
        dt, units, cores, desc, P_used = reference_sample_steps(name, steps=args.steps, warmup=args.warmup,
                                                                per_worker=1 if not cfg["n_analytic"] else None)
        value, kind, sample = units / dt, "reference", desc
        ms_per_step = dt / args.steps * 1e3
    else:
        value, kind, sample, cores = port["value"], "port", port["sample"] + (
            " (the reference has no hazard map / moving spheres)" if (cfg["hazard"] or cfg["n_dyn"]) else " (oracle/_ref absent)"), threads
        ms_per_step = pdt / max(1, min(args.steps, 20)) * 1e3
        P_used = P_sample
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if kind == "reference" else "f32",
            "data": "synthetic (same seeded workload as the GPU arm, bounded sample)",
            "config": {"workload": workload_label(name, cfg, cfg["P"]), "sample": sample, "sample_chromosomes_per_step": P_used},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "cpu_baseline_port": port,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--pop-per-gpu", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gate", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the selected workload (skip the other BASELINE configs)")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the C port even when oracle/_ref exists")
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
