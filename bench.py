#!/usr/bin/env python
"""bench.py — waypoint-segment fitness evaluations/s (+ GA generations/s) of the TraceFly GA hot
path on N B200s, next to the CPU restatement timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg4] [--impl reference]

A "step" is one full GA generation over one island: rank (tf_rank_topk) -> selection + crossover
+ mutation + fitness sweep (fused tf_generation).  With N > 1 (torchrun, one process per GPU) each
rank owns one island of the same size (weak scaling) and the islands exchange elites with an NCCL
all-gather every 10 generations.  Unit: one waypoint-segment evaluation: a
generation over P chromosomes of W waypoints is P*(W-1) units.

Workloads (BASELINE.json configs; synthetic, seeded, generated on the device):
  cfg3 (default, "the 1M-path config" the north_star target is quoted on):
        P = 1,048,576 x W = 64 vs a 1024^3 bit-packed voxel grid (128 MiB, rasterised from 16,384
        seeded spheres / cubes / cylinders), paths start uniformly over the world.
  cfg2  P = 65,536 x W = 32 vs 1,024 analytic sphere / cube / cylinder obstacles (FP32-pipe bound).
  cfg4  cfg3 + hazard heat-map f32[24,256,256] + 1,024 moving spheres.
Inputs are far larger than L2 (population 1 GiB, ping-ponged between two buffers; grid 128 MiB).

--impl reference times the CPU port of the same generation (oracle/tf_oracle.c, all host threads)
on a bounded sample of the same workload: the reference itself is Python and /root/reference does
not exist on the GPU box.
"""
import argparse
import json
import os
import subprocess
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
}
SEED = 0x5EED
MIGRATE_EVERY = 10
ELITES = 1024


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


def algorithmic_bytes_per_unit(W, fused=True):
    """fused generation per child = read 16W (parent halves) + write 16W + 4 (fitness)
    + 8 (fitness / rank reads)."""
    return (32 * W + 12) / (W - 1)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region through NVML (the same counters
    nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.* prints), every 5 ms."""

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
            time.sleep(0.005)

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
        env.build_dynamic_index(W, (-m, -m, -m), (wx + m, wy + m, wz + m), cell=float(os.environ.get("TF_DYN_CELL", 32)),
                                cap=int(os.environ.get("TF_DYN_CAP", 48)), band=int(os.environ.get("TF_DYN_BAND", 16)))
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


def run_gpu(args):
    out = _claim_stdout()
    import torch
    import torch.distributed as dist
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    from tracefly_b200.islands import IslandGA

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = dict(WORKLOADS[args.workload])
    if args.pop_per_gpu:
        cfg["P"] = args.pop_per_gpu
    P, W = cfg["P"], cfg["W"]
    env, params, pop, ga = build_gpu_workload(cfg, rank, torch, B, L)
    island = IslandGA(pop, ga, seed=SEED, elites=ELITES, interval=MIGRATE_EVERY, env=env, params=params)
    stream = torch.cuda.current_stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        island.step()
    if world > 1:
        island.migrate()  # untimed: NCCL communicator set-up and first-collective connect happen here
    sync_all()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
    gen_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.launch_count()
    sync_all()
    sampler.start()
    t0.record(stream)
    for i in range(args.steps):
        # the per-step events bracket ONLY the dominant kernel (tf_generation) for the roofline
        island.ops.rank(island.fit, island.top)
        gen_ev[i][0].record(stream)
        island.launch_generation()
        gen_ev[i][1].record(stream)
        island.pop, island.nxt = island.nxt, island.pop
        island.fit, island.nfit = island.nfit, island.fit
        island.gen += 1
        island.after_generation()   # N > 1: pipelined elite exchange (side stream) every 10 generations
    island.flush()
    t1.record(stream)
    sync_all()
    launches = L.launch_count() - launches0
    clocks = sampler.finish()
    ms_total = t0.elapsed_time(t1)
    gen_times = [a.elapsed_time(b) for a, b in gen_ev]
    gen_ms = float(np.mean(gen_times))
    q = max(1, len(gen_times) // 4)
    gen_ms_first, gen_ms_last = float(np.mean(gen_times[:q])), float(np.mean(gen_times[-q:]))
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    units_per_step = world * P * (W - 1)
    value = units_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: the public host-array API (batched.HostStepper): every step copies that step's
    #      population from pinned host memory to the GPU, runs tf_fitness + tf_rank_topk +
    #      tf_generation, and copies the fitness vector and the next population back; two requests are
    #      in flight so the D2H leg of step n overlaps the H2D leg of step n+1.
    e2e = None
    e2e_compact = None
    e2e_padded = None
    if not args.no_e2e:
        del island.nxt, island.nfit          # make room: the stepper owns its own device buffers
        torch.cuda.empty_cache()

        def run_e2e(fmt, steps):
            depth = int(os.environ.get("TF_E2E_DEPTH", 2))
            stepper = B.HostStepper(env, P, W, ga, params, fmt=fmt, depth=depth)
            hdt, hc = B.HostStepper._HOST[fmt]
            hosts = [torch.empty((P, W, hc), dtype=hdt).pin_memory() for _ in range(depth)]
            for h in hosts:
                h.copy_(island.pop[..., :hc].to(hdt))
            sync_all()
            tickets = []
            checksum = 0.0
            t_start = None
            for it in range(steps + depth):      # `depth` untimed warm-up requests
                if it == depth:
                    while tickets:
                        stepper.result(tickets.pop(0))
                    sync_all()
                    t_start = time.perf_counter()
                if len(tickets) == depth:
                    fit_h, next_h = stepper.result(tickets.pop(0))
                    checksum += float(fit_h[0])      # the caller really reads the result
                tickets.append(stepper.submit(hosts[it % depth], seed=SEED, generation=1000 + it))
            while tickets:
                fit_h, next_h = stepper.result(tickets.pop(0))
                checksum += float(fit_h[0])
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t_start) * 1e3
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            h2d, d2h = stepper.bytes_per_step()
            del stepper, hosts
            torch.cuda.empty_cache()
            return {"value": units_per_step * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms / steps, "steps": steps,
                    "host_format": {"f32": "float32[P,W,4] (the device layout)", "xyz": "float32[P,W,3] (the reference's 3-tuples as an array)",
                                    "i16": "int16[P,W,4] (lattice chromosomes)"}[fmt],
                    "what": "pinned host population in -> tf_fitness + tf_rank_topk + tf_generation -> host fitness "
                            f"vector + next population out; {depth} requests in flight (D2H of step n overlaps H2D of step n+1); "
                            "wall clock around the whole loop incl. the final drain"}

        e2e_steps = max(4, min(args.steps, 10))
        e2e = run_e2e("xyz", e2e_steps)
        e2e_padded = run_e2e("f32", e2e_steps)
        e2e_compact = run_e2e("i16", e2e_steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak_gbs()
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f).get(args.workload)
        if t and P == WORKLOADS[args.workload]["P"]:
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass
    bpu = algorithmic_bytes_per_unit(W)
    achieved = bpu * P * (W - 1) / (gen_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded lattice random walks, rasterised seeded obstacles; generated on device)",
        "generations_per_sec": args.steps / (ms_total * 1e-3),
        "config": {"workload": f"{args.workload}: pop {P:,} x {W} waypoints per GPU"
                               + (f" vs {cfg['world'][0]}^3 bit-packed voxel grid" if cfg["n_voxel_obs"] else "")
                               + (f" vs {cfg['n_analytic']} analytic sphere/cube/cylinder obstacles" if cfg["n_analytic"] else "")
                               + (" + hazard heat-map 24x256x256 + 1024 moving spheres" if cfg["hazard"] else ""),
                   "step": "one GA generation = tf_rank_topk + fused tf_generation (tournament k=4, single-point crossover, "
                           "lattice mutation 0.1, 2 elites, Philox4x32-10)",
                   "population_per_gpu": P, "waypoints": W, "islands": world,
                   "migration": f"{ELITES} elites NCCL all-gather every {MIGRATE_EVERY} generations, pipelined with the next generation on a side stream" if world > 1 else "none (1 island)",
                   "l2": "inputs larger than L2: population 2 x %.0f MiB ping-pong, voxel grid %d MiB" % (
                       P * W * 16 / 2 ** 20, (cfg["world"][0] * cfg["world"][1] * cfg["world"][2] // 8) >> 20 if cfg["n_voxel_obs"] else 0)},
        "roofline": {"bound": "hbm", "kernel": "tf_generation (TMA-staged fused generation kernel: duo_kernel<GA> without an obstacle table, stream_kernel<GA> with one)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bpu * P * (W - 1),
                     "algorithmic_bytes_per_unit": bpu, "kernel_ms": gen_ms,
                     "kernel_ms_first_quarter": gen_ms_first, "kernel_ms_last_quarter": gen_ms_last,
                     "kernel_share_of_step": gen_ms / (ms_total / args.steps)},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if e2e:
        line["e2e"] = e2e
        line["e2e_padded_host_format"] = e2e_padded
        line["e2e_compact_host_format"] = e2e_compact
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.workload, seconds=args.cpu_seconds)
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_generation_sample(workload, P_sample, threads, steps, warmup):
    """One GA generation (rank + GA step + fitness sweep) of the C port on P_sample chromosomes of
    the same workload (same environment, same operators, same Philox draws)."""
    from oracle import c_oracle as O
    cfg = WORKLOADS[workload]
    W = cfg["W"]
    wx, wy, wz = cfg["world"]
    O.set_threads(threads)
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
    rng = np.random.default_rng(SEED)
    if cfg["n_analytic"]:
        pop = O.init_walks(SEED, P_sample, W, (wx // 2, wy // 2, wz // 2), (0, 0, 0), (wx - 1, wy - 1, wz - 1))
    else:
        starts = np.zeros((P_sample, 4), np.float32)
        starts[:, :3] = np.floor(rng.uniform(0, 1, size=(P_sample, 3)) * np.array([wx, wy, wz]))
        pop = O.init_walks(SEED, P_sample, W, (0, 0, 0), (0, 0, 0), (wx - 1, wy - 1, wz - 1), start_pp=starts)
    ga = O.default_ga(selection=1, tournament_k=4, crossover=0, mutation=0, mutation_rate=0.1, n_elite=2)
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


def cpu_baseline(workload, seconds=15.0):
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = WORKLOADS[args.workload]
    P_sample = 4096 if cfg["n_analytic"] else 65536
    dt, units = cpu_generation_sample(args.workload, P_sample, threads, steps=args.steps, warmup=max(args.warmup, 1))
    value = units / dt
    sample = (f"each step = one generation (rank + GA step + fitness sweep) of {P_sample} chromosomes x {cfg['W']} "
              f"waypoints sampled from {args.workload}; C port oracle/tf_oracle.c (the reference is Python and "
              f"/root/reference is absent on the GPU box), {threads} pthreads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (same seeded workload as the GPU arm, bounded sample)",
            "config": {"workload": f"{args.workload}: sample of {P_sample} x {cfg['W']} waypoints", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--pop-per-gpu", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
