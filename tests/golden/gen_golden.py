#!/usr/bin/env python
"""Generate the golden input/output vectors under tests/golden/ by running the
UNMODIFIED TraceFly reference (imported read-only from /root/reference).

The reference has no tests and no golden vectors of its own, and
being Python it cannot travel to the GPU box, so the vectors made here are the
pin for oracle/ and, through it, for the CUDA path.

    python -B tests/golden/gen_golden.py            # writes tests/golden/*.npz

Recorded with CPython 3.12.3 / NumPy 2.3.5 (versions are stored in each file).
Every array is produced by reference code: HR = path_planner_hrmaga.py,
PF = "GA_complex _PF.py".  Nothing in this script restates the algorithm; it
only builds inputs, calls the reference and records what it returned (plus the
np.random draws it consumed, via wrappers, i.e. the "draw tape").
"""
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

META = {"numpy": np.__version__, "python": sys.version.split()[0],
        "generator": "tests/golden/gen_golden.py"}


# --------------------------------------------------------------------------- inputs
def lattice_walks(rng, P, W, start, lo, hi):
    """Seeded +-1 lattice random walks (what HR:91-99 mutation produces)."""
    steps = rng.integers(-1, 2, size=(P, W - 1, 3))
    wp = np.concatenate([np.zeros((P, 1, 3), np.int64), np.cumsum(steps, axis=1)], axis=1)
    wp = wp + np.asarray(start, np.int64)
    return np.clip(wp, lo, hi)


def as_tuples(row, n, to_py):
    """Waypoint tuples typed the way the reference's own generators type them (HR:68-80,
    PF:95-119): the first waypoint is the caller's tuple of Python scalars, every later one
    is a tuple of NumPy scalars (np.int64 / np.float64).  This matters: CPython >= 3.12
    sum() switches to compensated summation for runs of exact Python floats, which pure
    Python-scalar paths could trigger; NumPy-scalar paths sum naively, left to right."""
    np_t = np.int64 if to_py is int else np.float64
    return [tuple((to_py if i == 0 else np_t)(v) for v in row[i]) for i in range(n)]


def hr_obstacles(pos, size):
    out = []
    for p, s in zip(pos, size):
        d = {"position": [float(v) if isinstance(v, (float, np.floating)) else int(v) for v in p]}
        if s is not None and not (isinstance(s, float) and np.isnan(s)):
            d["size"] = s
        out.append(d)
    return out


# --------------------------------------------------------------------------- HR fitness
def gen_fitness_hr(hr):
    rng = np.random.default_rng(1234)
    cases = []

    def run_case(name, wp, lens, pos, size, to_py):
        P, W, _ = wp.shape
        obstacles = hr_obstacles(pos, size)
        fit = np.zeros(P)
        coll = np.zeros((P, W), bool)
        dnear = np.full((P, W), np.inf)
        for p in range(P):
            path = hr.Path(as_tuples(wp[p], lens[p], to_py))
            path.calculate_fitness(obstacles)
            fit[p] = path.fitness
            for i, pt in enumerate(path.waypoints):
                coll[p, i] = hr.is_collision(pt, obstacles)
                dnear[p, i] = hr.distance_to_nearest_obstacle(pt, obstacles)
        cases.append(name)
        sz = np.array([np.nan if s is None else float(s) for s in size], np.float64)
        return {f"{name}/wp": wp.astype(np.float64), f"{name}/lens": lens.astype(np.int32),
                f"{name}/obs_pos": np.asarray(pos, np.float64).reshape(-1, 3),
                f"{name}/obs_size": sz, f"{name}/fitness": fit, f"{name}/collide": coll,
                f"{name}/dnear": dnear}

    out = {}
    # lattice chromosomes, integer obstacles with integer sizes (PF:17-26 distribution)
    for (P, W, N) in [(64, 3, 4), (128, 16, 10), (128, 20, 10), (96, 32, 64), (48, 64, 32), (32, 33, 7)]:
        wp = lattice_walks(rng, P, W, (10, 10, 5), 0, 24)
        pos = rng.integers(0, 24, size=(N, 3))
        size = [int(s) for s in rng.integers(1, 5, size=N)]
        out.update(run_case(f"lat_W{W}_N{N}", wp, np.full(P, W), pos, size, int))
    # producer never writes "size" (lidar_hybrid_model_validate.py:150-154) -> default 1.0 (HR:20)
    wp = lattice_walks(rng, 64, 16, (8, 8, 4), 0, 16)
    pos = rng.integers(0, 16, size=(12, 3))
    out.update(run_case("lat_nosize", wp, np.full(64, 16), pos, [None] * 12, int))
    # N = 0 -> distance inf -> proximity term 0 (HR:26-27)
    out.update(run_case("lat_noobs", wp[:16], np.full(16, 16), np.zeros((0, 3)), [], int))
    # boundary: points exactly at ||p-c|| == size + 2 (Pythagorean offsets), and one step outside
    c = np.array([10, 10, 10])
    offs = np.array([[3, 0, 0], [0, 3, 0], [1, 2, 2], [2, 1, 2], [2, 2, 1], [0, 0, 3],
                     [4, 0, 0], [3, 1, 0], [2, 2, 2], [1, 1, 3], [0, 0, 0], [6, 0, 0],
                     [4, 4, 2], [3, 4, 0], [0, 5, 0], [5, 1, 0]])
    wp = (c + offs)[None, :, :].repeat(2, axis=0)
    wp[1] = wp[1][::-1]
    out.update(run_case("boundary_r3", wp, np.full(2, 16), c[None, :], [1], int))
    out.update(run_case("boundary_r6", wp, np.full(2, 16), c[None, :], [4], int))
    # all waypoints identical -> distance == 0 guard (HR:53), turns == 0
    wp = np.tile(np.array([[[3, 3, 3]]]), (4, 8, 1))
    out.update(run_case("zero_dist", wp, np.full(4, 8), rng.integers(0, 8, size=(3, 3)), [1, 2, 3], int))
    # float (non-lattice) waypoints as MAIN:52-56 produces, float obstacle positions/sizes
    wpf = lattice_walks(rng, 64, 12, (5, 5, 5), 0, 12).astype(np.float64)
    wpf += rng.uniform(-0.45, 0.45, size=wpf.shape)
    posf = rng.uniform(0, 12, size=(9, 3))
    sizef = [float(s) for s in rng.uniform(0.5, 3.0, size=9)]
    out.update(run_case("float_W12_N9", wpf, np.full(64, 12), posf, sizef, float))
    # ragged lengths inside one batch
    wp = lattice_walks(rng, 40, 24, (10, 10, 5), 0, 24)
    lens = rng.integers(3, 25, size=40)
    pos = rng.integers(0, 24, size=(10, 3))
    size = [int(s) for s in rng.integers(1, 5, size=10)]
    out.update(run_case("ragged_W24", wp, lens, pos, size, int))
    out["meta"] = np.array(json.dumps(dict(META, cases=cases, ref="HR:17-28,33-59")))
    np.savez_compressed(os.path.join(HERE, "fitness_hr.npz"), **out)
    print("fitness_hr.npz:", len(cases), "cases")


# --------------------------------------------------------------------------- PF fitness
KINDS = ["sphere", "cube", "cylinder"]


class _Env:
    pass


def gen_fitness_pf(pf):
    rng = np.random.default_rng(4321)
    cases = []
    out = {}

    def run_case(name, wp, lens, pos, size, kind, to_py):
        P, W, _ = wp.shape
        env = pf.ComplexEnvironment.__new__(pf.ComplexEnvironment)
        env.obstacles = [{"shape": KINDS[k], "position": tuple(to_py(v) for v in p), "size": to_py(s)}
                         for p, s, k in zip(pos, size, kind)]
        fit = np.zeros(P)
        coll = np.zeros((P, W), bool)
        dnear = np.full((P, W), np.inf)
        with np.errstate(divide="ignore"):
            for p in range(P):
                path = pf.Path(as_tuples(wp[p], lens[p], to_py))
                path.calculate_fitness(env)
                fit[p] = path.fitness
                for i, pt in enumerate(path.waypoints):
                    coll[p, i] = env.is_obstacle(pt)
                    dnear[p, i] = env.distance_to_obstacles(pt)
        cases.append(name)
        out.update({f"{name}/wp": wp.astype(np.float64), f"{name}/lens": lens.astype(np.int32),
                    f"{name}/obs_pos": np.asarray(pos, np.float64).reshape(-1, 3),
                    f"{name}/obs_size": np.asarray(size, np.float64),
                    f"{name}/obs_kind": np.asarray(kind, np.int32),
                    f"{name}/fitness": fit, f"{name}/collide": coll, f"{name}/dnear": dnear})

    for (P, W, N) in [(64, 3, 6), (128, 16, 10), (128, 21, 10), (96, 32, 96), (48, 64, 33)]:
        wp = lattice_walks(rng, P, W, (10, 10, 5), 0, 24)
        pos = rng.integers(0, 24, size=(N, 3))
        size = rng.integers(1, 5, size=N)
        kind = rng.integers(0, 3, size=N)
        run_case(f"lat_W{W}_N{N}", wp, np.full(P, W), pos, size, kind, int)
    # one kind only (exercises empty kind buckets)
    for k in range(3):
        wp = lattice_walks(rng, 32, 16, (8, 8, 4), 0, 16)
        pos = rng.integers(0, 16, size=(8, 3))
        size = rng.integers(1, 5, size=8)
        run_case(f"only_{KINDS[k]}", wp, np.full(32, 16), pos, size, np.full(8, k), int)
    run_case("noobs", wp[:8], np.full(8, 16), np.zeros((0, 3)), [], [], int)
    # boundary offsets around one obstacle of each kind with size 3
    c = np.array([10, 10, 10])
    offs = np.array([[3, 0, 0], [3, 3, 3], [1, 2, 2], [2, 2, 1], [3, 3, 6], [3, 3, 7], [0, 0, 6],
                     [4, 0, 0], [3, 1, 0], [2, 2, 2], [0, 0, 0], [3, 4, 0], [0, 0, 7], [4, 3, 6],
                     [1, 1, 3], [2, 1, 2]])
    wp = (c + offs)[None].repeat(2, axis=0)
    wp[1] = wp[1][::-1]
    for k in range(3):
        run_case(f"boundary_{KINDS[k]}", wp, np.full(2, 16), c[None], [3], [k], int)
    # float inputs
    wpf = lattice_walks(rng, 64, 12, (5, 5, 5), 0, 12).astype(np.float64)
    wpf += rng.uniform(-0.45, 0.45, size=wpf.shape)
    posf = rng.uniform(0, 12, size=(9, 3))
    sizef = rng.uniform(0.5, 3.0, size=9)
    run_case("float_W12_N9", wpf, np.full(64, 12), posf, sizef, rng.integers(0, 3, size=9), float)
    # ragged (PF populations have mixed lengths)
    wp = lattice_walks(rng, 40, 24, (10, 10, 5), 0, 24)
    lens = rng.integers(3, 25, size=40)
    pos = rng.integers(0, 24, size=(10, 3))
    run_case("ragged_W24", wp, lens, pos, rng.integers(1, 5, size=10), rng.integers(0, 3, size=10), int)
    # all-identical waypoints: PF has no distance guard (PF:163) -> numpy 1/0.0 = inf
    wp = np.tile(np.array([[[3, 3, 3]]]), (2, 6, 1))
    run_case("zero_dist", wp, np.full(2, 6), rng.integers(0, 8, size=(3, 3)), [1, 2, 3], [0, 1, 2], int)
    out["meta"] = np.array(json.dumps(dict(META, cases=cases, kinds=KINDS, ref="PF:28-59,141-170")))
    np.savez_compressed(os.path.join(HERE, "fitness_pf.npz"), **out)
    print("fitness_pf.npz:", len(cases), "cases")


# --------------------------------------------------------------------------- HR evolve traces
class Recorder:
    """Wraps np.random.{choice,randint,rand} to record the draws the reference
    consumes, in order, without changing what they return."""

    def __init__(self):
        self.events = []
        self._orig = {}

    def __enter__(self):
        for name in ("choice", "randint", "rand"):
            self._orig[name] = getattr(np.random, name)
            setattr(np.random, name, self._wrap(name, self._orig[name]))
        return self

    def __exit__(self, *exc):
        for name, fn in self._orig.items():
            setattr(np.random, name, fn)

    def _wrap(self, name, fn):
        def inner(*a, **k):
            r = fn(*a, **k)
            self.events.append((name, a, k, r))
            return r
        return inner


def pack_delta(d):
    return int((d[0] + 1) | ((d[1] + 1) << 2) | ((d[2] + 1) << 4))


def gen_evolve_hr(hr, pf, name, P, rate, G, start, end, obstacles, seed):
    sweeps = []  # one list per calculate_fitness call: (waypoints, fitness)
    orig_cf = hr.Path.calculate_fitness

    def cf(self, obs):
        orig_cf(self, obs)
        sweeps.append((np.array(self.waypoints, dtype=np.float64), float(self.fitness)))

    hr.Path.calculate_fitness = cf
    np.random.seed(seed)
    t0 = time.time()
    try:
        with Recorder() as rec:
            ga = hr.HRMAGAGenerator(P, rate, G)
            best = ga.evolve(start, end, obstacles)
    finally:
        hr.Path.calculate_fitness = orig_cf
    dt = time.time() - t0
    W = len(best)
    assert len(sweeps) == (G + 1) * P
    pops = np.stack([s[0] for s in sweeps]).reshape(G + 1, P, W, 3)
    fits = np.array([s[1] for s in sweeps]).reshape(G + 1, P)
    # decode the event stream into the canonical per-pair tape (see include/tracefly_b200.h)
    npairs = (P - 2 + 1) // 2
    tape = np.zeros((G, npairs, 8), np.int32)
    gates = np.zeros((G, npairs, 2), np.float64)
    ev = iter(rec.events)
    for g in range(G):
        for j in range(npairs):
            nm, a, k, r = next(ev)
            assert nm == "choice" and len(a[0]) == min(10, P)
            pool = a[0]
            ia = [i for i, x in enumerate(pool) if x is r[0]][0]
            ib = [i for i, x in enumerate(pool) if x is r[1]][0]
            nm, a, k, cp = next(ev)
            assert nm == "randint"
            row = [ia, ib, int(cp)]
            for c in range(2):
                nm, a, k, gate = next(ev)
                assert nm == "rand"
                gates[g, j, c] = gate
                if gate < rate:
                    nm, a, k, mp = next(ev)
                    assert nm == "randint"
                    nm, a, k, d = next(ev)
                    assert nm == "choice" and len(d) == 3
                    row += [int(mp), pack_delta(d)]
                else:
                    row += [-1, 0]
            tape[g, j, :7] = row
    assert next(ev, None) is None
    opos = np.array([o["position"] for o in obstacles], np.float64).reshape(-1, 3)
    osz = np.array([o.get("size", np.nan) for o in obstacles], np.float64)
    units = (G + 1) * P * (W - 1)
    meta = dict(META, P=P, rate=rate, G=G, start=list(start), end=list(end), seed=seed, W=W,
                ref="HR:101-130", ref_seconds=dt, units=units, units_per_s=units / dt)
    np.savez_compressed(os.path.join(HERE, f"evolve_hr_{name}.npz"),
                        pops=pops.astype(np.int16 if np.all(pops == np.round(pops)) else np.float64),
                        fits=fits, tape=tape, gates=gates, best=np.array(best, np.float64),
                        obs_pos=opos, obs_size=osz, meta=np.array(json.dumps(meta)))
    print(f"evolve_hr_{name}.npz: {dt:.1f}s reference wall, W={W}, {units / dt:.0f} units/s")


def pf_seeded_obstacles(pf, seed, bounds, n):
    """PF:17-26 obstacle generator driven at a fixed seed (PF:256-259 scenario)."""
    random.seed(seed)
    np.random.seed(seed)
    env = pf.ComplexEnvironment(*bounds, n)
    return env


def gen_evolve_pf(pf):
    """PF:255-267 default scenario at seed 0: records the (never-changing)
    population, the per-generation fitness and best_path()."""
    env = pf_seeded_obstacles(pf, 0, ((0, 20), (0, 20), (0, 10)), 10)
    t0 = time.time()
    ga = pf.GeneticAlgorithm(population_size=50, mutation_rate=0.1, generations=100, environment=env,
                             source=(0, 0, 0), destination=(15, 15, 8))
    init = [list(p.waypoints) for p in ga.population]
    ga.evolve()
    dt = time.time() - t0
    final = [list(p.waypoints) for p in ga.population]
    P = len(final)
    Wmax = max(len(p) for p in final)
    wp = np.zeros((P, Wmax, 3))
    lens = np.zeros(P, np.int32)
    for i, p in enumerate(final):
        lens[i] = len(p)
        wp[i, :len(p)] = np.array(p, np.float64)
    fit = np.array([p.fitness for p in ga.population])
    best = np.array(ga.best_path().waypoints, np.float64)
    units = 100 * int(np.sum(lens - 1))
    meta = dict(META, ref="PF:255-269", seed=0, ref_seconds=dt, units=units, units_per_s=units / dt,
                same_as_init=bool(sorted(map(str, init)) == sorted(map(str, final))))
    np.savez_compressed(os.path.join(HERE, "evolve_pf_default.npz"), wp=wp, lens=lens, fitness=fit, best=best,
                        obs_pos=np.array([o["position"] for o in env.obstacles], np.float64),
                        obs_size=np.array([o["size"] for o in env.obstacles], np.float64),
                        obs_kind=np.array([KINDS.index(o["shape"]) for o in env.obstacles], np.int32),
                        meta=np.array(json.dumps(meta)))
    print(f"evolve_pf_default.npz: {dt:.1f}s reference wall, lens {sorted(set(lens.tolist()))}")


def gen_find_path_pf(pf):
    """PF:93-132 PathPlanner.find_path on seeded PF environments: every dodge strategy, height variations
    -2..1 (what PF:185 draws) and a few larger ones, integer and float endpoints, dense obstacle fields that
    trigger the dodge and the loop guard."""
    cases = []
    scen = [  # (seed, bounds, n_obstacles, source, destination)
        (0, ((0, 20), (0, 20), (0, 10)), 10, (0, 0, 0), (15, 15, 8)),          # PF:256-266 scenario
        (1, ((0, 20), (0, 20), (0, 10)), 60, (0, 0, 0), (19, 19, 9)),          # dense: dodges
        (2, ((0, 12), (0, 12), (0, 6)), 80, (1, 1, 1), (11, 10, 5)),           # very dense: loop guard / stuck
        (3, ((0, 40), (0, 40), (0, 20)), 120, (39, 2, 1), (0, 37, 18)),        # negative directions, long
        (4, ((0, 20), (0, 20), (0, 10)), 30, (3, 3, 3), (3, 3, 3)),            # source == destination
        (5, ((0, 20), (0, 20), (0, 10)), 40, (0.5, 0.25, 0.0), (12.5, 9.25, 6.0)),  # float lattice offsets
    ]
    for seed, bounds, nobs, src, dst in scen:
        env = pf_seeded_obstacles(pf, seed, bounds, nobs)
        planner = pf.PathPlanner(src, dst, env)
        for dodge in ("up", "down", "left", "right", "forward", "backward"):
            for hv in (-2, -1, 0, 1, 3):
                path = planner.find_path(dodge_strategy=dodge, height_variation=hv)
                cases.append(dict(scen=seed, dodge=dodge, hv=hv, path=[[float(v) for v in p] for p in path]))
    Wmax = max(len(c["path"]) for c in cases)
    wp = np.zeros((len(cases), Wmax, 3))
    lens = np.zeros(len(cases), np.int32)
    for i, c in enumerate(cases):
        lens[i] = len(c["path"])
        wp[i, :lens[i]] = np.array(c["path"], np.float64)
    envs = {}
    for seed, bounds, nobs, src, dst in scen:
        env = pf_seeded_obstacles(pf, seed, bounds, nobs)
        envs[f"pos_{seed}"] = np.array([o["position"] for o in env.obstacles], np.float64)
        envs[f"size_{seed}"] = np.array([o["size"] for o in env.obstacles], np.float64)
        envs[f"kind_{seed}"] = np.array([KINDS.index(o["shape"]) for o in env.obstacles], np.int32)
        envs[f"src_{seed}"] = np.array(src, np.float64)
        envs[f"dst_{seed}"] = np.array(dst, np.float64)
    meta = dict(META, ref="PF:93-132", scenarios=[s[0] for s in scen])
    np.savez_compressed(os.path.join(HERE, "find_path_pf.npz"), wp=wp, lens=lens,
                        scen=np.array([c["scen"] for c in cases], np.int32),
                        hv=np.array([c["hv"] for c in cases], np.int32),
                        dodge=np.array([c["dodge"] for c in cases]), meta=np.array(json.dumps(meta)), **envs)
    print(f"find_path_pf.npz: {len(cases)} paths, lens {int(lens.min())}..{int(lens.max())}")


def main():
    hr = ref_loader.load_hr()
    pf = ref_loader.load_pf()
    which = set(sys.argv[1:]) or {"fitness", "evolve", "gae", "find_path", "ops_pf", "lidar", "codecs", "find_path_gae"}
    if "fitness" in which:
        gen_fitness_hr(hr)
        gen_fitness_pf(pf)
    if "evolve" in which:
        rng = np.random.default_rng(7)
        obs_small = hr_obstacles(rng.integers(0, 16, size=(10, 3)), [int(s) for s in rng.integers(1, 5, size=10)])
        # the reference's own defaults (HR:62): pop 50, mut 0.1, gens 100; SURVEY §6 scenario
        gen_evolve_hr(hr, pf, "default", 50, 0.1, 100, (0, 0, 0), (15, 15, 8), obs_small, seed=0)
        # high mutation rate, odd population (exercises the truncated last pair, HR:122)
        gen_evolve_hr(hr, pf, "odd_hot", 21, 0.9, 40, (2, 1, 0), (9, 12, 4), obs_small, seed=3)
        # P < 10 -> top pool is the whole population (HR:118)
        gen_evolve_hr(hr, pf, "tiny", 6, 0.5, 30, (0, 0, 0), (6, 3, 2), obs_small[:3], seed=5)
        # BASELINE configs[0] / SURVEY §8d cfg 1: pop 100, 20 waypoints, 200 generations
        env = pf_seeded_obstacles(pf, 0, ((0, 20), (0, 20), (0, 10)), 10)
        obs_cfg1 = [{"position": [int(v) for v in o["position"]], "size": int(o["size"])} for o in env.obstacles]
        gen_evolve_hr(hr, pf, "cfg1", 100, 0.1, 200, (0, 0, 0), (19, 19, 8), obs_cfg1, seed=0)
        gen_evolve_pf(pf)
    if "gae" in which:
        gen_run_gae()
    if "find_path" in which:
        gen_find_path_pf(pf)
    if "ops_pf" in which:
        gen_ops_pf(pf)
    if "lidar" in which:
        gen_downsample_lidar()
    if "codecs" in which:
        gen_waypoint_codecs()
    if "find_path_gae" in which:
        gen_find_path_gae()




# --------------------------------------------------------------------------- PF crossover / mutate called directly
def gen_ops_pf(pf):
    """GeneticAlgorithm.crossover (PF:196-200) and .mutate (PF:202-208) of the unmodified reference, called
    directly on seeded ragged parents (PF's own evolve drops its children, PF:216-218, so no evolve run can pin
    these two operators).  Records the parents, the np.random draws consumed (as tf_draw records: crossover point,
    gate outcome, mutated waypoint incl. endpoints, per-axis delta in {-1, +1}) and both children."""
    rng = np.random.default_rng(77)
    ga = pf.GeneticAlgorithm.__new__(pf.GeneticAlgorithm)          # PF:175-182 without the PathPlanner population
    n, Wmax = 400, 24
    lens = rng.integers(3, Wmax + 1, size=(n, 2)).astype(np.int32)
    lens[::5, 1] = lens[::5, 0]                                     # equal lengths too
    parents = np.zeros((n, 2, Wmax, 3), np.float64)
    for i in range(n):
        for k in range(2):
            L = lens[i, k]
            parents[i, k, :L] = np.cumsum(rng.integers(-1, 2, size=(L, 3)), axis=0) + rng.integers(0, 20, size=3)
    rates = np.where(np.arange(n) % 3 == 0, 1.0, 0.35)
    tape = np.zeros((n, 8), np.int32)
    clens = np.zeros((n, 2), np.int32)
    children = np.zeros((n, 2, Wmax, 3), np.float64)
    valid = np.zeros(n, bool)
    np.random.seed(123)
    for i in range(n):
        p1 = pf.Path([tuple(int(v) for v in w) for w in parents[i, 0, :lens[i, 0]]])
        p2 = pf.Path([tuple(int(v) for v in w) for w in parents[i, 1, :lens[i, 1]]])
        ga.mutation_rate = float(rates[i])
        with Recorder() as rec:
            c1, c2 = ga.crossover(p1, p2)
            c1 = ga.mutate(c1)
            c2 = ga.mutate(c2)
        ev = iter(rec.events)
        nm, a, k, cp = next(ev)
        assert nm == "randint" and a == (1, len(p1.waypoints) - 1)
        row = [0, 1, int(cp)]
        for c in (c1, c2):
            nm, a, k, gate = next(ev)
            assert nm == "rand"
            if gate < rates[i]:
                nm, a, k, mp = next(ev)
                assert nm == "randint" and a == (0, len(c.waypoints))
                d = []
                for _ in range(3):
                    nm, a, k, v = next(ev)
                    assert nm == "choice"
                    d.append(int(v))
                row += [int(mp), pack_delta(d)]
            else:
                row += [-1, 0]
        assert next(ev, None) is None
        tape[i, :7] = row
        for k, c in enumerate((c1, c2)):
            clens[i, k] = len(c.waypoints)
            children[i, k, :clens[i, k]] = np.array(c.waypoints, np.float64)
        valid[i] = True
    meta = dict(META, ref="PF:196-208", seed=123, note="tape rows: pa=0 pb=1 (the two parents of the row), cp, mp1, d1, mp2, d2; "
                "PF draws cp from randint(1, len(parent1) - 1): a cp beyond parent2's length is legal Python slicing")
    np.savez_compressed(os.path.join(HERE, "ops_pf.npz"), parents=parents.astype(np.int16), lens=lens, rates=rates, tape=tape,
                        children=children.astype(np.int16), clens=clens, valid=valid, meta=np.array(json.dumps(meta)))
    print(f"ops_pf.npz: {int(valid.sum())} pairs, mutated children {(tape[:, 3] >= 0).sum() + (tape[:, 5] >= 0).sum()}")


# --------------------------------------------------------------------------- LiDAR voxel down-sampling
def gen_downsample_lidar():
    """downsample_lidar (UAV_Navigation/data_compression/lidar_compression.py:242-246).  The file is a Colab export
    (`!pip install` at the top) and cannot be imported, so only the source lines of that one function are exec'd —
    unmodified — and called on seeded point clouds."""
    path = os.path.join(ref_loader.REF_ROOT, "UAV_Navigation/data_compression/lidar_compression.py")
    with open(path) as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("def downsample_lidar("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].strip() and not lines[i].startswith((" ", "\t")))
    ns = {"np": np}
    exec(compile("".join(lines[start:end]), path, "exec"), ns)
    fn = ns["downsample_lidar"]
    rng = np.random.default_rng(99)
    out = {}
    cases = [("unit", 1.0, (0.0, 48.0)), ("half", 0.5, (0.0, 20.0)), ("third", 0.3, (0.0, 14.0)), ("negative", 1.0, (-6.0, 30.0)),
             ("coarse", 2.5, (0.0, 100.0))]
    for name, vs, (lo, hi) in cases:
        pts = rng.uniform(lo, hi, size=(5000, 3)).astype(np.float32)   # float32-representable: what the device receives
        pts[:50] = np.round(pts[:50] / vs) * vs                         # points exactly on voxel boundaries
        pts = pts.astype(np.float32)
        vox = fn(pts.astype(np.float64), voxel_size=vs)
        out[f"pts_{name}"] = pts
        out[f"vox_{name}"] = np.asarray(vox, np.float64)
        out[f"vs_{name}"] = np.array(vs)
    meta = dict(META, ref="lidar_compression.py:242-246", cases=[c[0] for c in cases], lines=[start + 1, end])
    np.savez_compressed(os.path.join(HERE, "downsample_lidar.npz"), meta=np.array(json.dumps(meta)), **out)
    print("downsample_lidar.npz:", {c[0]: len(out[f"vox_{c[0]}"]) for c in cases})



# --------------------------------------------------------------------------- waypoint egress codecs
def _exec_functions(path, names, ns):
    """exec the source lines of top-level functions of a Colab export that cannot be imported.  names: (name, line)
    pairs — the file redefines some names several times; `line` (1-based) picks the definition."""
    with open(path) as f:
        lines = f.readlines()
    for name, line in names:
        start = line - 1
        assert lines[start].startswith(f"def {name}("), lines[start]
        end = next(i for i in range(start + 1, len(lines)) if lines[i].strip() and not lines[i].startswith((" ", "\t")))
        exec(compile("".join(lines[start:end]), path, "exec"), ns)
    return ns


def gen_waypoint_codecs():
    """The two lossy front ends the reference puts before its entropy coders
    (UAV_Navigation/data_compression/ga_waypoints_compression.py): compress_waypoints (:2090-2094, quantise to a
    0.05 grid -> float32 -> zlib) and delta_encode + quantize (:875-885, np.diff with a zero first row, rounded to 3
    decimals, packed as doubles), run UNMODIFIED on seeded routes (the file's own generate_route recipe, :2086-2087)
    and on lattice GA paths."""
    import zlib
    path = os.path.join(ref_loader.REF_ROOT, "UAV_Navigation/data_compression/ga_waypoints_compression.py")
    ns = _exec_functions(path, [("compress_waypoints", 2090), ("delta_encode", 875), ("quantize", 882)], {"np": np, "zlib": zlib})
    rng = np.random.default_rng(31)
    routes = {"route": np.cumsum(rng.standard_normal((300, 3)) * 10, axis=0),
              "lattice": np.cumsum(rng.integers(-1, 2, size=(64, 3)), axis=0).astype(np.float64) + 500.0,
              "ties": (rng.integers(-40, 40, size=(200, 3)) * 0.025),          # exact halves of the 0.05 grid
              "short": np.array([[1.0, 2.0, 3.0]])}
    out = {}
    for name, w in routes.items():
        comp = ns["compress_waypoints"](w, 0.05)
        out[f"in_{name}"] = w
        out[f"q32_{name}"] = np.frombuffer(zlib.decompress(comp), dtype=np.float32).reshape(w.shape)
        out[f"zlib_{name}"] = np.frombuffer(comp, dtype=np.uint8)
        out[f"dq_{name}"] = ns["quantize"](ns["delta_encode"](w))
    meta = dict(META, ref="ga_waypoints_compression.py:875-885,2090-2094", cases=list(routes), zlib=zlib.ZLIB_VERSION)
    np.savez_compressed(os.path.join(HERE, "waypoint_codecs.npz"), meta=np.array(json.dumps(meta)), **out)
    print("waypoint_codecs.npz:", {k: v.shape for k, v in routes.items()})



# --------------------------------------------------------------------------- GAE initial paths (find_path)
def gen_find_path_gae():
    """UAVPathPlanner.find_path (GA_enhanced.py GAE:166-179) on the file's own voxel Environment (GAE:87-122), P paths
    per scenario exactly as GeneticAlgorithm.initialize_population builds them (GAE:13-18).  random.shuffle is wrapped
    to record the order every dodge ends up with (GAE:157) — the "shuffle tape"."""
    gae = ref_loader.load_gae_classes()
    moves = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    scen = [  # seed, bounds, random obstacles, source, destination, P
        (11, ((0, 30), (0, 30), (0, 15)), 30, (0, 0, 0), (30, 30, 15), 20),       # the golden GAE run's world
        (12, ((0, 100), (0, 50), (0, 100)), 20, (1, 2, 3), (97, 45, 88), 12),     # SIMR:21 bounds
        (13, ((0, 12), (0, 12), (0, 6)), 200, (0, 0, 0), (12, 12, 6), 16),        # dense: many dodges
        (14, ((0, 30), (0, 30), (0, 15)), 30, (30, 2, 14), (0, 29, 0), 8),        # negative directions
        (15, ((0, 30), (0, 30), (0, 15)), 30, (5, 5, 5), (5, 5, 5), 3),           # source == destination
    ]
    out = {}
    for seed, bounds, nrand, src, dst, P in scen:
        random.seed(seed)
        np.random.seed(seed)
        env = gae.Environment(*bounds, num_random_obstacles=nrand)
        orders = []
        orig = random.shuffle

        def rec(x, _orig=orig):
            _orig(x)
            orders.append(sum(moves.index(m) << (3 * i) for i, m in enumerate(x)))

        random.shuffle = rec
        try:
            paths = [gae.UAVPathPlanner(src, dst, env).find_path() for _ in range(P)]
        finally:
            random.shuffle = orig
        after = random.random()
        Wmax = max(len(p) for p in paths)
        wp = np.zeros((P, Wmax, 3), np.int32)
        lens = np.zeros(P, np.int32)
        for i, pth in enumerate(paths):
            lens[i] = len(pth)
            wp[i, :len(pth)] = np.array(pth, np.int32)
        out.update({f"obs_{seed}": np.array(sorted(env.obstacles), np.int32), f"bounds_{seed}": np.array(bounds, np.int32),
                    f"src_{seed}": np.array(src, np.int32), f"dst_{seed}": np.array(dst, np.int32), f"wp_{seed}": wp,
                    f"lens_{seed}": lens, f"shuffles_{seed}": np.array(orders, np.int32), f"after_{seed}": np.array(after)})
        print(f"find_path_gae scenario {seed}: {P} paths, lens {int(lens.min())}..{int(lens.max())}, {len(orders)} dodges")
    meta = dict(META, ref="GAE:125-179", scenarios=[s[0] for s in scen])
    np.savez_compressed(os.path.join(HERE, "find_path_gae.npz"), meta=np.array(json.dumps(meta)), **out)



# --------------------------------------------------------------------------- GAE (backend GA) run
def gen_run_gae():
    """uav-backend/app/routers/GA_enhanced.py GeneticAlgorithm.run at fixed seeds (GAE:60-66), on its
    own voxel Environment (GAE:87-122).  `mutate` reads a module-global `environment` (GAE:47); the
    class definitions are exec'ed with that global bound to the environment under test."""
    gae = ref_loader.load_gae_classes()
    random.seed(11)
    np.random.seed(11)
    env = gae.Environment((0, 30), (0, 30), (0, 15))
    gae.GeneticAlgorithm.mutate.__globals__["environment"] = env
    ga = gae.GeneticAlgorithm(population_size=20, generations=30, mutation_rate=0.2)
    import contextlib, io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        best = ga.run((0, 0, 0), (30, 30, 15), env)
    bests = [int(l.split(":")[-1]) for l in buf.getvalue().strip().splitlines()]
    final = ga.population
    P = len(final)
    Wmax = max(len(p) for p in final)
    wp = np.zeros((P, Wmax, 3))
    lens = np.zeros(P, np.int32)
    for i, p in enumerate(final):
        lens[i] = len(p)
        wp[i, :len(p)] = np.array(p, np.float64)
    fit = np.array([ga.fitness_function(p) for p in final], np.int64)
    meta = dict(META, ref="GAE:6-66,87-179", seed=11, P=20, G=30, rate=0.2)
    np.savez_compressed(os.path.join(HERE, "run_gae.npz"), wp=wp, lens=lens, fitness=fit, best=np.array(best, np.float64),
                        best_per_generation=np.array(bests, np.int64),
                        obstacles=np.array(sorted(env.obstacles), np.int64), meta=np.array(json.dumps(meta)))
    print("run_gae.npz: final lens", sorted(set(lens.tolist())), "best fitness", bests[-1])


if __name__ == "__main__":
    main()
