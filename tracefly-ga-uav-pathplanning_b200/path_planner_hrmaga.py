"""Drop-in for UAV_Navigation/path_planning_GA/path_planner_hrmaga.py (HR) — same names,
argument order, defaults, return shapes and error behaviour — with the work done on the
B200 through libtracefly_b200.so:

    Path.calculate_fitness       HR:42-59   -> tf_fitness_f64 (reference-order float64 kernel)
    is_collision                 HR:17-23   -> tf_point_query_f64
    distance_to_nearest_obstacle HR:25-28   -> tf_point_query_f64
    HRMAGAGenerator.evolve       HR:101-130 -> tf_rank_topk_f64 + tf_ga_step_f64 + tf_fitness_f64
                                               + tf_track_best_f64 per generation, all on the GPU

Randomness: evolve() consumes np.random in EXACTLY the reference's order (HR:118-120:
choice -> randint -> rand [-> randint -> choice] x2 per child pair) and hands the draws to
the device as a tape, so `np.random.seed(s); HRMAGAGenerator().evolve(...)` returns the same
path as the reference does under the same seed (tests/test_dropin_hr_gpu.py checks it against
goldens recorded from the unmodified reference).

Deliberate, documented deviations:
  * returned waypoints are tuples of plain Python numbers (int when every input was integral,
    else float).  The reference returns NumPy scalars after the first tuple, which makes its
    own json.dump (HR:154) raise TypeError on integer inputs.
  * generate_local_path is ALSO a method of HRMAGAGenerator, because UAV_Navigation/main.py:59
    calls it as one (the reference only defines the module function, HR:138).
"""
import json
import os
from datetime import datetime
from typing import List, Tuple

import numpy as np
import torch

from . import _lib as L
from . import batched as B


def load_obstacles_from_json(filepath: str) -> List[dict]:
    """detected_objects.json -> list of {"position": [...], "size"?: s} (HR:10-15); a missing file means
    an empty world."""
    try:
        with open(filepath, "r") as handle:
            return json.load(handle).get("detected_objects", [])
    except FileNotFoundError:
        return []


def _env_for(obstacles, safety_radius=2.0):
    return B.Environment.from_hr_obstacles(obstacles, safety_radius)


def _pop_from_paths(paths, dtype=torch.float64):
    """list of waypoint lists -> (pop [P, Wmax, 4] cuda, lens i32[P] cuda)"""
    P = len(paths)
    Wmax = max(1, max(len(p) for p in paths))
    host = np.zeros((P, Wmax, 4), np.float64)
    lens = np.zeros(P, np.int32)
    for i, p in enumerate(paths):
        lens[i] = len(p)
        if len(p):
            host[i, :len(p), :3] = np.asarray(p, np.float64).reshape(len(p), -1)[:, :3]
    return torch.from_numpy(host).to("cuda", dtype), torch.from_numpy(lens).cuda()


def is_collision(point, obstacles, safety_radius=2.0):
    """HR:17-23 on the device (tf_point_query_f64)."""
    if not obstacles:
        return False
    coll, _ = B.point_query_f64([tuple(point)], _env_for(obstacles, safety_radius))
    return bool(coll[0])


def distance_to_nearest_obstacle(point, obstacles):
    """HR:25-28 on the device (tf_point_query_f64)."""
    if not obstacles:
        return float("inf")
    _, dn = B.point_query_f64([tuple(point)], _env_for(obstacles))
    return float(dn[0])


class Path:
    def __init__(self, waypoints: List[Tuple[int, int, int]]):
        self.waypoints = waypoints
        self.fitness = 0

    def is_turning(self, i):  # HR:38-40: heading change in the horizontal plane
        here, before = self.waypoints[i], self.waypoints[i - 1]
        return tuple(here[:2]) != tuple(before[:2])

    def calculate_fitness(self, obstacles):  # HR:42-59
        env = obstacles if isinstance(obstacles, B.Environment) else _env_for(obstacles)
        pop, lens = _pop_from_paths([self.waypoints])
        fit, _ = B.fitness_batch_f64(pop, lens, env)
        self.fitness = float(fit[0].item())


def _pythonize(rows, integral):
    out = []
    for r in rows:
        out.append(tuple(int(round(v)) for v in r[:3]) if integral else tuple(float(v) for v in r[:3]))
    return out


def _is_integral(*points):
    return all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for p in points for v in p)


class HRMAGAGenerator:
    def __init__(self, population_size=50, mutation_rate=0.1, generations=100):
        self.population_size = population_size
        self.mutation_rate = mutation_rate
        self.generations = generations
        self.emergency_stop = False

    # ---- HR:68-80
    def generate_initial_path(self, start, end):
        pop, n = B.init_diagonal(start, end, 1, f64=True)
        rows = pop[0, :n].cpu().numpy()
        path = _pythonize(rows, _is_integral(start, end))
        path[0] = start  # HR:69: path = [start]
        return path

    # ---- HR:82-83
    def initialize_population(self, start, end, obstacles):
        init = self.generate_initial_path(start, end)
        return [Path(list(init)) for _ in range(self.population_size)]

    # ---- HR:85-99 for single Path objects (evolve() does the same on the device for the whole
    #      population); both draw from np.random in the reference's order
    def crossover(self, p1, p2):
        a, b = p1.waypoints, p2.waypoints
        cut = np.random.randint(1, min(len(a), len(b)) - 1)
        return Path(a[:cut] + b[cut:]), Path(b[:cut] + a[cut:])

    def mutate(self, path):
        n = len(path.waypoints)
        if n and np.random.rand() < self.mutation_rate:
            at = np.random.randint(1, n - 1)
            delta = np.random.choice([-1, 0, 1], 3)
            path.waypoints[at] = tuple(v + int(d) for v, d in zip(path.waypoints[at], delta))
        return path

    # ---- the np.random draws of one generation, in the reference's order (HR:117-120)
    def _draw_generation(self, P, W, out):
        n_top = min(10, P)
        npairs = out.shape[0]
        rate = self.mutation_rate
        rnd = np.random
        for j in range(npairs):
            pa, pb = rnd.choice(n_top, 2, replace=False)          # HR:118
            cp = rnd.randint(1, W - 1)                            # HR:86
            row = [pa, pb, cp, -1, 0, -1, 0, 0]
            for c in range(2):
                if rnd.rand() < rate:                             # HR:94
                    mp = rnd.randint(1, W - 1)                    # HR:95
                    d = rnd.choice([-1, 0, 1], 3)                 # HR:96
                    row[3 + 2 * c] = mp
                    row[4 + 2 * c] = int((d[0] + 1) | ((d[1] + 1) << 2) | ((d[2] + 1) << 4))
            out[j] = row

    # ---- HR:101-130
    def evolve(self, start, end, obstacles):
        P = self.population_size
        env = obstacles if isinstance(obstacles, B.Environment) else _env_for(obstacles)
        pop, W = B.init_diagonal(start, end, P, f64=True)             # HR:102
        params = L.default_params(dist_zero_guard=1)
        fit, _ = B.fitness_batch_f64(pop, None, env, params)          # HR:103-104
        ga = L.default_ga_params(n_elite=2, n_top=10, selection=L.SEL_TRUNCATION, mutation_rate=self.mutation_rate)
        ne = min(2, P)
        npairs = (P - ne + 1) // 2
        k = min(10, P)
        topk = B.TopK(P, k, f64=True)
        tracker = B.BestTracker(W, f64=True)                          # HR:106-107
        nxt, nfit = torch.empty_like(pop), torch.empty_like(fit)
        top = torch.empty(k, dtype=torch.int32, device="cuda")
        G = self.generations
        tape_host = torch.zeros((max(G, 1), max(npairs, 1), 8), dtype=torch.int32).pin_memory()
        tape_dev = torch.empty_like(tape_host, device="cuda")
        tape_np = tape_host.numpy()
        for gen in range(G):
            if self.emergency_stop:                                   # HR:110-112
                print("[⚠️ EMERGENCY HALT TRIGGERED] Path planning interrupted.")
                return []
            if npairs:
                self._draw_generation(P, W, tape_np[gen, :npairs])
                tape_dev[gen].copy_(tape_host[gen], non_blocking=True)
            topk(fit, out=top)                                        # HR:114
            B.ga_step(pop, fit, top, ga, tape=tape_dev[gen], out=nxt, want_parity=False)   # HR:115-122
            B.fitness_batch_f64(nxt, None, env, params, out=nfit, want_collisions=False)  # HR:123-124
            pop, nxt = nxt, pop
            fit, nfit = nfit, fit
            tracker.update(pop, fit, 0, gen)                          # HR:126-128
        if G == 0:
            # HR:130 dereferences best_path = None when the loop never ran
            raise AttributeError("'NoneType' object has no attribute 'waypoints'")
        rows = tracker.best_path.cpu().numpy()
        path = _pythonize(rows, _is_integral(start, end))
        return path

    def emergency_halt(self):  # HR:132-133
        self.emergency_stop = True

    # MAIN:59 calls ga_system.generate_local_path(...): provided as a method too.
    def generate_local_path(self, start_point, end_point):
        return generate_local_path(start_point, end_point, planner=self)


DETECTIONS_FILE = "data/detected_objects.json"   # written by the LiDAR module (HR:139)
RESULT_FILE = "data/ga_path_output.json"         # read by the mission loop / compression stage (HR:140)


def generate_local_path(start_point, end_point, planner=None):
    """HR:138-160: plan one mission segment.  Obstacles are re-read from DETECTIONS_FILE on every call
    (the reference's notion of a dynamic world); a non-empty result is written to RESULT_FILE with the
    keys path / start / end / timestamp."""
    os.makedirs(os.path.dirname(RESULT_FILE), exist_ok=True)
    planner = planner if planner is not None else HRMAGAGenerator()
    best_path = planner.evolve(start_point, end_point, load_obstacles_from_json(DETECTIONS_FILE))
    if not best_path:
        print("❌ No path generated due to emergency stop or invalid configuration.")
        return best_path
    record = dict(path=best_path, start=start_point, end=end_point, timestamp=datetime.utcnow().isoformat())
    with open(RESULT_FILE, "w") as handle:
        json.dump(record, handle, indent=2)
    print(f"✅ Path saved to {RESULT_FILE}")
    return best_path
