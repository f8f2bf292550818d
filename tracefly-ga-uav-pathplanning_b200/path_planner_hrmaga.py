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

Speed: the draws of a block of generations are produced natively (tf_np_hr_tape restates NumPy's legacy MT19937
methods bit for bit and is checked against np.random before use) and the generations of a block replay as one CUDA
graph, so BASELINE configs[0] (100 paths, 200 generations) takes milliseconds instead of the reference's 30-50 s.

Deliberate, documented deviations:
  * returned waypoints are tuples of plain Python numbers (int when every input was integral,
    else float).  The reference returns NumPy scalars after the first tuple, which makes its
    own json.dump (HR:154) raise TypeError on integer inputs.
  * a segment longer than TF_W_MAX - 1 = 255 lattice steps raises ValueError (the reference would plan it; main.py
    plans 5-unit segments, MAIN:43-56).  A population of one path returns that path, as the reference does.
  * generate_local_path is ALSO a method of HRMAGAGenerator, because UAV_Navigation/main.py:59
    calls it as one (the reference only defines the module function, HR:138).
"""
import ctypes as C
import json
import os
from datetime import datetime
from typing import List, Tuple

import numpy as np
import torch

from . import _lib as L
from . import batched as B


_NATIVE_TAPE = None


def _native_tape_ok():
    """True when tf_np_hr_tape reproduces np.random's legacy generator on a probe (40 generations x 7 pairs, all
    four draw kinds, from a scratch RandomState): guards against a NumPy whose legacy streams differ from the
    restated ones (2.3.x).  Checked once per process; the global generator is not touched."""
    global _NATIVE_TAPE
    if _NATIVE_TAPE is None:
        try:
            rs = np.random.RandomState(20240229)
            st = rs.get_state()
            key = np.ascontiguousarray(st[1], np.uint32).copy()
            pos = C.c_int32(int(st[2]))
            got = np.zeros((40, 7, 8), np.int32)
            rc = L.lib.tf_np_hr_tape(key.ctypes.data_as(C.c_void_p), C.byref(pos), 40, 7, 6, 9, 0.4, got.ctypes.data_as(C.c_void_p))
            want = np.zeros_like(got)
            for g in range(40):
                for j in range(7):
                    pa, pb = rs.choice(6, 2, replace=False)
                    row = [pa, pb, rs.randint(1, 8), -1, 0, -1, 0, 0]
                    for c in range(2):
                        if rs.rand() < 0.4:
                            row[3 + 2 * c] = rs.randint(1, 8)
                            d = rs.choice([-1, 0, 1], 3)
                            row[4 + 2 * c] = int((d[0] + 1) | ((d[1] + 1) << 2) | ((d[2] + 1) << 4))
                    want[g, j] = row
            after = rs.get_state()
            _NATIVE_TAPE = bool(rc == 0 and np.array_equal(got, want) and np.array_equal(after[1], key) and int(after[2]) == pos.value)
        except Exception:
            _NATIVE_TAPE = False
    return _NATIVE_TAPE


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

    def _draw_block(self, P, W, out):
        """The draws of out.shape[0] consecutive generations into out[g, :npairs] (int32 [g, npairs, 8]), consuming
        np.random's global generator exactly as the reference does.  Fast path: the native restatement of NumPy's
        legacy generator (tf_np_hr_tape, ~1 us per child pair) advances a copy of the state, which is then
        installed with np.random.set_state; used only after it has reproduced np.random itself on a probe."""
        n_top = min(10, P)
        G, npairs = out.shape[0], out.shape[1]
        if npairs and W >= 3 and n_top >= 2 and _native_tape_ok():
            st = np.random.get_state()
            key = np.ascontiguousarray(st[1], np.uint32).copy()
            pos = C.c_int32(int(st[2]))
            L.check(L.lib.tf_np_hr_tape(key.ctypes.data_as(C.c_void_p), C.byref(pos), G, npairs, n_top, W,
                                        float(self.mutation_rate), out.ctypes.data_as(C.c_void_p)), "tf_np_hr_tape")
            np.random.set_state((st[0], key, pos.value, st[3], st[4]))
            return
        for g in range(G):
            self._draw_generation(P, W, out[g])

    # ---- HR:101-130
    def evolve(self, start, end, obstacles):
        P = self.population_size
        env = obstacles if isinstance(obstacles, B.Environment) else _env_for(obstacles)
        pop, W = B.init_diagonal(start, end, P, f64=True)             # HR:102
        if W > L.W_MAX:
            raise ValueError(f"start and end are {W - 1} lattice steps apart; this drop-in plans segments of at most "
                             f"{L.W_MAX - 1} steps (TF_W_MAX waypoints per chromosome) — split the segment as main.py does")
        params = L.default_params(dist_zero_guard=1)
        fit, _ = B.fitness_batch_f64(pop, None, env, params)          # HR:103-104
        G = self.generations
        tracker = B.BestTracker(W, f64=True)                          # HR:106-107
        if P < 2:
            # one path: no parents to pick (HR:117 never runs), every generation re-evaluates the same path and the
            # first one records it as best (HR:126-128); tf_ga_step needs two ranked parents, so this stays here
            for gen in range(G):
                if self.emergency_stop:
                    print("[⚠️ EMERGENCY HALT TRIGGERED] Path planning interrupted.")
                    return []
                tracker.update(pop, fit, 0, gen)
        else:
            halted = self._evolve_device(pop, fit, env, params, tracker, P, W, G)
            if halted:
                print("[⚠️ EMERGENCY HALT TRIGGERED] Path planning interrupted.")
                return []
        if G == 0:
            # HR:130 dereferences best_path = None when the loop never ran
            raise AttributeError("'NoneType' object has no attribute 'waypoints'")
        rows = tracker.best_path.cpu().numpy()
        path = _pythonize(rows, _is_integral(start, end))
        return path

    def _evolve_device(self, pop, fit, env, params, tracker, P, W, G):
        """HR:109-128 on the device.  The draws never depend on fitness (only on P, W and the mutation rate), so they
        are produced a block of K generations ahead; the K generations of a block — 4 kernels each: rank, select /
        crossover / mutate, fitness, best tracking — are captured once in a CUDA graph and replayed per block, so the
        loop is no longer bound by Python + launch latency (BASELINE configs[0]: 802 launches, 0.104 s -> 4 eager
        generations + 13 graph replays).  The emergency flag (HR:110-112) is polled before every block, i.e. within
        K <= 16 generations; the np.random state is then rewound to what the generations actually run consumed.
        Returns True when halted."""
        ga = L.default_ga_params(n_elite=2, n_top=10, selection=L.SEL_TRUNCATION, mutation_rate=self.mutation_rate)
        ne = min(2, P)
        npairs = (P - ne + 1) // 2
        k = min(10, P)
        topk = B.TopK(P, k, f64=True)
        nxt, nfit = torch.empty_like(pop), torch.empty_like(fit)
        top = torch.empty(k, dtype=torch.int32, device="cuda")
        # block size: even (the ping-pong ends where it started), <= 16, and a tape block of at most 16 MB
        K = max(1, min(16, G, (16 << 20) // max(1, npairs * 32)))
        K = K - (K & 1) if K > 1 else 1
        rows = max(npairs, 1)
        tape_host = [torch.zeros((K, rows, 8), dtype=torch.int32).pin_memory() for _ in range(2)]
        host_free = [None, None]
        tape_dev = torch.empty((K, rows, 8), dtype=torch.int32, device="cuda")
        state = {"pop": pop, "nxt": nxt, "fit": fit, "nfit": nfit}

        def generation(i):                                            # reads tape_dev[i]
            topk(state["fit"], out=top)                                                                   # HR:114
            B.ga_step(state["pop"], state["fit"], top, ga, tape=tape_dev[i], out=state["nxt"], want_parity=False)   # HR:115-122
            B.fitness_batch_f64(state["nxt"], None, env, params, out=state["nfit"], want_collisions=False)  # HR:123-124
            state["pop"], state["nxt"] = state["nxt"], state["pop"]
            state["fit"], state["nfit"] = state["nfit"], state["fit"]
            tracker.update(state["pop"], state["fit"], 0, 0)          # HR:126-128

        rng0 = np.random.get_state()
        graph, done, block = None, 0, 0
        use_graph = K >= 2 and G >= 3 * K and os.environ.get("TRACEFLY_HR_GRAPH", "1") != "0"
        while done < G:
            if self.emergency_stop:                                   # HR:110-112
                if npairs:   # the reference would have drawn for `done` generations only
                    np.random.set_state(rng0)
                    scratch = np.zeros((done, npairs, 8), np.int32)
                    self._draw_block(P, W, scratch)
                return True
            n = min(K, G - done)
            hb = block & 1
            if host_free[hb] is not None:
                host_free[hb].synchronize()                           # the H2D copy that read this buffer has finished
            if npairs:
                self._draw_block(P, W, tape_host[hb].numpy()[:n, :npairs])
                tape_dev[:n].copy_(tape_host[hb][:n], non_blocking=True)
                host_free[hb] = torch.cuda.Event()
                host_free[hb].record()
            if use_graph and n == K and block >= 1:
                if graph is None:                                     # block 0 ran eagerly: every kernel is loaded
                    graph = torch.cuda.CUDAGraph()
                    torch.cuda.synchronize()
                    with torch.cuda.graph(graph):
                        for i in range(K):
                            generation(i)
                graph.replay()
            else:
                for i in range(n):
                    generation(i)
            done += n
            block += 1
        torch.cuda.current_stream().synchronize()
        return False

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
