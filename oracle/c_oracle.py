"""ctypes binding of oracle/_build/libtf_oracle.so (built from oracle/tf_oracle.c by oracle/Makefile).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtf_oracle.so")

SEM_NONE, SEM_HR, SEM_PF = 0, 1, 2
KIND_SPHERE, KIND_CUBE, KIND_CYLINDER, KIND_AABB = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "tf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


class Env(C.Structure):
    _fields_ = [("semantics", C.c_int32), ("n", C.c_int32), ("obs", C.c_void_p), ("safety_radius", C.c_float),
                ("n_dyn", C.c_int32), ("dyn", C.c_void_p), ("dyn_safety", C.c_float),
                ("bricks", C.c_void_p), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("hazard", C.c_void_p), ("nt", C.c_int32), ("hny", C.c_int32), ("hnx", C.c_int32),
                ("sx", C.c_float), ("sy", C.c_float), ("ox", C.c_float), ("oy", C.c_float), ("st", C.c_float)]


class Params(C.Structure):
    _fields_ = [("w_safety", C.c_float), ("w_energy", C.c_float), ("w_wind", C.c_float), ("w_time", C.c_float),
                ("w_turn", C.c_float), ("dist_zero_guard", C.c_int32), ("hazard_weight", C.c_float),
                ("samples_per_segment", C.c_int32)]


class GaParams(C.Structure):
    _fields_ = [("n_elite", C.c_int32), ("selection", C.c_int32), ("n_top", C.c_int32), ("tournament_k", C.c_int32),
                ("crossover", C.c_int32), ("mutation", C.c_int32), ("mutation_rate", C.c_float), ("sigma", C.c_float),
                ("stream_id", C.c_uint32), ("_pad", C.c_int32)]


def default_params(guard=1, hazard_weight=0.0, samples=1):
    return Params(0.4, 0.2, 0.2, 0.1, 0.1, guard, hazard_weight, samples)


def default_ga(**kw):
    g = GaParams(2, 0, 10, 4, 0, 0, 0.1, 1.0, 0, 0)
    for k, v in kw.items():
        setattr(g, k, v)
    return g


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_fitness_one_f32.restype = C.c_float
        _lib.orc_fitness_one_f64.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def obstacle_table(pos, size, kind=None, half=None, dtype=np.float32):
    """n x 8 table in tf_obstacle field order.  size NaN/None -> -1 (absent, HR:20)."""
    pos = np.asarray(pos, np.float64).reshape(-1, 3)
    n = len(pos)
    t = np.zeros((n, 8), dtype)
    t[:, 0:3] = pos
    sz = np.asarray(size, np.float64).reshape(-1) if n else np.zeros(0)
    sz = np.where(np.isnan(sz), -1.0, sz)
    t[:, 3] = sz
    if half is not None:
        t[:, 4:7] = half
    k = np.zeros(n, np.int32) if kind is None else np.asarray(kind, np.int32)
    if dtype == np.float32:
        t[:, 7] = k.view(np.float32)
    else:
        t[:, 7] = k
    return np.ascontiguousarray(t)


class OracleEnv:
    """Holds numpy arrays alive and the C struct that points into them."""

    def __init__(self, semantics=SEM_NONE, obs=None, safety_radius=2.0, dyn=None, dyn_safety=2.0,
                 bricks=None, dims=None, hazard=None, hazard_map=None):
        self.semantics = semantics
        self.obs = np.zeros((0, 8), np.float32) if obs is None else np.ascontiguousarray(obs, np.float32)
        self.safety64 = float(safety_radius)
        self.dyn = None if dyn is None else np.ascontiguousarray(dyn, np.float32)
        self.bricks = None if bricks is None else np.ascontiguousarray(bricks, np.uint32)
        self.hazard = None if hazard is None else np.ascontiguousarray(hazard, np.float32)
        e = Env()
        e.semantics = semantics
        e.n = len(self.obs)
        e.obs = self.obs.ctypes.data
        e.safety_radius = safety_radius
        if self.dyn is not None:
            e.n_dyn, e.dyn, e.dyn_safety = len(self.dyn), self.dyn.ctypes.data, dyn_safety
        if self.bricks is not None:
            e.bricks = self.bricks.ctypes.data
            e.nx, e.ny, e.nz = dims
        if self.hazard is not None:
            e.hazard = self.hazard.ctypes.data
            e.nt, e.hny, e.hnx = self.hazard.shape
            e.sx, e.sy, e.ox, e.oy, e.st = hazard_map
        self.c = e
        self.obs64 = None

    def with_obs64(self, obs64):
        self.obs64 = np.ascontiguousarray(obs64, np.float64)
        return self


def set_threads(n):
    lib().orc_set_threads(int(n))


def fitness_f32(env, pop, lens=None, params=None):
    pop = np.ascontiguousarray(pop, np.float32)
    P, W, _ = pop.shape
    params = params or default_params()
    fit = np.zeros(P, np.float32)
    coll = np.zeros((P, (W + 31) // 32), np.uint32)
    lens_c = None if lens is None else np.ascontiguousarray(lens, np.int32)
    lib().orc_fitness_f32(C.byref(env.c), C.byref(params), _p(pop), _p(lens_c), P, W, _p(fit), _p(coll))
    return fit, coll


def fitness_f64(env, pop64, lens=None, params=None, want_dnear=False):
    pop64 = np.ascontiguousarray(pop64, np.float64)
    P, W, _ = pop64.shape
    params = params or default_params()
    fit = np.zeros(P, np.float64)
    coll = np.zeros((P, (W + 31) // 32), np.uint32)
    dn = np.full((P, W), np.inf) if want_dnear else None
    lens_c = None if lens is None else np.ascontiguousarray(lens, np.int32)
    obs64 = env.obs64 if env.obs64 is not None else _widen(env.obs)
    lib().orc_fitness_f64(C.byref(env.c), _p(obs64), C.c_double(env.safety64), C.byref(params), _p(pop64),
                          _p(lens_c), P, W, _p(fit), _p(coll), _p(dn))
    return (fit, coll, dn) if want_dnear else (fit, coll)


def _widen(obs32):
    o = np.zeros(obs32.shape, np.float64)
    o[:, :7] = obs32[:, :7]
    o[:, 7] = np.ascontiguousarray(obs32[:, 7]).view(np.int32)
    return o


def topk(fit, k, worst=False):
    fit = np.ascontiguousarray(fit)
    idx = np.zeros(k, np.int32)
    fn = lib().orc_topk_f64 if fit.dtype == np.float64 else lib().orc_topk_f32
    if fit.dtype not in (np.float32, np.float64):
        raise TypeError(fit.dtype)
    fn(_p(fit), len(fit), k, int(worst), _p(idx))
    return idx


def select_set(fit, k, worst=False):
    """The k best / worst under the stable order, as a set in ascending index order."""
    return np.sort(topk(fit, k, worst=worst)).astype(np.int32)


def philox(key0, key1, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(C.c_uint32(key0), C.c_uint32(key1), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2),
                            C.c_uint32(c3), out)
    return [int(v) for v in out]


def ga_step(pop, fit, top_idx, ga, tape=None, seed=0, gen=0, lens=None):
    """Returns next_pop, next_lens, sel[npairs,2], cp[npairs]; dtype of pop picks the f32/f64 model."""
    is64 = pop.dtype == np.float64
    pop = np.ascontiguousarray(pop)
    fit = np.ascontiguousarray(fit, np.float64 if is64 else np.float32)
    P, W, _ = pop.shape
    ne = min(ga.n_elite, P)
    npairs = (P - ne + 1) // 2
    nxt = np.zeros_like(pop)
    nlens = np.zeros(P, np.int32)
    sel = np.zeros((npairs, 2), np.int32)
    cp = np.zeros(npairs, np.int32)
    lens_c = None if lens is None else np.ascontiguousarray(lens, np.int32)
    tape_c = None if tape is None else np.ascontiguousarray(tape, np.int32)
    top_idx = np.ascontiguousarray(top_idx, np.int32)
    fn = lib().orc_ga_step_f64 if is64 else lib().orc_ga_step_f32
    fn(_p(pop), _p(lens_c), _p(fit), _p(top_idx), _p(tape_c), C.c_uint64(seed), C.c_uint32(gen), C.byref(ga), P, W,
       _p(nxt), _p(nlens), _p(sel), _p(cp))
    return nxt, nlens, sel, cp


def fitness_gae(pop, lens=None):
    pop = np.ascontiguousarray(pop, np.float32)
    P, W, _ = pop.shape
    fit = np.zeros(P, np.int32)
    lens_c = None if lens is None else np.ascontiguousarray(lens, np.int32)
    lib().orc_fitness_gae(_p(pop), _p(lens_c), P, W, _p(fit))
    return fit


def roulette_cumweights(score):
    """GAE:28-34 / random.choices: cumulative weights s / total, accumulated left to right in float64."""
    score = np.ascontiguousarray(score, np.int32)
    cum = np.zeros(len(score), np.float64)
    lib().orc_roulette_cumweights(_p(score), len(score), _p(cum))
    return cum


def select_roulette(cum, u):
    """bisect_right(cum, u * (cum[-1] + 0.0), 0, n - 1) for every uniform draw u."""
    cum = np.ascontiguousarray(cum, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    idx = np.zeros(len(u), np.int32)
    lib().orc_select_roulette(_p(cum), len(cum), _p(u), len(u), _p(idx))
    return idx


def voxelize_points(pts, dims, origin=(0.0, 0.0, 0.0), voxel_size=1.0):
    """downsample_lidar's voxel assignment (lidar_compression.py:242-246) as occupancy bits in brick layout."""
    pts = np.ascontiguousarray(pts, np.float32)
    nx, ny, nz = dims
    bricks = np.zeros(voxel_words(nx, ny, nz), np.uint32)
    o = (C.c_double * 3)(*[float(v) for v in origin])
    lib().orc_voxelize_points(_p(pts), C.c_int64(len(pts)), o, C.c_double(float(voxel_size)), _p(bricks), nx, ny, nz)
    return bricks


def encode_waypoints(paths, mode, param):
    """paths f64[n, W, 4] -> mode 0: f32[n, W, 3] quantised to a `param` grid (ga_waypoints_compression.py:2090-2093);
    mode 1: f64[n, W, 3] deltas rounded to `param` decimals (:875-884)."""
    paths = np.ascontiguousarray(paths, np.float64)
    n, W, _ = paths.shape
    out = np.zeros((n, W, 3), np.float32 if mode == 0 else np.float64)
    lib().orc_encode_waypoints(_p(paths), C.c_int64(n), W, int(mode), C.c_double(float(param)), _p(out))
    return out


def find_path_gae(obstacles, bounds, src, dst, shuffles, n_paths, cap=512):
    """UAVPathPlanner.find_path (GAE:166-179) for n_paths in sequence; shuffles: packed random.shuffle orders.
    -> (paths i32[n, cap, 4], lens i32[n], consumed, status)"""
    obs = np.zeros((len(obstacles), 4), np.int32)
    if len(obstacles):
        obs[:, :3] = np.asarray(obstacles, np.int32).reshape(len(obstacles), -1)[:, :3]
    b = np.ascontiguousarray(np.asarray(bounds, np.int32).reshape(6))
    s3, d3 = np.asarray(src, np.int32), np.asarray(dst, np.int32)
    sh = np.ascontiguousarray(shuffles, np.int32)
    paths = np.zeros((n_paths, cap, 4), np.int32)
    lens = np.zeros(n_paths, np.int32)
    used, status = C.c_int32(0), C.c_int32(0)
    lib().orc_find_path_gae(_p(obs), len(obs), _p(b), _p(s3), _p(d3), _p(sh), len(sh), n_paths, cap, _p(paths), _p(lens),
                            C.byref(used), C.byref(status))
    return paths, lens, used.value, status.value


def voxel_words(nx, ny, nz):
    return (nx // 8) * (ny // 8) * (nz // 8) * 16


def voxelize(env, nx, ny, nz):
    bricks = np.zeros(voxel_words(nx, ny, nz), np.uint32)
    lib().orc_voxelize(C.byref(env.c), _p(bricks), nx, ny, nz)
    return bricks


def voxel_pack_linear(lin, nx, ny, nz):
    lin = np.ascontiguousarray(lin, np.uint32)
    bricks = np.zeros(voxel_words(nx, ny, nz), np.uint32)
    lib().orc_voxel_pack_linear(_p(lin), _p(bricks), nx, ny, nz)
    return bricks


def voxel_test(bricks, dims, x, y, z):
    return int(lib().orc_voxel_test(_p(bricks), dims[0], dims[1], dims[2], int(x), int(y), int(z)))


def init_diagonal(start, end, wcap=0):
    s = (C.c_double * 3)(*[float(v) for v in start])
    e = (C.c_double * 3)(*[float(v) for v in end])
    n = lib().orc_init_diagonal(s, e, 0, None)
    out = np.zeros((max(n, wcap), 4), np.float64)
    lib().orc_init_diagonal(s, e, len(out), _p(out))
    return out[:n]


DODGE_MOVES = {"up": (0, 0, 1), "down": (0, 0, -1), "left": (-1, 0, 0), "right": (1, 0, 0), "forward": (0, 1, 0),
               "backward": (0, -1, 0)}


def find_path_pf(env, source, destination, height_variations, dodge="up", cap=256):
    """PF:93-132 for a batch of height variations -> (paths f64[n, cap, 4], lens i32[n], overflow count)."""
    hv = np.ascontiguousarray(height_variations, np.int32)
    n = len(hv)
    paths = np.zeros((n, cap, 4), np.float64)
    lens = np.zeros(n, np.int32)
    over = np.zeros(1, np.int32)
    d3 = lambda v: (C.c_double * 3)(*[float(x) for x in v])
    obs64 = env.obs64 if env.obs64 is not None else _widen(env.obs)
    lib().orc_find_path_pf(C.byref(env.c), _p(obs64), C.c_double(env.safety64), d3(source), d3(destination), _p(hv), n,
                           d3(DODGE_MOVES[dodge]), cap, _p(paths), _p(lens), _p(over))
    return paths, lens, int(over[0])


def init_walks(seed, P, W, start, lo, hi, start_pp=None):
    pop = np.zeros((P, W, 4), np.float32)
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    spp = None if start_pp is None else np.ascontiguousarray(start_pp, np.float32)
    lib().orc_init_walks(C.c_uint64(seed), _p(spp), f3(start), f3(lo), f3(hi), P, W, _p(pop))
    return pop


def pack_elites(pop, fit, idx):
    pop = np.ascontiguousarray(pop, np.float32)
    fit = np.ascontiguousarray(fit, np.float32)
    idx = np.ascontiguousarray(idx, np.int32)
    E, W = len(idx), pop.shape[1]
    rows = np.zeros((E, W, 4), np.float32)
    rfit = np.zeros(E, np.float32)
    lib().orc_pack_elites(_p(pop), _p(fit), _p(idx), E, W, _p(rows), _p(rfit))
    return rows, rfit


def replace_worst(pop, fit, worst, grows, gfit, world, self_rank, E):
    """In place on pop / fit (must be contiguous float32)."""
    assert pop.flags.c_contiguous and fit.flags.c_contiguous and pop.dtype == np.float32 and fit.dtype == np.float32
    worst = np.ascontiguousarray(worst, np.int32)
    grows = np.ascontiguousarray(grows, np.float32)
    gfit = np.ascontiguousarray(gfit, np.float32)
    lib().orc_replace_worst(_p(pop), _p(fit), _p(worst), pop.shape[1], _p(grows), _p(gfit), world, self_rank, E)
