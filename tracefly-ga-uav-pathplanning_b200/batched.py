"""Array-form entry points of the GA hot path ("chromosome-as-waypoint-array in, fitness vector
and next population out", BASELINE.json north_star).  torch is used for device memory and
streams only; every computation is a call through the C-ABI of libtracefly_b200.so.

Layouts: population f32[P, W, 4] (x, y, z, pad) on the GPU, fitness f32[P],
collision flags u32[P, ceil(W/32)] (returned as int32 tensors; bit i of word k = waypoint 32k+i).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._lib import check, lib

KINDS = {"sphere": L.KIND_SPHERE, "cube": L.KIND_CUBE, "cylinder": L.KIND_CYLINDER, "aabb": L.KIND_AABB}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tracefly_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk_tensor(t, dtype, name, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims")


class Environment:
    """Device-side obstacle set: analytic obstacles (HR or PF rules), optional moving spheres,
    optional bit-packed voxel grid, optional hazard heat-map.  Owns a tf_env."""

    def __init__(self, semantics=L.SEM_NONE, table64=None, safety_radius=2.0):
        _require_cuda()
        self.semantics = semantics
        self.safety_radius = float(safety_radius)
        t = np.zeros((0, 8), np.float64) if table64 is None else np.ascontiguousarray(table64, np.float64)
        self.table64 = t
        self._h = C.c_void_p()
        check(lib.tf_env_create_f64(semantics, t.ctypes.data_as(C.c_void_p), len(t), self.safety_radius,
                                    C.byref(self._h)), "tf_env_create_f64")
        self._keep = {}

    # ---- constructors mirroring the reference's obstacle formats
    @classmethod
    def from_hr_obstacles(cls, obstacles, safety_radius=2.0):
        """obstacles: list of {"position": [x,y,z], "size": s?} (detected_objects.json, HR:10-15).
        A missing size stays missing (-1) and takes the HR:20 default 1.0 on the device."""
        t = np.zeros((len(obstacles), 8), np.float64)
        for i, o in enumerate(obstacles):
            t[i, 0:3] = np.asarray(o["position"], np.float64)[:3]
            t[i, 3] = float(o["size"]) if "size" in o else -1.0
        return cls(L.SEM_HR, t, safety_radius)

    @classmethod
    def from_hr_obstacles_indexed(cls, obstacles, box_lo, box_hi, cell, safety_radius=2.0, cap=64):
        """HR rules (HR:17-28) for LARGE sphere sets (up to 2048) through the candidate index: the
        spheres are stored as moving spheres with zero velocity, so a waypoint only visits the few
        spheres that can be nearest to / collide with a point of its grid cell.  The arithmetic per
        sphere is the same (c + 0*t == c), hence results are bit-identical to `from_hr_obstacles`
        for the FP32 kernels (the float64 reference-order kernels need the plain constructor)."""
        env = cls(L.SEM_NONE, None, 0.0)
        env.semantics = L.SEM_HR            # callers pick the HR distance guard from this
        dyn = np.zeros((len(obstacles), 8), np.float32)
        for i, o in enumerate(obstacles):
            dyn[i, 0:3] = np.asarray(o["position"], np.float32)[:3]
            dyn[i, 3] = float(o["size"]) if "size" in o else -1.0
        env.set_dynamic(dyn, safety_radius)
        env.build_dynamic_index(L.W_MAX, box_lo, box_hi, cell, cap=cap, band=L.W_MAX)
        return env

    @classmethod
    def from_pf_obstacles(cls, obstacles):
        """obstacles: list of {"shape", "position", "size"} (ComplexEnvironment.obstacles, PF:17-26);
        an "aabb" shape with "half": (hx,hy,hz) is accepted as an extension."""
        t = np.zeros((len(obstacles), 8), np.float64)
        for i, o in enumerate(obstacles):
            t[i, 0:3] = np.asarray(o["position"], np.float64)[:3]
            t[i, 3] = float(o["size"])
            t[i, 7] = KINDS[o["shape"]]
            if o["shape"] == "aabb":
                t[i, 4:7] = np.asarray(o["half"], np.float64)
        return cls(L.SEM_PF, t, 0.0)

    @classmethod
    def empty(cls):
        return cls(L.SEM_NONE, None, 0.0)

    # ---- optional attachments
    def set_dynamic(self, dyn, safety_radius=2.0):
        """dyn: float array [n, 7|8] = cx cy cz size vx vy vz — moving spheres (extension)."""
        d = np.zeros((len(dyn), 8), np.float32)
        d[:, :7] = np.asarray(dyn, np.float32)[:, :7]
        check(lib.tf_env_set_dynamic(self._h, d.ctypes.data_as(C.c_void_p), len(d), float(safety_radius)),
              "tf_env_set_dynamic")
        return self

    def build_static_index(self, box_lo, box_hi, cell, cap=64, z_cell=1.0, z_cap=32):
        """Candidate index for the static analytic obstacles (HR or PF rules): same results, a waypoint
        visits tens of obstacles instead of all N.  See tf_env_build_static_index."""
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        check(lib.tf_env_build_static_index(self._h, f3(box_lo), f3(box_hi), float(cell), int(cap), float(z_cell),
                                            int(z_cap), _stream()), "tf_env_build_static_index")
        return self

    def static_index_stats(self):
        n, over, mean, mx, zmean = C.c_int64(), C.c_int64(), C.c_double(), C.c_int(), C.c_double()
        check(lib.tf_env_static_index_stats(self._h, C.byref(n), C.byref(over), C.byref(mean), C.byref(mx), C.byref(zmean)),
              "tf_env_static_index_stats")
        return dict(lists=n.value, overflowed=over.value, mean_len=mean.value, max_len=mx.value, z_mean_len=zmean.value)

    def build_dynamic_index(self, time_slices, box_lo, box_hi, cell, cap=48, band=1):
        """Candidate index for the moving spheres (same results, far fewer sphere tests):
        waypoint i at integer time i < time_slices inside [box_lo, box_hi) only visits the spheres
        that can be nearest to / collide with a point of its cell."""
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        check(lib.tf_env_build_dynamic_index(self._h, int(time_slices), int(band), f3(box_lo), f3(box_hi), float(cell),
                                             int(cap), _stream()), "tf_env_build_dynamic_index")
        return self

    def dynamic_index_stats(self):
        n, over, mean, mx = C.c_int64(), C.c_int64(), C.c_double(), C.c_int()
        check(lib.tf_env_dynamic_index_stats(self._h, C.byref(n), C.byref(over), C.byref(mean), C.byref(mx)),
              "tf_env_dynamic_index_stats")
        return dict(lists=n.value, overflowed=over.value, mean_len=mean.value, max_len=mx.value)

    def set_voxels(self, bricks, dims):
        """bricks: int32/uint32 CUDA tensor in brick layout (see voxelize / voxel_pack_linear)."""
        if bricks is None:
            check(lib.tf_env_set_voxels(self._h, None, 0, 0, 0), "tf_env_set_voxels")
            self._keep.pop("bricks", None)
            return self
        nx, ny, nz = dims
        need = lib.tf_voxel_grid_bytes(nx, ny, nz)
        if need < 0 or bricks.numel() * bricks.element_size() != need or not bricks.is_cuda:
            raise ValueError("bricks must be a CUDA tensor of tf_voxel_grid_bytes(nx,ny,nz) bytes")
        check(lib.tf_env_set_voxels(self._h, _ptr(bricks), nx, ny, nz), "tf_env_set_voxels")
        self._keep["bricks"] = bricks
        return self

    def set_hazard(self, hazard, sx=1.0, sy=1.0, ox=0.0, oy=0.0, st=1.0):
        """hazard: f32[nt, ny, nx] CUDA tensor; waypoint i at (x, y) samples (x*sx+ox, y*sy+oy, i*st)."""
        if hazard is None:
            check(lib.tf_env_set_hazard(self._h, None, 0, 0, 0, 1, 1, 0, 0, 1), "tf_env_set_hazard")
            self._keep.pop("hazard", None)
            return self
        _chk_tensor(hazard, torch.float32, "hazard", 3)
        nt, ny, nx = hazard.shape
        check(lib.tf_env_set_hazard(self._h, _ptr(hazard), nt, ny, nx, sx, sy, ox, oy, st), "tf_env_set_hazard")
        self._keep["hazard"] = hazard
        return self

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.tf_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------ voxels
def voxel_grid_words(nx, ny, nz):
    b = lib.tf_voxel_grid_bytes(nx, ny, nz)
    if b < 0:
        raise ValueError("voxel grid dims must be positive multiples of 8")
    return b // 4


def obstacle_records(table):
    """float table [n, 8] (cx cy cz size hx hy hz kind) -> tf_obstacle records as a uint8 CPU tensor."""
    t = np.asarray(table, np.float64)
    rec = np.zeros((len(t), 8), np.float32)
    rec[:, :7] = t[:, :7]
    rec[:, 7] = t[:, 7].astype(np.int32).view(np.float32)
    return torch.from_numpy(rec)


def voxelize(table, dims, semantics=L.SEM_PF, safety_radius=0.0, out=None):
    """Rasterise analytic obstacles into a brick-layout occupancy grid on the GPU."""
    _require_cuda()
    nx, ny, nz = dims
    words = voxel_grid_words(nx, ny, nz)
    if out is None:
        out = torch.zeros(words, dtype=torch.int32, device="cuda")
    rec = obstacle_records(table).cuda()
    check(lib.tf_voxelize(semantics, _ptr(rec), rec.shape[0], float(safety_radius), _ptr(out), nx, ny, nz, _stream()),
          "tf_voxelize")
    return out


def voxelize_points(points, dims, origin=(0.0, 0.0, 0.0), voxel_size=1.0, out=None):
    """LiDAR point cloud (CUDA tensor f32[n, 4]) -> brick-layout occupancy grid; may be called repeatedly
    on the same grid to accumulate scans."""
    _chk_tensor(points, torch.float32, "points", 2)
    nx, ny, nz = dims
    if out is None:
        out = torch.zeros(voxel_grid_words(nx, ny, nz), dtype=torch.int32, device=points.device)
    o = (C.c_double * 3)(*[float(v) for v in origin])
    check(lib.tf_voxelize_points(_ptr(points), points.shape[0], o, float(voxel_size), _ptr(out), nx, ny, nz, _stream()),
          "tf_voxelize_points")
    return out


def voxel_pack_linear(linear, dims):
    """linear: int32 CUDA tensor, bit index x + nx*(y + ny*z) -> brick layout."""
    nx, ny, nz = dims
    out = torch.empty(voxel_grid_words(nx, ny, nz), dtype=torch.int32, device="cuda")
    check(lib.tf_voxel_pack_linear(_ptr(linear), _ptr(out), nx, ny, nz, _stream()), "tf_voxel_pack_linear")
    return out


# ------------------------------------------------------------------------------------ fitness
def fitness_batch(pop, lens, env, params=None, out=None, want_collisions=True):
    """Path.calculate_fitness (HR:42-59 / PF:144-170) for a whole population (FP32 kernel).
    pop f32[P,W,4] cuda; lens i32[P] cuda or None.  -> (fitness f32[P], collide i32[P, ceil(W/32)])"""
    _chk_tensor(pop, torch.float32, "pop", 3)
    P, W, four = pop.shape
    if four != 4:
        raise ValueError("pop must be [P, W, 4]")
    if lens is not None:
        _chk_tensor(lens, torch.int32, "lens", 1)
    params = params or L.default_params(dist_zero_guard=1 if env.semantics != L.SEM_PF else 0)
    fit = out if out is not None else torch.empty(P, dtype=torch.float32, device=pop.device)
    coll = torch.empty((P, (W + 31) // 32), dtype=torch.int32, device=pop.device) if want_collisions else None
    if P == 0:
        return fit, coll
    check(lib.tf_fitness(env.handle, _ptr(pop), _ptr(lens), P, W, C.byref(params), _ptr(fit), _ptr(coll), _stream()),
          "tf_fitness")
    return fit, coll


def fitness_batch_f64(pop64, lens, env, params=None, out=None, want_collisions=True):
    """Reference-order float64 fitness: bit-identical to the reference on identical inputs."""
    _chk_tensor(pop64, torch.float64, "pop64", 3)
    P, W, _ = pop64.shape
    if lens is not None:
        _chk_tensor(lens, torch.int32, "lens", 1)
    params = params or L.default_params(dist_zero_guard=1 if env.semantics != L.SEM_PF else 0)
    fit = out if out is not None else torch.empty(P, dtype=torch.float64, device=pop64.device)
    coll = torch.empty((P, (W + 31) // 32), dtype=torch.int32, device=pop64.device) if want_collisions else None
    check(lib.tf_fitness_f64(env.handle, _ptr(pop64), _ptr(lens), P, W, C.byref(params), _ptr(fit), _ptr(coll),
                             _stream()), "tf_fitness_f64")
    return fit, coll


def fitness_gae(pop, lens):
    """Backend GA fitness (GAE:20-26) for a ragged population: -> int32[P] cuda."""
    _chk_tensor(pop, torch.float32, "pop", 3)
    P, W, _ = pop.shape
    if lens is not None:
        _chk_tensor(lens, torch.int32, "lens", 1)
    fit = torch.empty(P, dtype=torch.int32, device=pop.device)
    if P:
        check(lib.tf_fitness_gae(_ptr(pop), _ptr(lens), P, W, _ptr(fit), _stream()), "tf_fitness_gae")
    return fit


def roulette_cumweights(score):
    """Cumulative roulette weights of the backend GA (GAE:28-34) for int32 scores on the device -> f64[P] cuda,
    accumulated left to right exactly like random.choices does."""
    _chk_tensor(score, torch.int32, "score", 1)
    cum = torch.empty(score.numel(), dtype=torch.float64, device=score.device)
    check(lib.tf_roulette_cumweights(_ptr(score), score.numel(), _ptr(cum), _stream()), "tf_roulette_cumweights")
    return cum


def select_roulette(cum, u):
    """random.choices' pick for every uniform draw u (f64 cuda tensor): bisect_right(cum, u * total) -> i32[n]."""
    _chk_tensor(cum, torch.float64, "cum", 1)
    _chk_tensor(u, torch.float64, "u", 1)
    idx = torch.empty(u.numel(), dtype=torch.int32, device=cum.device)
    check(lib.tf_select_roulette(_ptr(cum), cum.numel(), _ptr(u), u.numel(), _ptr(idx), _stream()), "tf_select_roulette")
    return idx


def point_query_f64(points, env):
    """is_collision / distance_to_nearest_obstacle (HR:17-28) or is_obstacle / distance_to_obstacles
    (PF:28-59) for a batch of points, float64 reference arithmetic.  points: array-like [n, 3].
    -> (collide bool[n] numpy, dnear f64[n] numpy)"""
    _require_cuda()
    pts = np.zeros((len(points), 4), np.float64)
    if len(points):
        pts[:, :3] = np.asarray(points, np.float64).reshape(len(points), -1)[:, :3]
    d = torch.from_numpy(pts).cuda()
    n = len(pts)
    coll = torch.empty(n, dtype=torch.int32, device="cuda")
    dn = torch.empty(n, dtype=torch.float64, device="cuda")
    check(lib.tf_point_query_f64(env.handle, _ptr(d), n, _ptr(coll), _ptr(dn), _stream()), "tf_point_query_f64")
    return coll.cpu().numpy().astype(bool), dn.cpu().numpy()


def find_path_pf(env, source, destination, height_variations, dodge="up", cap=64, max_cap=1 << 16):
    """`PathPlanner.find_path` (PF:93-132) for a batch of height variations: one launch, one thread per path.
    -> (paths f64[n, cap, 4] cuda, lens i32[n] cuda).  The capacity is doubled (and the launch repeated)
    until every path fits."""
    _require_cuda()
    hv = torch.as_tensor(np.ascontiguousarray(height_variations, np.int32)).cuda()
    n = hv.numel()
    src = (C.c_double * 3)(*[float(v) for v in source])
    dst = (C.c_double * 3)(*[float(v) for v in destination])
    over = torch.zeros(1, dtype=torch.int32, device="cuda")
    while True:
        paths = torch.empty((n, cap, 4), dtype=torch.float64, device="cuda")
        lens = torch.empty(n, dtype=torch.int32, device="cuda")
        over.zero_()
        check(lib.tf_find_path_pf(env.handle, src, dst, _ptr(hv), n, L.DODGE[dodge], cap, _ptr(paths), _ptr(lens), _ptr(over),
                                  _stream()), "tf_find_path_pf")
        if n == 0 or int(over.item()) == 0:
            return paths, lens
        if cap >= max_cap:
            raise RuntimeError(f"find_path_pf: a path needs more than {max_cap} waypoints")
        cap *= 2


def find_path_gae(obstacles, bounds, source, destination, n_paths, cap=256, max_cap=1 << 16):
    """`UAVPathPlanner.find_path()` of the backend GA (GAE:166-179) n_paths times — what
    `GeneticAlgorithm.initialize_population` does (GAE:13-18) — in ONE launch.  The dodges take their
    `random.shuffle` orders from a tape drawn natively from a copy of CPython's global generator
    (tf_py_shuffle6_tape); afterwards the generator is advanced by exactly the shuffles the walks consumed, so the
    caller's `random` stream continues as it would after the reference's own loop.
    obstacles: iterable of integer voxels; bounds: ((xlo, xhi), (ylo, yhi), (zlo, zhi)) inclusive.
    -> list of n_paths lists of (x, y, z) int tuples."""
    import random
    _require_cuda()
    obs = np.zeros((len(obstacles), 4), np.int32)
    if len(obstacles):
        obs[:, :3] = np.asarray(list(obstacles), np.int64).reshape(len(obstacles), -1)[:, :3]
    obs_d = torch.from_numpy(obs).cuda()
    i3 = lambda v: (C.c_int32 * 3)(*[int(x) for x in v])
    b6 = (C.c_int32 * 6)(*[int(v) for pair in bounds for v in pair])
    state = random.getstate()
    key0 = np.array(state[1][:624], np.uint32)
    pos0 = int(state[1][624])

    def tape(n):
        key, pos = key0.copy(), C.c_int32(pos0)
        out = np.zeros(max(n, 1), np.int32)
        check(lib.tf_py_shuffle6_tape(key.ctypes.data_as(C.c_void_p), C.byref(pos), n, out.ctypes.data_as(C.c_void_p)),
              "tf_py_shuffle6_tape")
        return out, key, pos.value

    n_sh = 64 * max(n_paths, 1)
    while True:
        orders, _, _ = tape(n_sh)
        sh_d = torch.from_numpy(orders).cuda()
        paths = torch.empty((max(n_paths, 1), cap, 4), dtype=torch.int32, device="cuda")
        lens = torch.zeros(max(n_paths, 1), dtype=torch.int32, device="cuda")
        res = torch.zeros(2, dtype=torch.int32, device="cuda")
        check(lib.tf_find_path_gae(_ptr(obs_d), len(obs), b6, i3(source), i3(destination), _ptr(sh_d), n_sh, n_paths, cap,
                                   _ptr(paths), _ptr(lens), _ptr(res), _stream()), "tf_find_path_gae")
        used, status = (int(v) for v in res.cpu().tolist())
        if status == 0:
            break
        if status == 1:
            if cap >= max_cap:
                raise RuntimeError(f"find_path_gae: a path needs more than {max_cap} waypoints")
            cap *= 2
        else:
            n_sh *= 4
    _, key, pos = tape(used)                                  # advance the caller's generator by what was consumed
    random.setstate((state[0], tuple(int(v) for v in key) + (pos,), state[2]))
    rows, counts = paths.cpu().numpy(), lens.cpu().numpy()
    return [[tuple(int(v) for v in w[:3]) for w in rows[i, :counts[i]]] for i in range(n_paths)]


# ------------------------------------------------------------------------------------ ranking
class TopK:
    """Reusable top-k plan (owns the workspace): order = (fitness desc, index asc) — the stable
    descending sort of HR:114.  worst=True returns the tail of that order, worst first."""

    def __init__(self, P, k, f64=False, device="cuda"):
        self.P, self.k, self.f64 = P, k, f64
        nbytes = lib.tf_topk_workspace_bytes(P, k, int(f64))
        if nbytes < 0:
            raise ValueError(f"top-k: unsupported P={P}, k={k}")
        self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.nbytes = int(nbytes)

    def __call__(self, fit, worst=False, out=None):
        idx = out if out is not None else torch.empty(self.k, dtype=torch.int32, device=fit.device)
        fn = lib.tf_rank_topk_f64 if self.f64 else lib.tf_rank_topk
        _chk_tensor(fit, torch.float64 if self.f64 else torch.float32, "fit", 1)
        check(fn(_ptr(fit), self.P, self.k, int(worst), _ptr(idx), _ptr(self.ws), self.nbytes, _stream()), "tf_rank_topk")
        return idx


class SelectSet:
    """The k best / worst chromosomes as a set in ascending index order (tf_select_set)."""

    def __init__(self, P, k, device="cuda"):
        self.P, self.k = P, k
        nbytes = lib.tf_select_set_workspace_bytes(P)
        if nbytes < 0 or k < 1 or k > P:
            raise ValueError(f"select_set: unsupported P={P}, k={k}")
        self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.nbytes = int(nbytes)

    def __call__(self, fit, worst=False, out=None):
        _chk_tensor(fit, torch.float32, "fit", 1)
        idx = out if out is not None else torch.empty(self.k, dtype=torch.int32, device=fit.device)
        check(lib.tf_select_set(_ptr(fit), self.P, self.k, int(worst), _ptr(idx), _ptr(self.ws), self.nbytes, _stream()),
              "tf_select_set")
        return idx


def rank_topk(fit, k, worst=False):
    return TopK(fit.numel(), k, f64=fit.dtype == torch.float64, device=fit.device)(fit, worst)


# ------------------------------------------------------------------------------------ GA step
def _as_tape(tape, device):
    if tape is None:
        return None
    if isinstance(tape, np.ndarray):
        tape = torch.from_numpy(np.ascontiguousarray(tape, np.int32)).to(device)
    _chk_tensor(tape, torch.int32, "tape", 2)
    if tape.shape[1] != 8:
        raise ValueError("tape must be int32[npairs, 8] (tf_draw records)")
    return tape


def ga_step(pop, fit, top_idx, ga=None, tape=None, seed=0, generation=0, lens=None, out=None, want_parity=True):
    """Selection + crossover + mutation (HR:114-122).  pop f32 or f64 [P,W,4].
    -> (next_pop, next_lens or None, sel i32[npairs,2] or None, cp i32[npairs] or None)"""
    ga = ga or L.default_ga_params()
    f64 = pop.dtype == torch.float64
    _chk_tensor(pop, torch.float64 if f64 else torch.float32, "pop", 3)
    P, W, _ = pop.shape
    ne = min(ga.n_elite, P)
    npairs = (P - ne + 1) // 2
    tape = _as_tape(tape, pop.device)
    if tape is not None and tape.shape[0] < npairs:
        raise ValueError("tape shorter than the number of child pairs")
    nxt = out if out is not None else torch.empty_like(pop)
    next_lens = torch.empty(P, dtype=torch.int32, device=pop.device) if lens is not None else None
    sel = torch.empty((npairs, 2), dtype=torch.int32, device=pop.device) if want_parity else None
    cp = torch.empty(npairs, dtype=torch.int32, device=pop.device) if want_parity else None
    fn = lib.tf_ga_step_f64 if f64 else lib.tf_ga_step
    check(fn(_ptr(pop), _ptr(lens), _ptr(fit), _ptr(top_idx), _ptr(tape), seed, generation, C.byref(ga), P, W,
             _ptr(nxt), _ptr(next_lens), _ptr(sel), _ptr(cp), _stream()), "tf_ga_step")
    return nxt, next_lens, sel, cp


def generation(env, pop, fit, top_idx, ga=None, params=None, tape=None, seed=0, generation=0, out_pop=None,
               out_fit=None, want_collisions=False, want_parity=False, next_top=None, topk=None):
    """Fused GA step + fitness sweep (FP32, fixed W): the child never round-trips through HBM.
    -> (next_pop, next_fit, next_coll or None, sel or None, cp or None)
    next_top (i32[k], with topk = the TopK plan that owns the workspace): also filled with the ranking of the
    children — rank_topk(next_fit, k) — so the generation loop needs no ranking launch (tf_generation_ranked)."""
    ga = ga or L.default_ga_params()
    _chk_tensor(pop, torch.float32, "pop", 3)
    P, W, _ = pop.shape
    params = params or L.default_params(dist_zero_guard=1 if env.semantics != L.SEM_PF else 0)
    ne = min(ga.n_elite, P)
    npairs = (P - ne + 1) // 2
    tape = _as_tape(tape, pop.device)
    nxt = out_pop if out_pop is not None else torch.empty_like(pop)
    nfit = out_fit if out_fit is not None else torch.empty(P, dtype=torch.float32, device=pop.device)
    coll = torch.empty((P, (W + 31) // 32), dtype=torch.int32, device=pop.device) if want_collisions else None
    sel = torch.empty((npairs, 2), dtype=torch.int32, device=pop.device) if want_parity else None
    cp = torch.empty(npairs, dtype=torch.int32, device=pop.device) if want_parity else None
    if next_top is not None:
        if topk is None or topk.f64 or topk.P != P or next_top.numel() != topk.k:
            raise ValueError("generation(next_top=...) needs topk = TopK(P, len(next_top))")
        rc = lib.tf_generation_ranked(env.handle, _ptr(pop), _ptr(fit), _ptr(top_idx), _ptr(tape), seed, generation,
                                      C.byref(ga), C.byref(params), P, W, _ptr(nxt), _ptr(nfit), _ptr(coll), _ptr(sel),
                                      _ptr(cp), _ptr(next_top), topk.k, _ptr(topk.ws), topk.nbytes, _stream())
    else:
        rc = lib.tf_generation(env.handle, _ptr(pop), _ptr(fit), _ptr(top_idx), _ptr(tape), seed, generation,
                               C.byref(ga), C.byref(params), P, W, _ptr(nxt), _ptr(nfit), _ptr(coll), _ptr(sel), _ptr(cp),
                               _stream())
    if rc == L.TF_ERR_UNSUPPORTED:
        # more than 2048 analytic obstacles without a candidate index: the fused kernel stages one table chunk only.
        # The unfused pair gives the same results (tf_ga_step + tf_fitness walk the table in chunks).
        check(lib.tf_ga_step(_ptr(pop), None, _ptr(fit), _ptr(top_idx), _ptr(tape), seed, generation, C.byref(ga), P, W,
                             _ptr(nxt), None, _ptr(sel), _ptr(cp), _stream()), "tf_ga_step")
        check(lib.tf_fitness(env.handle, _ptr(nxt), None, P, W, C.byref(params), _ptr(nfit), _ptr(coll), _stream()),
              "tf_fitness")
        if next_top is not None:
            topk(nfit, out=next_top)
        return nxt, nfit, coll, sel, cp
    check(rc, "tf_generation")
    return nxt, nfit, coll, sel, cp


# ------------------------------------------------------------------------------------ initial populations
def init_diagonal(start, end, P, f64=False, W=None):
    """HR:68-83: the diagonal lattice walk start -> end replicated P times.  -> pop [P, W, 4]"""
    _require_cuda()
    s = (C.c_double * 3)(*[float(v) for v in start])
    e = (C.c_double * 3)(*[float(v) for v in end])
    n = C.c_int(0)
    fn = lib.tf_init_diagonal_f64 if f64 else lib.tf_init_diagonal
    check(fn(s, e, 0, 0, None, C.byref(n), None), "tf_init_diagonal(size)")
    W = n.value if W is None else W
    pop = torch.zeros((P, W, 4), dtype=torch.float64 if f64 else torch.float32, device="cuda")
    check(fn(s, e, P, W, _ptr(pop), C.byref(n), _stream()), "tf_init_diagonal")
    return pop, n.value


def init_walks(seed, P, W, start=(0, 0, 0), lo=(-1e30,) * 3, hi=(1e30,) * 3, start_per_path=None):
    """Seeded +-1 lattice random walks (the synthetic populations of BASELINE configs[1..4])."""
    _require_cuda()
    pop = torch.empty((P, W, 4), dtype=torch.float32, device="cuda")
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    if start_per_path is not None:
        _chk_tensor(start_per_path, torch.float32, "start_per_path", 2)
    check(lib.tf_init_walks(seed, _ptr(start_per_path), f3(start), f3(lo), f3(hi), P, W, _ptr(pop), _stream()),
          "tf_init_walks")
    return pop


# ------------------------------------------------------------------------------------ evolve
class BestTracker:
    """HR:126-128 on the device: best <- pop[slot] whenever fit[slot] > best_fit."""

    def __init__(self, W, f64=False, device="cuda"):
        dt = torch.float64 if f64 else torch.float32
        self.f64 = f64
        self.W = W
        self.best_fit = torch.full((1,), -1.0, dtype=dt, device=device)  # HR:107
        self.best_path = torch.zeros((W, 4), dtype=dt, device=device)
        self.best_gen = torch.full((1,), -1, dtype=torch.int32, device=device)

    def update(self, pop, fit, slot, generation):
        fn = lib.tf_track_best_f64 if self.f64 else lib.tf_track_best
        check(fn(_ptr(pop), _ptr(fit), slot, self.W, generation, _ptr(self.best_fit), _ptr(self.best_path),
                 _ptr(self.best_gen), _stream()), "tf_track_best")


def evolve_batch(pop, env, generations, ga=None, params=None, seed=0, tapes=None, fused=True, start_generation=0,
                 should_stop=None, record_best=True, inplace=False, history=False, resume=None):
    """Runs `generations` GA generations on the device (HR:109-128 generalised):
    rank -> (select, crossover, mutate) -> fitness, with HR's best tracking.

    tapes        optional int32[G, npairs, 8] host draws (np.random order, HR:118-120); else Philox.
    should_stop  callable polled before every generation (the emergency flag of HR:110-112) -> returns None.
    history      True -> per-generation best / mean fitness, accumulated on the device (no host sync).
    resume       a checkpoint made by `checkpoint()` of a previous call's result: the run continues at the
                 saved generation counter.  Philox is counter-based (pair, generation, island), so a resumed
                 run is bit-identical to an uninterrupted one.
    Returns dict(pop, fit, generation, best_path [W,4], best_fit, best_gen[, best_history, mean_history])."""
    ga = ga or L.default_ga_params()
    if resume is not None:
        pop, start_generation, seed = resume["pop"].to("cuda"), int(resume["generation"]), int(resume["seed"])
        inplace = True
    f64 = pop.dtype == torch.float64
    P, W, _ = pop.shape
    params = params or L.default_params(dist_zero_guard=1 if env.semantics != L.SEM_PF else 0)
    k = max(min(ga.n_elite, P), min(ga.n_top, P) if ga.selection == L.SEL_TRUNCATION else 1, 1)
    topk = TopK(P, k, f64=f64, device=pop.device)
    tracker = BestTracker(W, f64=f64, device=pop.device) if record_best else None
    if tracker is not None and resume is not None and "best_fit" in resume:
        tracker.best_fit.copy_(resume["best_fit"])
        tracker.best_path.copy_(resume["best_path"])
        tracker.best_gen.copy_(resume["best_gen"])
    fit = (fitness_batch_f64(pop, None, env, params) if f64 else fitness_batch(pop, None, env, params, want_collisions=False))[0]
    cur, nxt = (pop if inplace else pop.clone()), torch.empty_like(pop)  # ping-pong buffers
    nfit = torch.empty_like(fit)
    top = torch.empty(k, dtype=torch.int32, device=pop.device)
    if tapes is not None and isinstance(tapes, np.ndarray):
        tapes = torch.from_numpy(np.ascontiguousarray(tapes, np.int32)).to(pop.device)
    hist_best = torch.empty(generations, dtype=fit.dtype, device=pop.device) if history else None
    hist_mean = torch.empty(generations, dtype=fit.dtype, device=pop.device) if history else None
    # the fused f32 generation ranks its own children (tf_generation_ranked): one ranking launch before the first
    # generation, none inside the loop
    ranked_gen = fused and not f64
    ntop = torch.empty_like(top) if ranked_gen else None
    for g in range(generations):
        if should_stop is not None and should_stop():
            return None
        if g == 0 or not ranked_gen:
            topk(fit, out=top)
        tape = None if tapes is None else tapes[g]
        gen_no = start_generation + g
        if f64 or not fused:
            ga_step(cur, fit, top, ga, tape=tape, seed=seed, generation=gen_no, out=nxt, want_parity=False)
            if f64:
                fitness_batch_f64(nxt, None, env, params, out=nfit, want_collisions=False)
            else:
                fitness_batch(nxt, None, env, params, out=nfit, want_collisions=False)
        else:
            generation(env, cur, fit, top, ga, params, tape=tape, seed=seed, generation=gen_no, out_pop=nxt, out_fit=nfit,
                       next_top=ntop, topk=topk)
            top, ntop = ntop, top
        cur, nxt = nxt, cur
        fit, nfit = nfit, fit
        if tracker is not None:
            tracker.update(cur, fit, 0, gen_no)
        if history:
            hist_best[g] = fit.max()
            hist_mean[g] = fit.mean()
    out = dict(pop=cur, fit=fit, generation=start_generation + generations, seed=seed)
    if tracker is not None:
        out.update(best_path=tracker.best_path, best_fit=tracker.best_fit, best_gen=tracker.best_gen)
    if history:
        out.update(best_history=hist_best, mean_history=hist_mean)
    return out


def checkpoint(result, path=None):
    """Checkpoint of an `evolve_batch` result (population, generation counter, Philox seed, best-so-far).
    With `path` it is written with torch.save; the dict is returned either way."""
    keep = ("pop", "generation", "seed", "best_path", "best_fit", "best_gen")
    state = {k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in result.items() if k in keep}
    if path is not None:
        torch.save(state, path)
    return state


# ------------------------------------------------------------------------------------ result egress
def encode_waypoints(paths, mode=L.CODEC_QUANT_F32, param=None):
    """The lossy front ends of the reference's waypoint compression (ga_waypoints_compression.py) on the device:
    paths f32 / f64 [n, W, 4] cuda -> CODEC_QUANT_F32: f32[n, W, 3] = np.round(w / q) * q (:2090-2093, q = 0.05);
    CODEC_DELTA_ROUND: f64[n, W, 3] = np.round(np.diff(w, prepend=w[:1]), d) (:875-884, d = 3).  The entropy coder
    (zlib / LZMA) stays on the host: `zlib.compress(encode_waypoints(p)[0].cpu().numpy().tobytes())`."""
    if paths.dim() == 2:
        paths = paths.unsqueeze(0)
    f32 = paths.dtype == torch.float32
    _chk_tensor(paths, torch.float32 if f32 else torch.float64, "paths", 3)
    n, W, four = paths.shape
    if four != 4:
        raise ValueError("paths must be [n, W, 4]")
    if param is None:
        param = 0.05 if mode == L.CODEC_QUANT_F32 else 3
    out = torch.empty((n, W, 3), dtype=torch.float32 if mode == L.CODEC_QUANT_F32 else torch.float64, device=paths.device)
    check(lib.tf_encode_waypoints(_ptr(paths), int(f32), n, W, int(mode), float(param), _ptr(out), _stream()),
          "tf_encode_waypoints")
    return out


# ------------------------------------------------------------------------------------ host round trips
class HostStepper:
    """The array API for callers whose populations live in HOST memory (what a drop-in call from
    the reference's process looks like): each `submit` copies a pinned host population to the GPU,
    runs fitness -> rank -> fused generation there, and copies the fitness vector and the next
    population back.  `depth` requests are in flight on separate streams, so the D2H leg of step n
    overlaps the H2D leg of step n+1 (PCIe is full duplex).  Host formats: "xyz" float32[P, W, 3] (the
    reference's 3-tuples as an array, 12 B per waypoint; tf_unpack_xyz / tf_pack_xyz), "f32" float32[P, W, 4]
    (the device layout, 16 B), "i16" int16[P, W, 4] (lattice chromosomes, 8 B; tf_unpack_i16 / tf_pack_i16)."""

    _HOST = {"f32": (torch.float32, 4), "xyz": (torch.float32, 3), "i16": (torch.int16, 4)}

    def __init__(self, env, P, W, ga=None, params=None, fmt="f32", depth=2, device="cuda"):
        _require_cuda()
        if fmt not in self._HOST:
            raise ValueError("fmt must be 'f32', 'xyz' or 'i16'")
        self.env, self.P, self.W, self.fmt = env, P, W, fmt
        self.ga = ga or L.default_ga_params()
        self.params = params or L.default_params(dist_zero_guard=1 if env.semantics != L.SEM_PF else 0)
        k = max(min(self.ga.n_elite, P), min(self.ga.n_top, P) if self.ga.selection == L.SEL_TRUNCATION else 1, 1)
        hdt, hc = self._HOST[fmt]
        self.slots = []
        for _ in range(depth):
            s = dict(stream=torch.cuda.Stream(), pop=torch.empty((P, W, 4), dtype=torch.float32, device=device),
                     nxt=torch.empty((P, W, 4), dtype=torch.float32, device=device),
                     fit=torch.empty(P, dtype=torch.float32, device=device),
                     nfit=torch.empty(P, dtype=torch.float32, device=device),
                     top=torch.empty(k, dtype=torch.int32, device=device), topk=TopK(P, k, device=device),
                     host_fit=torch.empty(P, dtype=torch.float32).pin_memory(),
                     host_next=torch.empty((P, W, hc), dtype=hdt).pin_memory(), busy=False)
            if fmt != "f32":
                s["stage_in"] = torch.empty((P, W, hc), dtype=hdt, device=device)
                s["stage_out"] = torch.empty((P, W, hc), dtype=hdt, device=device)
            self.slots.append(s)
        self._n = 0

    def bytes_per_step(self):
        hdt, hc = self._HOST[self.fmt]
        row = self.P * self.W * hc * (4 if hdt == torch.float32 else 2)
        return row, row + self.P * 4

    def submit(self, host_pop, seed=0, generation=0):
        """host_pop: pinned CPU tensor in the stepper's host format.  Returns a ticket."""
        s = self.slots[self._n % len(self.slots)]
        if s["busy"]:
            raise RuntimeError("HostStepper: collect the oldest result before submitting more work")
        n4 = self.P * self.W
        # the environment (voxel grid, candidate indices, hazard map) and the caller's earlier work may still be in
        # flight on the caller's stream; the stepper's own streams do not synchronise with it implicitly
        s["stream"].wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s["stream"]):
            if self.fmt == "f32":
                s["pop"].copy_(host_pop, non_blocking=True)
            else:
                s["stage_in"].copy_(host_pop, non_blocking=True)
                unpack = lib.tf_unpack_i16 if self.fmt == "i16" else lib.tf_unpack_xyz
                check(unpack(_ptr(s["stage_in"]), _ptr(s["pop"]), n4, _stream()), "tf_unpack")
            fitness_batch(s["pop"], None, self.env, self.params, out=s["fit"], want_collisions=False)
            s["topk"](s["fit"], out=s["top"])
            check(lib.tf_generation(self.env.handle, _ptr(s["pop"]), _ptr(s["fit"]), _ptr(s["top"]), None, seed,
                                    generation, C.byref(self.ga), C.byref(self.params), self.P, self.W, _ptr(s["nxt"]),
                                    _ptr(s["nfit"]), None, None, None, _stream()), "tf_generation")
            s["host_fit"].copy_(s["fit"], non_blocking=True)
            if self.fmt == "f32":
                s["host_next"].copy_(s["nxt"], non_blocking=True)
            else:
                pack = lib.tf_pack_i16 if self.fmt == "i16" else lib.tf_pack_xyz
                check(pack(_ptr(s["nxt"]), _ptr(s["stage_out"]), n4, _stream()), "tf_pack")
                s["host_next"].copy_(s["stage_out"], non_blocking=True)
        s["busy"] = True
        self._n += 1
        return self._n - 1

    def result(self, ticket):
        """Blocks until that request's D2H copies have landed.  -> (fitness f32[P], next population) pinned
        host tensors owned by the stepper (valid until the slot is reused `depth` submits later)."""
        s = self.slots[ticket % len(self.slots)]
        s["stream"].synchronize()
        s["busy"] = False
        return s["host_fit"], s["host_next"]
