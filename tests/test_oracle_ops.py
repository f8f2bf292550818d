"""Pins the oracle's restatements of three operators whose reference code could not be pinned through whole runs:

  * PF GeneticAlgorithm.crossover / .mutate (PF:196-208): PF's own evolve() drops its children (PF:216-218), so the
    operators are recorded by calling the unmodified methods directly (tests/golden/ops_pf.npz);
  * downsample_lidar (lidar_compression.py:242-246), the voxel assignment behind tf_voxelize_points
    (tests/golden/downsample_lidar.npz, the function's own source lines exec'd unmodified);
  * the roulette selection of the backend GA (GAE:28-34), i.e. CPython's random.choices, which is available here:
    the oracle is checked against random.choices itself.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import c_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_ops_pf():
    g = np.load(os.path.join(GOLD, "ops_pf.npz"))
    n, _, Wmax, _ = g["parents"].shape
    pop = np.zeros((2 * n, Wmax, 4), np.float64)
    pop[:, :, :3] = g["parents"].reshape(2 * n, Wmax, 3)
    lens = g["lens"].reshape(2 * n).astype(np.int32)
    tape = g["tape"].copy()
    tape[:, 0] = 2 * np.arange(n)          # the parents of pair j are rows 2j and 2j + 1 of the population
    tape[:, 1] = 2 * np.arange(n) + 1
    return g, pop, lens, tape


def check_children(g, nxt, nlens):
    n = g["parents"].shape[0]
    want, wlen = g["children"].astype(np.float64), g["clens"]
    assert np.array_equal(nlens.reshape(n, 2), wlen)
    for i in range(n):
        for c in range(2):
            L = wlen[i, c]
            assert np.array_equal(nxt[2 * i + c, :L, :3], want[i, c, :L]), (i, c)


def test_pf_crossover_and_mutate_vs_reference_tape():
    g, pop, lens, tape = load_ops_pf()
    assert bool(g["valid"].all())
    # some recorded crossover points lie beyond the second parent (PF:197 only looks at parent1): general slicing
    assert int((g["tape"][:, 2] > g["lens"][:, 1]).sum()) > 10
    # endpoints are mutated too (PF:204 randint(0, len))
    mp = np.concatenate([g["tape"][:, 3], g["tape"][:, 5]])
    assert (mp == 0).any()
    ga = O.default_ga(n_elite=0, selection=1, mutation=1)        # tape: pa / pb are population indices
    for dtype in (np.float64, np.float32):
        nxt, nlens, sel, cp = O.ga_step(pop.astype(dtype), np.zeros(len(pop), dtype), np.zeros(2, np.int32), ga, tape=tape, lens=lens)
        check_children(g, nxt.astype(np.float64), nlens)
        assert np.array_equal(cp, g["tape"][:, 2])


def test_top2_selection_is_pf_select_parents():
    """TF_SEL_TOP2 (Philox mode): every pair's parents are the two best of the stable ranking (PF:192-194)."""
    rng = np.random.default_rng(1)
    P, W = 64, 12
    pop = np.zeros((P, W, 4), np.float32)
    pop[:, :, :3] = rng.integers(0, 9, size=(P, W, 3))
    fit = rng.integers(0, 5, size=P).astype(np.float32)           # ties: the stable order decides
    top = O.topk(fit, 2)
    order = sorted(range(P), key=lambda i: -fit[i])                # Python's stable sort, reverse via negation
    assert list(top) == order[:2]
    ga = O.default_ga(n_elite=2, selection=2, mutation=1, mutation_rate=0.0)
    nxt, _, sel, cp = O.ga_step(pop, fit, top, ga, seed=5, gen=1)
    assert (sel[:, 0] == top[0]).all() and (sel[:, 1] == top[1]).all()
    for j in range(len(cp)):
        a, b = pop[top[0]], pop[top[1]]
        assert np.array_equal(nxt[2 + 2 * j], np.concatenate([a[:cp[j]], b[cp[j]:]]))


@pytest.mark.parametrize("case", ["unit", "half", "third", "negative", "coarse"])
def test_voxelize_points_is_downsample_lidar(case):
    g = np.load(os.path.join(GOLD, "downsample_lidar.npz"))
    pts, vox, vs = g[f"pts_{case}"], g[f"vox_{case}"], float(g[f"vs_{case}"])
    dims = (64, 64, 64)
    p4 = np.zeros((len(pts), 4), np.float32)
    p4[:, :3] = pts
    bricks = O.voxelize_points(p4, dims, (0.0, 0.0, 0.0), vs)
    got = {(x, y, z) for x in range(dims[0]) for y in range(dims[1]) for z in range(dims[2])
           if O.voxel_test(bricks, dims, x, y, z)} if case == "unit" else None
    # reference voxels -> integer grid coordinates (the function returns unique_voxels * voxel_size)
    grid = np.unique((pts.astype(np.float64) / vs).astype(int), axis=0)
    assert np.array_equal(grid * vs, vox), "golden self-consistency: np.unique(grid) * voxel_size"
    inside = grid[np.all((grid >= 0) & (grid < np.array(dims)), axis=1)]
    want = {tuple(int(v) for v in r) for r in inside}
    if got is None:
        got = {tuple(int(v) for v in r) for r in inside if O.voxel_test(bricks, dims, *r)}
        n_bits = int(sum(bin(int(w)).count("1") for w in bricks))
        assert n_bits == len(want), "no voxel outside the reference's set may be occupied"
    assert got == want and len(want) > 1000


@pytest.mark.parametrize("P,seed", [(2, 0), (20, 1), (50, 2), (1000, 3)])
def test_roulette_is_random_choices(P, seed):
    rng = np.random.default_rng(seed)
    score = -rng.integers(20, 90, size=P).astype(np.int32)        # GAE:20-26: always negative
    total = int(score.sum())
    weights = [int(s) / total for s in score]                     # GAE:30-32
    cum = O.roulette_cumweights(score)
    import itertools
    assert np.array_equal(cum, np.array(list(itertools.accumulate(weights))))
    random.seed(seed)
    st = random.getstate()
    picks = random.choices(range(P), weights, k=400)              # GAE:33
    random.setstate(st)
    u = np.array([random.random() for _ in range(400)])
    assert list(O.select_roulette(cum, u)) == picks
    # the quirk: the WORST score carries the largest weight
    assert int(np.argmax(np.diff(np.concatenate([[0.0], cum])))) == int(np.argmin(score))


@pytest.mark.parametrize("case", ["route", "lattice", "ties", "short"])
def test_waypoint_codecs_vs_reference(case):
    """orc_encode_waypoints == compress_waypoints' quantiser (ga_waypoints_compression.py:2090-2093; the zlib stream
    of our bytes is the reference's stream) and == delta_encode + quantize (:875-884), bit for bit."""
    import zlib
    g = np.load(os.path.join(GOLD, "waypoint_codecs.npz"))
    w = g[f"in_{case}"]
    p4 = np.zeros((1, len(w), 4))
    p4[0, :, :3] = w
    q = O.encode_waypoints(p4, 0, 0.05)[0]
    assert np.array_equal(q.view(np.uint32), g[f"q32_{case}"].view(np.uint32))
    if json.loads(str(g["meta"]))["zlib"] == zlib.ZLIB_VERSION:
        assert np.array_equal(np.frombuffer(zlib.compress(q.tobytes()), np.uint8), g[f"zlib_{case}"])
    d = O.encode_waypoints(p4, 1, 3)[0]
    assert np.array_equal(d.view(np.uint64), g[f"dq_{case}"].view(np.uint64))


def _gae_scenarios():
    g = np.load(os.path.join(GOLD, "find_path_gae.npz"))
    return g, json.loads(str(g["meta"]))["scenarios"]


@pytest.mark.parametrize("seed", _gae_scenarios()[1])
def test_find_path_gae_vs_reference(seed):
    """orc_find_path_gae == UAVPathPlanner.find_path of the unmodified GA_enhanced.py (GAE:166-179) on its own voxel
    world, with the recorded random.shuffle orders: every path, waypoint for waypoint, and the shuffle count."""
    g, _ = _gae_scenarios()
    wp, lens, sh = g[f"wp_{seed}"], g[f"lens_{seed}"], g[f"shuffles_{seed}"]
    bounds = g[f"bounds_{seed}"].reshape(6)
    P, cap = len(lens), int(lens.max()) + 5
    paths, olens, used, status = O.find_path_gae(g[f"obs_{seed}"], bounds, g[f"src_{seed}"], g[f"dst_{seed}"], sh, P, cap)
    assert status == 0 and used == len(sh)
    assert np.array_equal(olens, lens)
    for i in range(P):
        assert np.array_equal(paths[i, :lens[i], :3], wp[i, :lens[i]]), i
    if len(sh):      # too short a tape / too small a capacity are reported, not silently truncated
        assert O.find_path_gae(g[f"obs_{seed}"], bounds, g[f"src_{seed}"], g[f"dst_{seed}"], sh[:-1], P, cap)[3] == 2
    if lens.max() > 1:
        assert O.find_path_gae(g[f"obs_{seed}"], bounds, g[f"src_{seed}"], g[f"dst_{seed}"], sh, P, int(lens.max()) - 1)[3] == 1


@pytest.mark.parametrize("seed", _gae_scenarios()[1])
def test_native_shuffle_tape_is_random_shuffle(seed):
    """tf_py_shuffle6_tape (host function of the C-ABI) == the orders random.shuffle produced inside the reference's
    dodge_obstacle (GAE:157), and it leaves CPython's generator where the reference left it."""
    import ctypes as C
    import __graft_entry__ as ge
    ge.build()
    from tracefly_b200 import _lib as L
    g, _ = _gae_scenarios()
    sh = g[f"shuffles_{seed}"]
    random.seed(int(seed))
    st = random.getstate()
    key = np.array(st[1][:624], np.uint32)
    pos = C.c_int32(int(st[1][624]))
    out = np.zeros(max(len(sh), 1), np.int32)
    assert L.lib.tf_py_shuffle6_tape(key.ctypes.data_as(C.c_void_p), C.byref(pos), len(sh), out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out[:len(sh)], sh)
    random.setstate((st[0], tuple(int(v) for v in key) + (pos.value,), st[2]))
    assert random.random() == float(g[f"after_{seed}"])
