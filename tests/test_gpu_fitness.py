"""GPU parity: tf_fitness / tf_fitness_f64 / tf_point_query_f64 through the C-ABI (ctypes) vs
  (1) goldens recorded from the unmodified reference, and (2) the CPU oracle on seeded inputs.
Bars: collision flags bit-exact; FP32 fitness bit-exact vs the fixed-order oracle and within
1e-5 relative of the reference (north_star); FP64 fitness bit-identical to the reference."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as O
from tests.helpers import (as_u32, golden_case, load_golden, oracle_env_from_table, random_table, rel_err,
                           unpack_bits, valid_mask)

pytestmark = pytest.mark.gpu
FIT_RTOL = 1e-5


@pytest.fixture(scope="module")
def tf():
    import torch
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    assert torch.cuda.is_available()
    return torch, L, B


def _cases(fam):
    return load_golden(f"fitness_{fam}.npz")[1]["cases"]


@pytest.mark.parametrize("fam,case", [("hr", c) for c in _cases("hr")] + [("pf", c) for c in _cases("pf")])
def test_golden_cases(tf, fam, case):
    torch, L, B = tf
    g, _ = load_golden(f"fitness_{fam}.npz")
    d = golden_case(g, fam, case)
    sem = L.SEM_HR if fam == "hr" else L.SEM_PF
    env = B.Environment(sem, d["obs64"], 2.0)
    P, W, _ = d["pop64"].shape
    valid = valid_mask(d["lens"], W)
    lens = torch.from_numpy(d["lens"]).cuda()
    prm = L.default_params(dist_zero_guard=1 if fam == "hr" else 0)

    # float64 reference-order kernel: bit-identical to the reference
    pop64 = torch.from_numpy(d["pop64"]).cuda()
    f64, c64 = B.fitness_batch_f64(pop64, lens, env, prm)
    assert np.array_equal(f64.cpu().numpy(), d["fitness"])
    assert np.array_equal(unpack_bits(as_u32(c64), W) & valid, d["collide"] & valid)

    # FP32 throughput kernel: flags bit-exact, fitness bit-exact vs oracle and 1e-5 vs reference
    pop32 = pop64.float()
    f32, c32 = B.fitness_batch(pop32, lens, env, prm)
    assert np.array_equal(unpack_bits(as_u32(c32), W) & valid, d["collide"] & valid)
    assert rel_err(f32.cpu().numpy(), d["fitness"]) <= FIT_RTOL
    oenv = O.OracleEnv(O.SEM_HR if fam == "hr" else O.SEM_PF, d["obs32"], 2.0)
    of32, oc32 = O.fitness_f32(oenv, d["pop64"].astype(np.float32), d["lens"], O.default_params(guard=prm.dist_zero_guard))
    assert np.array_equal(of32, f32.cpu().numpy())
    assert np.array_equal(oc32, as_u32(c32))

    # point queries (is_collision / distance_to_nearest_obstacle, is_obstacle / distance_to_obstacles)
    pts = d["pop64"][:, :, :3].reshape(-1, 3)
    coll, dn = B.point_query_f64(pts, env)
    vm = valid.reshape(-1)
    assert np.array_equal(coll[vm], d["collide"].reshape(-1)[vm])
    assert np.array_equal(dn[vm], d["dnear"].reshape(-1)[vm])


@pytest.mark.parametrize("sem", ["hr", "pf"])
@pytest.mark.parametrize("P,W,N", [(1, 3, 1), (777, 20, 10), (4096, 32, 1024), (1500, 64, 257), (300, 100, 64),
                                   (200, 256, 33), (512, 32, 5000), (64, 1, 5), (100, 2, 0)])
def test_random_vs_oracle(tf, sem, P, W, N):
    torch, L, B = tf
    rng = np.random.default_rng(P * 1000 + W + N)
    world = 64
    table = random_table(rng, N, world, sem_pf=(sem == "pf"), aabb=True)
    semc = L.SEM_PF if sem == "pf" else L.SEM_HR
    env = B.Environment(semc, table, 2.0)
    pop_h = O.init_walks(P + W, P, W, (32, 32, 32), (0, 0, 0), (63, 63, 63))
    pop = torch.from_numpy(pop_h).cuda()
    fit, coll = B.fitness_batch(pop, None, env)
    oenv = oracle_env_from_table(O.SEM_PF if sem == "pf" else O.SEM_HR, table, 2.0)
    O.set_threads(8)
    ofit, ocoll = O.fitness_f32(oenv, pop_h, None, O.default_params(guard=0 if sem == "pf" else 1))
    O.set_threads(1)
    assert np.array_equal(ocoll, as_u32(coll))
    assert np.array_equal(ofit, fit.cpu().numpy())


def test_float_inputs_and_ragged_lens(tf):
    torch, L, B = tf
    rng = np.random.default_rng(5)
    P, W, N = 2048, 48, 100
    table = random_table(rng, N, 32, sem_pf=True)
    table[:, :3] += rng.uniform(-0.5, 0.5, size=(N, 3))
    table[:, 3] = rng.uniform(0.5, 3.5, size=N)
    pop_h = (rng.uniform(0, 32, size=(P, W, 4))).astype(np.float32)
    lens_h = rng.integers(0, W + 1, size=P).astype(np.int32)
    for sem, osem in ((L.SEM_PF, O.SEM_PF), (L.SEM_HR, O.SEM_HR)):
        env = B.Environment(sem, table, 1.5)
        fit, coll = B.fitness_batch(torch.from_numpy(pop_h).cuda(), torch.from_numpy(lens_h).cuda(), env)
        oenv = oracle_env_from_table(osem, table, 1.5)
        ofit, ocoll = O.fitness_f32(oenv, pop_h, lens_h, O.default_params(guard=0 if sem == L.SEM_PF else 1))
        assert np.array_equal(ocoll, as_u32(coll))
        assert np.array_equal(ofit, fit.cpu().numpy(), equal_nan=True)


def _hazard_field(rng, nt, ny, nx):
    return rng.uniform(0, 1, size=(nt, ny, nx)).astype(np.float32)


def test_voxel_hazard_dynamic_samples(tf):
    """Extensions (no reference implementation): CUDA == CPU restatement."""
    torch, L, B = tf
    rng = np.random.default_rng(11)
    P, W, N = 3000, 64, 300
    dims = (64, 64, 32)
    table = random_table(rng, N, 64, sem_pf=True, max_size=3)
    table[:, 2] = rng.integers(0, 32, size=N)
    pop_h = O.init_walks(3, P, W, (32, 32, 16), (0, 0, 0), (63, 63, 31))
    pop = torch.from_numpy(pop_h).cuda()
    oenv_an = oracle_env_from_table(O.SEM_PF, table, 0.0)
    bricks_h = O.voxelize(oenv_an, *dims)
    bricks = B.voxelize(table, dims, L.SEM_PF)
    assert np.array_equal(bricks_h, as_u32(bricks)), "voxeliser differs"
    # (1) voxel-only environment: collision flags == analytic flags on lattice paths
    env_v = B.Environment.empty().set_voxels(bricks, dims)
    fit_v, coll_v = B.fitness_batch(pop, None, env_v)
    env_a = B.Environment(L.SEM_PF, table, 0.0)
    fit_a, coll_a = B.fitness_batch(pop, None, env_a)
    assert np.array_equal(as_u32(coll_v), as_u32(coll_a)), "voxel flags must equal analytic flags"
    ov = O.OracleEnv(O.SEM_NONE, None, bricks=bricks_h, dims=dims)
    ofit_v, ocoll_v = O.fitness_f32(ov, pop_h, None, O.default_params(guard=1))
    assert np.array_equal(ocoll_v, as_u32(coll_v)) and np.array_equal(ofit_v, fit_v.cpu().numpy())
    # (2) everything at once: voxels + moving spheres + hazard map + 3 samples per segment
    dyn = np.zeros((40, 8), np.float32)
    dyn[:, :3] = rng.uniform(0, 64, size=(40, 3))
    dyn[:, 2] = rng.uniform(0, 32, size=40)
    dyn[:, 3] = rng.integers(1, 4, size=40)
    dyn[:, 4:7] = rng.uniform(-1, 1, size=(40, 3))
    haz_h = _hazard_field(rng, 24, 32, 32)
    hmap = (0.5, 0.5, 0.0, 0.0, 23.0 / (W - 1))
    env = B.Environment(L.SEM_HR, table, 2.0).set_voxels(bricks, dims).set_dynamic(dyn, 1.0)
    env.set_hazard(torch.from_numpy(haz_h).cuda(), *hmap)
    for S in (1, 3):
        prm = L.default_params(hazard_weight=0.25, samples_per_segment=S)
        fit, coll = B.fitness_batch(pop, None, env, prm)
        oe = oracle_env_from_table(O.SEM_HR, table, 2.0, dyn=dyn, dyn_safety=1.0, bricks=bricks_h, dims=dims,
                                   hazard=haz_h, hazard_map=hmap)
        O.set_threads(8)
        ofit, ocoll = O.fitness_f32(oe, pop_h, None, O.default_params(guard=1, hazard_weight=0.25, samples=S))
        O.set_threads(1)
        assert np.array_equal(ocoll, as_u32(coll))
        assert np.array_equal(ofit, fit.cpu().numpy())


def test_voxel_pack_linear(tf):
    torch, L, B = tf
    rng = np.random.default_rng(2)
    dims = (32, 16, 24)
    nbits = dims[0] * dims[1] * dims[2]
    lin = rng.integers(0, 2 ** 32, size=nbits // 32, dtype=np.uint64).astype(np.uint32)
    got = B.voxel_pack_linear(torch.from_numpy(lin.view(np.int32)).cuda(), dims)
    assert np.array_equal(O.voxel_pack_linear(lin, *dims), as_u32(got))


def test_error_behaviour(tf):
    torch, L, B = tf
    env = B.Environment.empty()
    with pytest.raises(TypeError):
        B.fitness_batch(torch.zeros((4, 8, 4)), None, env)          # CPU tensor
    with pytest.raises(ValueError):
        B.fitness_batch(torch.zeros((4, 300, 4), device="cuda"), None, env)  # W > TF_W_MAX
    with pytest.raises(ValueError):
        B.Environment(7, None, 0.0)
    fit, coll = B.fitness_batch(torch.zeros((0, 8, 4), device="cuda"), None, env)
    assert fit.numel() == 0
    # the small misc entry points of the C-ABI
    import ctypes as C
    sms, major, minor, hbm = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    assert L.lib.tf_device_info(C.byref(sms), C.byref(major), C.byref(minor), C.byref(hbm)) == 0
    assert major.value == 10 and sms.value >= 100 and hbm.value > (100 << 30)
    assert L.lib.tf_last_cuda_error() == 0
    n0 = L.launch_count()
    B.fitness_batch(torch.zeros((4, 8, 4), device="cuda"), None, env)
    assert L.launch_count() == n0 + 1
    assert L.lib.tf_set_option(99, 0) < 0 and L.lib.tf_set_option(L.OPT_STREAM_SCHEDULE, 3) < 0


def test_dynamic_index_equals_brute_force(tf):
    """tf_env_build_dynamic_index: bit-identical fitness / flags, with points inside and outside the
    index box, integer and fractional (sampled) times, and a cell capacity small enough to overflow."""
    torch, L, B = tf
    rng = np.random.default_rng(21)
    P, W, n = 3000, 64, 700
    world = 256
    dyn = np.zeros((n, 8), np.float32)
    dyn[:, :3] = rng.uniform(0, world, size=(n, 3))
    dyn[:, 3] = rng.integers(1, 5, size=n)
    dyn[:, 4:7] = rng.uniform(-1, 1, size=(n, 3))
    pop_h = O.init_walks(13, P, W, (128, 128, 128), (-20, -20, -20), (world + 20, world + 20, world + 20),
                         start_pp=np.concatenate([rng.uniform(-10, world + 10, size=(P, 3)).astype(np.float32).round(),
                                                  np.zeros((P, 1), np.float32)], axis=1))
    pop = torch.from_numpy(pop_h).cuda()
    oe = O.OracleEnv(O.SEM_NONE, None, 0.0, dyn=dyn, dyn_safety=2.0)
    for S in (1, 2):
        prm = L.default_params(samples_per_segment=S)
        ofit, ocoll = None, None
        for cap, band in ((0, 1), (64, 1), (3, 1), (128, 16), (256, 64)):
            env = B.Environment.empty().set_dynamic(dyn, 2.0)
            if cap:
                env.build_dynamic_index(W, (0, 0, 0), (world, world, world), cell=32.0, cap=cap, band=band)
            fit, coll = B.fitness_batch(pop, None, env, prm)
            if ofit is None:
                O.set_threads(8)
                ofit, ocoll = O.fitness_f32(oe, pop_h, None, O.default_params(guard=1, samples=S))
                O.set_threads(1)
            assert np.array_equal(ocoll, as_u32(coll)), (S, cap, band)
            assert np.array_equal(ofit, fit.cpu().numpy()), (S, cap, band)


def test_indexed_hr_spheres_equal_plain_hr(tf):
    """Environment.from_hr_obstacles_indexed (candidate index, zero-velocity spheres) == from_hr_obstacles."""
    torch, L, B = tf
    rng = np.random.default_rng(8)
    world, n, P, W = 256, 1500, 4000, 32
    obstacles = [{"position": [int(v) for v in rng.integers(0, world, size=3)], "size": int(rng.integers(1, 5))} for _ in range(n)]
    for o in obstacles[::7]:
        del o["size"]                       # HR:20 default
    pop = B.init_walks(3, P, W, lo=(-5, -5, -5), hi=(world + 5,) * 3,
                       start_per_path=torch.from_numpy(np.concatenate([rng.integers(-3, world + 3, size=(P, 3)), np.zeros((P, 1))], axis=1).astype(np.float32)).cuda())
    plain = B.Environment.from_hr_obstacles(obstacles, 2.0)
    fast = B.Environment.from_hr_obstacles_indexed(obstacles, (0, 0, 0), (world,) * 3, cell=16.0, safety_radius=2.0)
    assert fast.dynamic_index_stats()["overflowed"] == 0
    f0, c0 = B.fitness_batch(pop, None, plain)
    f1, c1 = B.fitness_batch(pop, None, fast)
    assert torch.equal(f0, f1) and torch.equal(c0, c1)


@pytest.mark.parametrize("sem", ["hr", "pf"])
@pytest.mark.parametrize("W", [20, 32, 64])
def test_static_index_equals_full_table_loop(tf, sem, W):
    """tf_env_build_static_index: bit-identical fitness / flags to the brute-force table loop, for points
    inside and outside the box, generous and overflowing list capacities, float obstacle data."""
    torch, L, B = tf
    rng = np.random.default_rng(W + (7 if sem == "pf" else 0))
    world, n, P = 128, 900, 3000
    table = random_table(rng, n, world, sem_pf=(sem == "pf"), aabb=True)
    table[: n // 3, :3] += rng.uniform(-0.5, 0.5, size=(n // 3, 3))      # non-lattice centres
    table[: n // 5, 3] = rng.uniform(0.5, 4.0, size=n // 5)              # non-integer sizes
    semc = L.SEM_PF if sem == "pf" else L.SEM_HR
    starts = np.concatenate([rng.integers(-6, world + 6, size=(P, 3)), np.zeros((P, 1))], axis=1).astype(np.float32)
    pop = B.init_walks(5, P, W, lo=(-8, -8, -8), hi=(world + 8,) * 3, start_per_path=torch.from_numpy(starts).cuda())
    pop[: P // 4, :, :3] += torch.rand((P // 4, W, 3), device="cuda") - 0.5   # non-lattice waypoints too
    plain = B.Environment(semc, table, 2.0)
    f0, c0 = B.fitness_batch(pop, None, plain)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, mutation_rate=0.5)
    top = B.rank_topk(f0, 2)
    g0 = B.generation(plain, pop, f0, top, ga, seed=1, generation=0, want_collisions=True)
    for cell, cap, zc, zcap in ((8.0, 96, 1.0, 32), (16.0, 128, 2.0, 64), (32.0, 6, 4.0, 2)):
        env = B.Environment(semc, table, 2.0).build_static_index((0, 0, 0), (world,) * 3, cell, cap, zc, zcap)
        f1, c1 = B.fitness_batch(pop, None, env)
        assert torch.equal(f0, f1) and torch.equal(c0, c1), (cell, cap)
        g1 = B.generation(env, pop, f0, top, ga, seed=1, generation=0, want_collisions=True)
        assert torch.equal(g0[0], g1[0]) and torch.equal(g0[1], g1[1]) and torch.equal(g0[2], g1[2])
    st = B.Environment(semc, table, 2.0).build_static_index((0, 0, 0), (world,) * 3, 8.0, 96).static_index_stats()
    assert st["overflowed"] == 0 and st["mean_len"] < 40


def test_voxelize_points(tf):
    """LiDAR points -> occupancy bits: downsample_lidar's own arithmetic (lidar_compression.py:242-246, float64
    division, truncation toward zero).  CUDA == oracle on random clouds with an origin, and CUDA == the voxels the
    UNMODIFIED reference function returned (tests/golden/downsample_lidar.npz)."""
    torch, L, B = tf
    rng = np.random.default_rng(4)
    dims, origin, vs = (64, 32, 16), (-3.0, 2.0, 0.5), 0.25
    pts = rng.uniform(-5, 15, size=(20000, 4)).astype(np.float32)
    got = as_u32(B.voxelize_points(torch.from_numpy(pts).cuda(), dims, origin, vs))
    assert np.array_equal(got, O.voxelize_points(pts, dims, origin, vs)) and int(got.any())
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "downsample_lidar.npz"))
    dims = (64, 64, 64)
    for case in json.loads(str(g["meta"]))["cases"]:
        p4 = np.zeros((len(g[f"pts_{case}"]), 4), np.float32)
        p4[:, :3] = g[f"pts_{case}"]
        vs = float(g[f"vs_{case}"])
        bricks = as_u32(B.voxelize_points(torch.from_numpy(p4).cuda(), dims, (0.0, 0.0, 0.0), vs))
        assert np.array_equal(bricks, O.voxelize_points(p4, dims, (0.0, 0.0, 0.0), vs))
        grid = np.rint(g[f"vox_{case}"] / vs).astype(np.int64)          # the reference's unique voxels
        inside = grid[np.all((grid >= 0) & (grid < np.array(dims)), axis=1)]
        assert all(O.voxel_test(bricks, dims, *r) for r in inside)
        assert int(sum(bin(int(w)).count("1") for w in bricks)) == len(inside)


def test_library_sqrt_is_correctly_rounded_for_every_float(tf):
    """The FP32 model needs IEEE sqrt (HR:51-52, PF:162).  The library's own sequence (rsqrt seed + fma
    Newton step, zero folded in) must equal sqrt.rn.f32 for all 2^32 bit patterns."""
    torch, L, B = tf
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for first in (0, 1 << 31):
        L.check(L.lib.tf_selftest_sqrt(first, 1 << 31, bad.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "tf_selftest_sqrt")
    assert int(bad.item()) == 0


def test_library_reciprocal_is_correctly_rounded_for_every_float(tf):
    """HR:46's 1 / (1 + d): the library's reciprocal (MUFU.RCP seed + fma Newton step for 1 <= y < 2^126, div.rn for
    every other input) must equal div.rn.f32(1, y) for all 2^32 bit patterns, scalar and batched form."""
    torch, L, B = tf
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for first in (0, 1 << 31):
        L.check(L.lib.tf_selftest_rcp(first, 1 << 31, bad.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "tf_selftest_rcp")
    assert int(bad.item()) == 0
