"""Drop-in for "GA_complex _PF.py": under the reference's seeds the default scenario (PF:255-269)
must reproduce the obstacle set, the (never-changing) ragged population, every float64 fitness and
best_path() of the unmodified reference (tests/golden/evolve_pf_default.npz)."""
import random

import numpy as np
import pytest

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
KINDS = ["sphere", "cube", "cylinder"]


def test_pf_default_scenario_matches_reference():
    import torch
    assert torch.cuda.is_available()
    from tracefly_b200 import ga_complex_pf as pf
    g, meta = load_golden("evolve_pf_default.npz")
    random.seed(0)
    np.random.seed(0)
    env = pf.ComplexEnvironment((0, 20), (0, 20), (0, 10), 10)                      # PF:256-261
    assert [KINDS.index(o["shape"]) for o in env.obstacles] == g["obs_kind"].tolist()
    assert np.array_equal(np.array([o["position"] for o in env.obstacles], float), g["obs_pos"])
    ga = pf.GeneticAlgorithm(population_size=50, mutation_rate=0.1, generations=100, environment=env,
                             source=(0, 0, 0), destination=(15, 15, 8))           # PF:265
    ga.evolve()                                                                    # PF:267
    assert [len(p.waypoints) for p in ga.population] == g["lens"].tolist()
    for p, wp, n, f in zip(ga.population, g["wp"], g["lens"], g["fitness"]):
        assert np.array_equal(np.array(p.waypoints, float), wp[:n])
        assert p.fitness == f                                                      # float64, bit-identical
    assert np.array_equal(np.array(ga.best_path().waypoints, float), g["best"])   # PF:269
    # the point queries behind PathPlanner
    assert env.is_obstacle(env.obstacles[0]["position"]) is True
    assert env.distance_to_obstacles(env.obstacles[0]["position"]) == 0.0
    p = pf.Path([(0, 0, 0), (1, 1, 0), (1, 1, 1)])
    assert p.is_turning(1) and not p.is_turning(2)


def test_find_path_kernel_matches_reference_walks():
    """tf_find_path_pf (PF:93-132) through the C-ABI: the 180 walks recorded from the unmodified PathPlanner,
    waypoint for waypoint, plus the drop-in's list-of-tuples form and the capacity-retry path."""
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    from tracefly_b200 import ga_complex_pf as pf
    g, meta = load_golden("find_path_pf.npz")
    for scen in meta["scenarios"]:
        obstacles = [{"shape": KINDS[k], "position": tuple(int(v) for v in p), "size": int(s)}
                     for k, p, s in zip(g[f"kind_{scen}"], g[f"pos_{scen}"], g[f"size_{scen}"])]
        env = B.Environment.from_pf_obstacles(obstacles)
        rows = np.nonzero(g["scen"] == scen)[0]
        src, dst = g[f"src_{scen}"], g[f"dst_{scen}"]
        for dodge in L.DODGE:
            sel = [r for r in rows if str(g["dodge"][r]) == dodge]
            paths, lens = B.find_path_pf(env, src, dst, g["hv"][sel], dodge, cap=4)     # cap 4 forces the retry loop
            lens = lens.cpu().numpy()
            assert np.array_equal(lens, g["lens"][sel])
            paths = paths.cpu().numpy()
            for i, r in enumerate(sel):
                assert np.array_equal(paths[i, :lens[i], :3], g["wp"][r, :lens[i]])
        # the drop-in surface
        world = pf.ComplexEnvironment((0, 1), (0, 1), (0, 1), 0)
        world.obstacles = obstacles
        whole = all(float(v).is_integer() for v in list(src) + list(dst))
        as_t = (lambda v: tuple(int(x) for x in v)) if whole else (lambda v: tuple(float(x) for x in v))
        planner = pf.PathPlanner(as_t(src), as_t(dst), world)
        r = rows[7]
        got = planner.find_path(str(g["dodge"][r]), height_variation=int(g["hv"][r]))
        assert got == [as_t(q) for q in g["wp"][r, :g["lens"][r]]]
        assert got[0] is planner.source


def test_find_path_kernel_equals_oracle_on_dense_worlds():
    """20,000 walks through random dense shaped-obstacle fields (dodges, stuck walks, loop guard): device == CPU
    restatement, every waypoint."""
    from oracle import c_oracle as O
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    from tests.helpers import oracle_env_from_table, random_table
    rng = np.random.default_rng(11)
    for trial, (world, nobs) in enumerate([(24, 150), (48, 1500)]):
        t64 = random_table(rng, nobs, world, sem_pf=True)
        env = B.Environment(L.SEM_PF, t64, 0.0)
        oenv = oracle_env_from_table(O.SEM_PF, t64, 0.0)
        hv = rng.integers(-3, 4, size=10000).astype(np.int32)
        src, dst = (0, 1, 2), (world - 1, world - 2, world // 2)
        for dodge in ("up", "left"):
            paths, lens = B.find_path_pf(env, src, dst, hv, dodge, cap=32)
            cap = paths.shape[1]
            epaths, elens, over = O.find_path_pf(oenv, src, dst, hv, dodge, cap=cap)
            assert over == 0
            assert np.array_equal(lens.cpu().numpy(), elens)
            got = paths.cpu().numpy()
            mask = np.arange(cap)[None, :] < elens[:, None]
            assert np.array_equal(got[mask], epaths[mask])

