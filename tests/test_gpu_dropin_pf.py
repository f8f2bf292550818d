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
