"""Drop-in for the backend GA (uav-backend/app/routers/GA_enhanced.py): same seeds -> same result as
the unmodified reference (tests/golden/run_gae.npz), and tf_fitness_gae vs the CPU restatement."""
import random

import numpy as np
import pytest

from oracle import c_oracle as O
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def test_backend_ga_run_matches_reference(capsys):
    import torch
    assert torch.cuda.is_available()
    from tracefly_b200 import ga_enhanced as gae
    g, meta = load_golden("run_gae.npz")
    random.seed(meta["seed"])
    np.random.seed(meta["seed"])
    env = gae.Environment((0, 30), (0, 30), (0, 15))
    assert sorted(env.obstacles) == [tuple(o) for o in g["obstacles"].tolist()]
    ga = gae.GeneticAlgorithm(population_size=meta["P"], generations=meta["G"], mutation_rate=meta["rate"])
    best = ga.run((0, 0, 0), (30, 30, 15), env)                           # GAE:60-66 / SIMR:24
    printed = [int(l.split(":")[-1]) for l in capsys.readouterr().out.strip().splitlines()]
    assert printed == g["best_per_generation"].tolist()
    assert np.array_equal(np.array(best, float), g["best"])
    assert [len(p) for p in ga.population] == g["lens"].tolist()
    for p, wp, n, f in zip(ga.population, g["wp"], g["lens"], g["fitness"]):
        assert np.array_equal(np.array(p, float), wp[:n]) and ga.fitness_function(p) == f
    assert all(isinstance(v, int) for w in best for v in w)               # JSON-serialisable (SIMR:26-29)


def test_fitness_gae_kernel_vs_oracle():
    import torch
    from tracefly_b200 import batched as B
    rng = np.random.default_rng(0)
    P, W = 5000, 97
    pop = np.zeros((P, W, 4), np.float32)
    pop[:, :, :3] = rng.integers(0, 3, size=(P, W, 3))
    lens = rng.integers(0, W + 1, size=P).astype(np.int32)
    got = B.fitness_gae(torch.from_numpy(pop).cuda(), torch.from_numpy(lens).cuda()).cpu().numpy()
    assert np.array_equal(got, O.fitness_gae(pop, lens))
    got = B.fitness_gae(torch.from_numpy(pop).cuda(), None).cpu().numpy()
    assert np.array_equal(got, O.fitness_gae(pop, None))
