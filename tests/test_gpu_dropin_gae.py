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


def test_find_path_gae_kernel_vs_reference_walks():
    """tf_find_path_gae against the walks of the unmodified UAVPathPlanner (tests/golden/find_path_gae.npz), fed the
    recorded shuffle orders through the C-ABI, and through batched.find_path_gae (natively drawn tape under the same
    random.seed): same paths, same generator state afterwards."""
    import ctypes as C
    import json
    import os
    import random
    import torch
    from tracefly_b200 import _lib as L, batched as B
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "find_path_gae.npz"))
    for seed in json.loads(str(g["meta"]))["scenarios"]:
        wp, lens, sh = g[f"wp_{seed}"], g[f"lens_{seed}"], g[f"shuffles_{seed}"]
        obs = np.zeros((len(g[f"obs_{seed}"]), 4), np.int32)
        obs[:, :3] = g[f"obs_{seed}"]
        P, cap = len(lens), int(lens.max()) + 3
        paths = torch.zeros((P, cap, 4), dtype=torch.int32, device="cuda")
        dl = torch.zeros(P, dtype=torch.int32, device="cuda")
        res = torch.zeros(2, dtype=torch.int32, device="cuda")
        i3 = lambda v: (C.c_int32 * 3)(*[int(x) for x in v])
        b6 = (C.c_int32 * 6)(*[int(v) for v in g[f"bounds_{seed}"].reshape(6)])
        shd = torch.from_numpy(np.ascontiguousarray(sh if len(sh) else np.zeros(1, np.int32))).cuda()
        obs_d = torch.from_numpy(obs).cuda()
        L.check(L.lib.tf_find_path_gae(obs_d.data_ptr(), len(obs), b6, i3(g[f"src_{seed}"]), i3(g[f"dst_{seed}"]),
                                       shd.data_ptr(), len(sh), P, cap, paths.data_ptr(), dl.data_ptr(), res.data_ptr(), None), "tf_find_path_gae")
        assert res.cpu().tolist() == [len(sh), 0]
        assert np.array_equal(dl.cpu().numpy(), lens)
        got = paths.cpu().numpy()
        for i in range(P):
            assert np.array_equal(got[i, :lens[i], :3], wp[i, :lens[i]]), (seed, i)
        random.seed(int(seed))
        trails = B.find_path_gae([tuple(o) for o in g[f"obs_{seed}"].tolist()], g[f"bounds_{seed}"].tolist(),
                                 g[f"src_{seed}"].tolist(), g[f"dst_{seed}"].tolist(), P, cap=8)   # forces the capacity retry
        assert [len(t) for t in trails] == lens.tolist()
        assert all(np.array_equal(np.array(t), wp[i, :lens[i]]) for i, t in enumerate(trails))
        assert random.random() == float(g[f"after_{seed}"])
