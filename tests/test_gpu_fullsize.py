"""Parity at BASELINE.json's full sizes (configs[1]: 65,536 x 32 vs 1,024 analytic obstacles; configs[2]:
1,048,576 x 64 vs a 1024^3 voxel grid; configs[3]: the same + hazard map + 1,024 moving spheres).  The CPU oracle cannot sweep these in seconds, so the checks
are (a) the oracle on a seeded sample of the same launch, and (b) size-independent properties:
batch independence, permutation equivariance, voxel flags == analytic flags, TMA path == LSU path,
fused == unfused, elites carried over, children are splices of the recorded parents, stable top-k."""
import numpy as np
import pytest

import bench
from oracle import c_oracle as O
from tests.helpers import as_u32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tf():
    import torch
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    assert torch.cuda.is_available()
    return torch, L, B


def _oracle_env(cfg, bricks_h=None):
    kw = {}
    sem, obs32 = O.SEM_NONE, None
    if cfg["n_analytic"]:
        t = bench.obstacle_table(cfg, cfg["n_analytic"], 4, bench.SEED + 1)
        obs32, sem = O.obstacle_table(t[:, :3], t[:, 3], t[:, 7].astype(np.int32)), O.SEM_PF
    if bricks_h is not None:
        kw.update(bricks=bricks_h, dims=cfg["world"])
    return O.OracleEnv(sem, obs32, 0.0, **kw)


def test_cfg2_full_size(tf):
    torch, L, B = tf
    cfg = dict(bench.WORKLOADS["cfg2"])
    P, W = cfg["P"], cfg["W"]
    env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
    fit, coll = B.fitness_batch(pop, None, env, params)
    # (a) oracle on a seeded sample of the same launch
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(P, 1024, replace=False))
    O.set_threads(8)
    ofit, ocoll = O.fitness_f32(_oracle_env(cfg), pop[idx].cpu().numpy(), None, O.default_params(guard=0))
    O.set_threads(1)
    assert np.array_equal(ofit, fit[idx].cpu().numpy()) and np.array_equal(ocoll, as_u32(coll[idx]))
    # (b) batch independence + permutation equivariance
    perm = torch.randperm(P, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    fit_p, coll_p = B.fitness_batch(pop[perm].contiguous(), None, env, params)
    assert torch.equal(fit_p, fit[perm]) and torch.equal(coll_p, coll[perm])
    sub = pop[1000:1000 + 4097].contiguous()
    fit_s, _ = B.fitness_batch(sub, None, env, params)
    assert torch.equal(fit_s, fit[1000:1000 + 4097])
    # TMA-staged path == LSU path at full size
    L.set_option(L.OPT_DISABLE_STREAM, 1)
    try:
        fit_l, coll_l = B.fitness_batch(pop, None, env, params)
    finally:
        L.set_option(L.OPT_DISABLE_STREAM, 0)
    assert torch.equal(fit_l, fit) and torch.equal(coll_l, coll)
    # one full generation: stable top-k, elites, children are splices of the recorded parents
    top = B.rank_topk(fit, 10)
    order = torch.sort(fit, descending=True, stable=True).indices[:10]
    assert torch.equal(top.long(), order)
    nxt, nfit, ncoll, sel, cp = B.generation(env, pop, fit, top, ga, params, seed=3, generation=1,
                                             want_collisions=True, want_parity=True)
    nxt_u, _, sel_u, cp_u = B.ga_step(pop, fit, top, ga, seed=3, generation=1)
    fit_u, coll_u = B.fitness_batch(nxt_u, None, env, params)
    assert torch.equal(nxt, nxt_u) and torch.equal(nfit, fit_u) and torch.equal(ncoll, coll_u) and torch.equal(sel, sel_u)
    assert torch.equal(nxt[:2], pop[top[:2].long()])
    j = torch.arange(0, 2000, device="cuda")
    c1 = nxt[2 + 2 * j]
    pa, pb = pop[sel[j, 0].long()], pop[sel[j, 1].long()]
    k = torch.arange(W, device="cuda")[None, :, None]
    splice = torch.where(k < cp[j].long()[:, None, None], pa, pb)
    differs = (c1 != splice).any(dim=2).sum(dim=1)            # only the (optional) mutated waypoint may differ
    assert int(differs.max()) <= 1
    assert 0.0 < float((differs == 1).float().mean()) < 0.25   # mutation rate 0.1 (minus zero deltas)


def test_cfg3_full_size(tf):
    torch, L, B = tf
    cfg = dict(bench.WORKLOADS["cfg3"])
    P, W = cfg["P"], cfg["W"]
    env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
    table = bench.obstacle_table(cfg, cfg["n_voxel_obs"], 10, bench.SEED + 2)
    fit, coll = B.fitness_batch(pop, None, env, params)
    assert float(coll.ne(0).any(dim=1).float().mean()) > 0.05      # the grid is really being hit
    # voxel flags == analytic flags (lattice paths; obstacles rasterised from the same table), on a slice
    n = 1 << 14
    env_a = B.Environment(L.SEM_PF, table, 0.0)
    _, coll_a = B.fitness_batch(pop[:n].contiguous(), None, env_a, params)
    assert torch.equal(coll_a, coll[:n])
    # oracle (voxeliser + FP32 model) on a seeded sample
    venv = O.OracleEnv(O.SEM_PF, O.obstacle_table(table[:, :3], table[:, 3], table[:, 7].astype(np.int32)), 0.0)
    bricks_h = O.voxelize(venv, *cfg["world"])
    assert np.array_equal(bricks_h, as_u32(env._keep["bricks"]))
    idx = np.sort(np.random.default_rng(3).choice(P, 4096, replace=False))
    O.set_threads(8)
    ofit, ocoll = O.fitness_f32(_oracle_env(cfg, bricks_h), pop[idx].cpu().numpy(), None, O.default_params(guard=1))
    O.set_threads(1)
    assert np.array_equal(ofit, fit[idx].cpu().numpy()) and np.array_equal(ocoll, as_u32(coll[idx]))
    # TMA path == LSU path, fused == unfused, elites, stable ranking — at 1M x 64
    top = B.rank_topk(fit, 10)
    assert torch.equal(top.long(), torch.sort(fit, descending=True, stable=True).indices[:10])
    worst = B.rank_topk(fit, 1024, worst=True)
    assert torch.equal(worst.long(), torch.sort(fit, descending=True, stable=True).indices.flip(0)[:1024])
    nxt, nfit, *_ = B.generation(env, pop, fit, top, ga, params, seed=5, generation=2)
    L.set_option(L.OPT_DISABLE_STREAM, 1)
    try:
        fit_l, coll_l = B.fitness_batch(pop, None, env, params)
        nxt_l, nfit_l, *_ = B.generation(env, pop, fit, top, ga, params, seed=5, generation=2)
    finally:
        L.set_option(L.OPT_DISABLE_STREAM, 0)
    assert torch.equal(fit_l, fit) and torch.equal(coll_l, coll) and torch.equal(nxt_l, nxt) and torch.equal(nfit_l, nfit)
    assert torch.equal(nxt[:2], pop[top[:2].long()])
    refit, _ = B.fitness_batch(nxt, None, env, params)
    assert torch.equal(refit, nfit)                                  # the fused sweep scored what it wrote
    # the whole next population (1M x 64: selection, crossover points, mutations, elites) vs the CPU restatement
    oga = O.default_ga(selection=ga.selection, tournament_k=ga.tournament_k, crossover=ga.crossover, mutation=ga.mutation,
                       mutation_rate=ga.mutation_rate, n_elite=ga.n_elite, stream_id=ga.stream_id)
    O.set_threads(16)
    onxt, _, osel, ocp = O.ga_step(pop.cpu().numpy(), fit.cpu().numpy(), top.cpu().numpy(), oga, seed=5, gen=2)
    O.set_threads(1)
    assert np.array_equal(onxt, nxt.cpu().numpy())
    _, _, _, sel, cp = B.generation(env, pop, fit, top, ga, params, seed=5, generation=2, want_parity=True)
    assert np.array_equal(osel, sel.cpu().numpy()) and np.array_equal(ocp, cp.cpu().numpy())


def test_cfg4_full_size(tf):
    """configs[3]: cfg3 + hazard map + 1,024 moving spheres (candidate index with a margin).  The oracle on a
    seeded sample of the full launch, index == brute force, and the fused generation scores what it writes —
    on a population evolved until waypoints have drifted past the world box."""
    torch, L, B = tf
    cfg = dict(bench.WORKLOADS["cfg4"])
    P, W = cfg["P"], cfg["W"]
    env, params, pop, ga = bench.build_gpu_workload(cfg, 0, torch, B, L)
    fit, coll = B.fitness_batch(pop, None, env, params)
    top = B.rank_topk(fit, 2)
    for g in range(12):                                               # let mutations push waypoints around
        pop, fit, *_ = B.generation(env, pop, fit, top, ga, params, seed=bench.SEED, generation=g)
        top = B.rank_topk(fit, 2)
    fit, coll = B.fitness_batch(pop, None, env, params)
    nxt, nfit, *_ = B.generation(env, pop, fit, top, ga, params, seed=5, generation=99)
    refit, _ = B.fitness_batch(nxt, None, env, params)
    assert torch.equal(refit, nfit)
    # oracle (brute force over the spheres, its own hazard sampler) on a sample that includes out-of-box chromosomes
    wx, wy, wz = cfg["world"]
    outside = ((pop[..., :3] < 0) | (pop[..., :3] >= torch.tensor([wx, wy, wz], device="cuda"))).any(-1).any(-1)
    idx = np.random.default_rng(4).choice(P, 2048, replace=False)
    idx = np.unique(np.concatenate([idx, torch.nonzero(outside).flatten()[:512].cpu().numpy()]))
    table = bench.obstacle_table(cfg, cfg["n_voxel_obs"], 10, bench.SEED + 2)
    venv = O.OracleEnv(O.SEM_PF, O.obstacle_table(table[:, :3], table[:, 3], table[:, 7].astype(np.int32)), 0.0)
    bricks_h = O.voxelize(venv, *cfg["world"])
    oenv = O.OracleEnv(O.SEM_NONE, None, 0.0, bricks=bricks_h, dims=cfg["world"], dyn=bench.dyn_table(cfg, cfg["n_dyn"], bench.SEED),
                       dyn_safety=2.0, hazard=bench.hazard_field(), hazard_map=(0.25, 0.25, 0.0, 0.0, 23.0 / (W - 1)))
    O.set_threads(16)
    ofit, ocoll = O.fitness_f32(oenv, pop[idx].cpu().numpy(), None, O.default_params(guard=1, hazard_weight=0.5))
    O.set_threads(1)
    assert np.array_equal(ofit, fit[idx].cpu().numpy()) and np.array_equal(ocoll, as_u32(coll[idx]))
