"""GPU parity for ranking, selection / crossover / mutation and the fused generation.
Selected indices, crossover points and the produced populations must be BIT-EXACT:
  * against the reference itself under its own recorded draw tape (golden evolve traces), and
  * against the CPU oracle under identical Philox draws (extensions have no reference code)."""
import numpy as np
import pytest

from oracle import c_oracle as O
from tests.helpers import as_u32, load_golden, oracle_env_from_table, random_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tf():
    import torch
    from tracefly_b200 import _lib as L
    from tracefly_b200 import batched as B
    assert torch.cuda.is_available()
    return torch, L, B


@pytest.mark.parametrize("P,k", [(1, 1), (7, 7), (100, 10), (2048, 2), (2049, 10), (50_000, 10), (300_000, 1024),
                                 (1_000_003, 10), (70_000, 7168)])
@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_topk_stable_order(tf, P, k, dtype):
    torch, L, B = tf
    if dtype == "f64" and k > 4096:
        pytest.skip("f64 keys: k <= 4096")
    rng = np.random.default_rng(P + k)
    # heavy ties (few distinct values) + signed zeros + infinities: the stable order must hold
    vals = rng.integers(0, max(2, P // 50), size=P).astype(np.float64) / 7.0
    vals[rng.integers(0, P, size=max(1, P // 100))] = np.inf
    vals[rng.integers(0, P, size=max(1, P // 100))] = -0.0
    vals = vals.astype(np.float32 if dtype == "f32" else np.float64)
    fit = torch.from_numpy(vals).cuda()
    for worst in (False, True):
        got = B.rank_topk(fit, k, worst=worst).cpu().numpy()
        assert np.array_equal(got, O.topk(vals, k, worst=worst))


def test_all_equal_population_keeps_index_order(tf):
    """Generation 0 of HR is P identical paths (HR:82-83): the stable sort must return 0..k-1."""
    torch, L, B = tf
    fit = torch.full((5000,), 0.625, device="cuda")
    assert B.rank_topk(fit, 10).cpu().tolist() == list(range(10))
    assert B.rank_topk(fit, 3, worst=True).cpu().tolist() == [4999, 4998, 4997]


@pytest.mark.parametrize("name", ["tiny", "odd_hot", "default", "cfg1"])
def test_ga_step_replays_reference_generations(tf, name):
    """Each recorded generation of the unmodified reference: rank -> tape-driven GA step must
    reproduce the next population bit for bit, and the FP64 kernel its fitness bit for bit."""
    torch, L, B = tf
    g, meta = load_golden(f"evolve_hr_{name}.npz")
    P, G, W = meta["P"], meta["G"], meta["W"]
    table = np.zeros((len(g["obs_pos"]), 8))
    table[:, :3] = g["obs_pos"]
    table[:, 3] = np.where(np.isnan(g["obs_size"]), -1.0, g["obs_size"])
    env = B.Environment(L.SEM_HR, table, 2.0)
    pops = np.zeros((G + 1, P, W, 4))
    pops[..., :3] = g["pops"]
    ga = L.default_ga_params(mutation_rate=meta["rate"])
    k = min(10, P)
    gens = range(G) if G <= 40 else list(range(0, G, 7)) + [G - 1]
    for gen in gens:
        pop = torch.from_numpy(pops[gen]).cuda()
        fit = torch.from_numpy(g["fits"][gen]).cuda()
        top = B.rank_topk(fit, k)
        order = np.lexsort((np.arange(P), -g["fits"][gen]))
        assert np.array_equal(top.cpu().numpy(), order[:k])
        nxt, _, sel, cp = B.ga_step(pop, fit, top, ga, tape=g["tape"][gen])
        assert np.array_equal(nxt.cpu().numpy(), pops[gen + 1]), f"generation {gen}: population differs"
        assert np.array_equal(cp.cpu().numpy(), g["tape"][gen][:, 2])
        assert np.array_equal(sel.cpu().numpy(), order[:k][g["tape"][gen][:, :2]])
        f64, _ = B.fitness_batch_f64(nxt, None, env)
        assert np.array_equal(f64.cpu().numpy(), g["fits"][gen + 1]), f"generation {gen}: fitness differs"


COMBOS = [
    dict(selection=0, crossover=0, mutation=0),   # the reference's own operators, Philox draws
    dict(selection=1, crossover=0, mutation=0, tournament_k=4),
    dict(selection=1, crossover=1, mutation=2, tournament_k=7, sigma=0.75),
    dict(selection=0, crossover=1, mutation=1, n_top=32),
    dict(selection=1, crossover=0, mutation=1, tournament_k=32, n_elite=5),
    dict(selection=1, crossover=0, mutation=0, tournament_k=1, n_elite=0),
]


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("P,W", [(1001, 20), (4096, 64), (513, 100)])
def test_ga_step_philox_vs_oracle(tf, combo, P, W):
    torch, L, B = tf
    rng = np.random.default_rng(P + W)
    pop_h = O.init_walks(9, P, W, (32, 32, 32), (0, 0, 0), (63, 63, 63))
    fit_h = rng.integers(0, 300, size=P).astype(np.float32) / 301.0   # ties included
    kw = dict(mutation_rate=0.4, stream_id=3)
    kw.update(combo)
    ga, oga = L.default_ga_params(**kw), O.default_ga(**kw)
    k = max(min(ga.n_elite, P), ga.n_top if ga.selection == 0 else 1, 1)
    pop, fit = torch.from_numpy(pop_h).cuda(), torch.from_numpy(fit_h).cuda()
    top = B.rank_topk(fit, k)
    otop = O.topk(fit_h, k)
    assert np.array_equal(otop, top.cpu().numpy())
    nxt, _, sel, cp = B.ga_step(pop, fit, top, ga, seed=0xDEADBEEFCAFE, generation=17)
    onxt, _, osel, ocp = O.ga_step(pop_h, fit_h, otop, oga, seed=0xDEADBEEFCAFE, gen=17)
    assert np.array_equal(osel, sel.cpu().numpy()), "selected parent indices differ"
    assert np.array_equal(ocp, cp.cpu().numpy()), "crossover points differ"
    assert np.array_equal(onxt, nxt.cpu().numpy()), "children differ"
    if ga.n_elite:
        assert np.array_equal(nxt[:ga.n_elite].cpu().numpy(), pop_h[otop[:ga.n_elite]])


def test_ga_step_variable_lengths(tf):
    """PF/GAE populations are ragged: child1 = p1[:cp] + p2[cp:] has len(p2)."""
    torch, L, B = tf
    rng = np.random.default_rng(3)
    P, W = 600, 40
    pop_h = O.init_walks(4, P, W, (32, 32, 32), (0, 0, 0), (63, 63, 63))
    lens_h = rng.integers(3, W + 1, size=P).astype(np.int32)
    fit_h = rng.random(P).astype(np.float32)
    for mut in (0, 1):
        ga, oga = L.default_ga_params(mutation=mut, mutation_rate=0.5), O.default_ga(mutation=mut, mutation_rate=0.5)
        pop, fit, lens = torch.from_numpy(pop_h).cuda(), torch.from_numpy(fit_h).cuda(), torch.from_numpy(lens_h).cuda()
        top = B.rank_topk(fit, 10)
        nxt, nlens, sel, cp = B.ga_step(pop, fit, top, ga, seed=5, generation=2, lens=lens)
        onxt, onlens, osel, ocp = O.ga_step(pop_h, fit_h, O.topk(fit_h, 10), oga, seed=5, gen=2, lens=lens_h)
        assert np.array_equal(onlens, nlens.cpu().numpy())
        assert np.array_equal(osel, sel.cpu().numpy()) and np.array_equal(ocp, cp.cpu().numpy())
        assert np.array_equal(onxt, nxt.cpu().numpy())


@pytest.mark.parametrize("combo", COMBOS[:4])
@pytest.mark.parametrize("sem,W,N", [("hr", 20, 10), ("pf", 32, 1024), ("pf", 64, 200), ("none", 100, 0)])
def test_fused_generation_equals_unfused_and_oracle(tf, combo, sem, W, N):
    torch, L, B = tf
    P = 3001
    rng = np.random.default_rng(W + N)
    semc = {"hr": L.SEM_HR, "pf": L.SEM_PF, "none": L.SEM_NONE}[sem]
    table = random_table(rng, N, 64, sem_pf=(sem == "pf"))
    env = B.Environment(semc, table, 2.0)
    pop_h = O.init_walks(21, P, W, (32, 32, 32), (0, 0, 0), (63, 63, 63))
    pop = torch.from_numpy(pop_h).cuda()
    fit, _ = B.fitness_batch(pop, None, env)
    kw = dict(mutation_rate=0.3)
    kw.update(combo)
    ga = L.default_ga_params(**kw)
    k = max(min(ga.n_elite, P), ga.n_top if ga.selection == 0 else 1, 1)
    top = B.rank_topk(fit, k)
    nxt_f, fit_f, coll_f, sel_f, cp_f = B.generation(env, pop, fit, top, ga, seed=99, generation=5,
                                                     want_collisions=True, want_parity=True)
    nxt_u, _, sel_u, cp_u = B.ga_step(pop, fit, top, ga, seed=99, generation=5)
    fit_u, coll_u = B.fitness_batch(nxt_u, None, env)
    assert torch.equal(nxt_f, nxt_u) and torch.equal(fit_f, fit_u) and torch.equal(coll_f, coll_u)
    assert torch.equal(sel_f, sel_u) and torch.equal(cp_f, cp_u)
    oenv = oracle_env_from_table({"hr": O.SEM_HR, "pf": O.SEM_PF, "none": O.SEM_NONE}[sem], table, 2.0)
    fit_h = fit.cpu().numpy()
    onxt, _, osel, ocp = O.ga_step(pop_h, fit_h, O.topk(fit_h, k), O.default_ga(**kw), seed=99, gen=5)
    O.set_threads(8)
    ofit, ocoll = O.fitness_f32(oenv, onxt, None, O.default_params(guard=0 if sem == "pf" else 1))
    O.set_threads(1)
    assert np.array_equal(onxt, nxt_f.cpu().numpy()) and np.array_equal(osel, sel_f.cpu().numpy())
    assert np.array_equal(ofit, fit_f.cpu().numpy()) and np.array_equal(ocoll, as_u32(coll_f))


@pytest.mark.parametrize("P,W,k,world", [(400_000, 64, 2, "voxel"), (400_000, 64, 1, "voxel"), (400_000, 48, 2, "empty"),
                                          (400_000, 64, 3, "voxel"), (3001, 32, 2, "table"), (65_536, 32, 2, "table"), (65_536, 64, 1, "table"), (600_000, 32, 2, "flat")])
def test_generation_ranked_returns_the_ranking_of_the_children(tf, P, W, k, world):
    """tf_generation_ranked: next_top == tf_rank_topk(next_fit, k) == the oracle's stable order, whether the ranking
    is fused into the persistent kernel (large P, k <= 2), runs behind it (k = 3, small P, obstacle table), and when
    nearly all children tie ("flat": identical chromosomes, no mutation -> index order decides); and the generation
    itself is unchanged by it."""
    torch, L, B = tf
    rng = np.random.default_rng(P + W + k)
    if world == "table":
        env = B.Environment(L.SEM_PF, random_table(rng, 40, 64, sem_pf=True), 0.0)
    else:
        env = B.Environment.empty()
    if world == "voxel":
        env.set_voxels(B.voxelize(random_table(rng, 300, 64, sem_pf=True), (64, 64, 64), L.SEM_PF), (64, 64, 64))
    if world == "flat":
        pop = B.init_diagonal((0, 0, 0), (W - 1, 0, 0), P, W=W)[0]
    else:
        pop = B.init_walks(5, P, W, start=(32, 32, 32), lo=(0, 0, 0), hi=(63, 63, 63))
    prm = L.default_params(dist_zero_guard=0 if world == "table" else 1)
    fit, _ = B.fitness_batch(pop, None, env, prm)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.0 if world == "flat" else 0.2, n_elite=2)
    top = B.rank_topk(fit, 2)
    plan = B.TopK(P, k)
    ntop = torch.full((k,), -1, dtype=torch.int32, device="cuda")
    for rep in range(2):     # twice: the launch counters and the workspace are handed back clean
        nxt, nfit, *_ = B.generation(env, pop, fit, top, ga, prm, seed=3, generation=rep, next_top=ntop, topk=plan)
        ref_nxt, ref_fit, *_ = B.generation(env, pop, fit, top, ga, prm, seed=3, generation=rep)
        assert torch.equal(nxt, ref_nxt) and torch.equal(nfit, ref_fit)
        assert torch.equal(ntop, B.rank_topk(nfit, k))
        assert np.array_equal(ntop.cpu().numpy(), O.topk(nfit.cpu().numpy(), k))


def test_track_best_lag_quirk(tf):
    """HR:126-128: best follows population[0] (last generation's elite) with a strict '>'."""
    torch, L, B = tf
    W = 8
    tr = B.BestTracker(W)
    pop = torch.arange(3 * W * 4, dtype=torch.float32, device="cuda").reshape(3, W, 4)
    for gen, f0 in enumerate([0.5, 0.5, 0.7, 0.6]):
        fit = torch.tensor([f0, 9.0, 9.0], device="cuda")
        tr.update(pop + gen, fit, 0, gen)
    assert float(tr.best_fit) == pytest.approx(0.7) and int(tr.best_gen) == 2
    assert torch.equal(tr.best_path, (pop + 2)[0])


def test_init_populations(tf):
    torch, L, B = tf
    for start, end in [((0, 0, 0), (15, 15, 8)), ((0, 0, 0), (19, 19, 8)), ((2, 1, 0), (9, 12, 4)),
                       ((0.0, 0.0, 0.0), (2.886751345948129, 2.886751345948129, 0.7698003589195010)), ((3, 3, 3), (3, 3, 3))]:
        ref = O.init_diagonal(start, end)
        pop, n = B.init_diagonal(start, end, 37, f64=True)
        assert n == len(ref)
        assert np.array_equal(pop.cpu().numpy(), np.repeat(ref[None], 37, axis=0))
        pop32, _ = B.init_diagonal(start, end, 5)
        assert np.array_equal(pop32.cpu().numpy(), np.repeat(ref[None], 5, axis=0).astype(np.float32))
    starts = torch.rand((1000, 4), device="cuda") * 50
    starts = starts.floor()
    got = B.init_walks(77, 1000, 33, lo=(0, 0, 0), hi=(49, 49, 49), start_per_path=starts)
    exp = O.init_walks(77, 1000, 33, (0, 0, 0), (0, 0, 0), (49, 49, 49), start_pp=starts.cpu().numpy())
    assert np.array_equal(exp, got.cpu().numpy())


@pytest.mark.parametrize("P", [5003, 230003, 700001])   # tiles of 4 / of 32 (static) / dynamic tickets; ragged
@pytest.mark.parametrize("W", [3, 20, 32, 48, 63, 64, 70, 100, 128, 200, 256])
@pytest.mark.parametrize("env_kind", ["none", "pf", "vox+dyn+haz", "vox+dyn+haz, no table"])
def test_tma_stream_kernels_equal_lsu_kernels(tf, W, env_kind, P):
    """The TMA-staged kernels (W = 32 / 64 everywhere; any W <= 256 with masked lanes where there is no obstacle
    table), in both schedules (one chromosome per warp step / two), must give bit-identical results to the
    LSU-path kernels."""
    torch, L, B = tf
    rng = np.random.default_rng(W)
    dims = (64, 64, 64)
    table = random_table(rng, 150, 64, sem_pf=True)
    if env_kind == "none":
        env = B.Environment.empty()
    elif env_kind == "pf":
        env = B.Environment(L.SEM_PF, table, 0.0)
    else:
        env = B.Environment(L.SEM_HR, table[:40], 2.0) if env_kind == "vox+dyn+haz" else B.Environment.empty()
        env.set_voxels(B.voxelize(table, dims, L.SEM_PF), dims)
        dyn = np.zeros((16, 8), np.float32)
        dyn[:, :3] = rng.uniform(0, 64, size=(16, 3))
        dyn[:, 3] = 2
        dyn[:, 4:7] = rng.uniform(-1, 1, size=(16, 3))
        env.set_dynamic(dyn, 1.0)
        env.set_hazard(torch.from_numpy(rng.uniform(0, 1, size=(8, 16, 16)).astype(np.float32)).cuda(), 0.25, 0.25, 0, 0, 7.0 / (W - 1))
    prm = L.default_params(hazard_weight=0.3)
    prm3 = L.default_params(hazard_weight=0.3, samples_per_segment=3)
    pop = B.init_walks(5, P, W, start=(32, 32, 32), lo=(0, 0, 0), hi=(63, 63, 63))
    if P == 5003:   # off-lattice waypoints (what MAIN passes): fractional coordinates through the same kernels
        g = torch.Generator(device="cuda")
        g.manual_seed(W)
        pop[..., :3] += torch.rand(pop[..., :3].shape, device="cuda", generator=g) * 0.75 - 0.25
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.5, n_elite=3)
    gau = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, crossover=L.CX_UNIFORM, mutation=L.MUT_GAUSSIAN,
                              mutation_rate=0.5, sigma=0.75, n_elite=3)
    res = []
    for disable, schedule in ((0, 0), (0, 1), (0, 2), (1, 0)):
        L.set_option(L.OPT_DISABLE_STREAM, disable)
        L.set_option(L.OPT_STREAM_SCHEDULE, schedule)
        try:
            fit, coll = B.fitness_batch(pop, None, env, prm)
            top = B.rank_topk(fit, 3)
            out = B.generation(env, pop, fit, top, ga, prm, seed=77, generation=9, want_collisions=True, want_parity=True)
            # uniform crossover + Gaussian mutation: both parent rows staged (stream_kernel<..., UX>) for W = 32 / 64
            uni = B.generation(env, pop, fit, top, gau, prm, seed=78, generation=4, want_collisions=True, want_parity=True)
            # S = 3 samples per segment (XF 2: the one-chromosome-per-step TMA kernel for W = 32 / 64)
            fit3, coll3 = B.fitness_batch(pop, None, env, prm3)
            gen3 = B.generation(env, pop, fit3, B.rank_topk(fit3, 3), ga, prm3, seed=79, generation=2, want_collisions=True,
                                want_parity=True)
            torch.cuda.synchronize()
            res.append((fit, coll) + out + uni + (fit3, coll3) + gen3)
        finally:
            L.set_option(L.OPT_DISABLE_STREAM, 0)
            L.set_option(L.OPT_STREAM_SCHEDULE, 0)
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert torch.equal(a, b)


@pytest.mark.parametrize("fmt", ["f32", "xyz", "i16"])
def test_host_stepper_round_trip(tf, fmt):
    """Host-array API: pinned population in -> fitness + next population out equals the device API."""
    torch, L, B = tf
    P, W = 4099, 64
    rng = np.random.default_rng(1)
    table = random_table(rng, 60, 64, sem_pf=True)
    env = B.Environment(L.SEM_PF, table, 0.0)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=4, mutation_rate=0.2)
    prm = L.default_params(dist_zero_guard=0)
    pops = [B.init_walks(s, P, W, start=(32, 32, 32), lo=(0, 0, 0), hi=(63, 63, 63)) for s in (1, 2, 3)]
    stepper = B.HostStepper(env, P, W, ga, prm, fmt=fmt, depth=2)
    hdt = torch.int16 if fmt == "i16" else torch.float32
    hc = 3 if fmt == "xyz" else 4
    hosts = [p[..., :hc].to(hdt).contiguous().cpu().pin_memory() for p in pops]
    tickets = [stepper.submit(hosts[0], seed=5, generation=0), stepper.submit(hosts[1], seed=5, generation=1)]
    got = []
    for i in range(3):
        fit_h, next_h = stepper.result(tickets.pop(0))
        got.append((fit_h.clone(), next_h.clone()))
        if i == 0:
            tickets.append(stepper.submit(hosts[2], seed=5, generation=2))
    for g, (pop, (fit_h, next_h)) in enumerate(zip(pops, got)):
        fit, _ = B.fitness_batch(pop, None, env, prm)
        top = B.rank_topk(fit, 2)
        nxt, nfit, *_ = B.generation(env, pop, fit, top, ga, prm, seed=5, generation=g)
        assert torch.equal(fit.cpu(), fit_h)
        assert torch.equal(nxt[..., :hc].to(hdt).cpu(), next_h)
    h2d, d2h = stepper.bytes_per_step()
    assert h2d == P * W * {"f32": 16, "xyz": 12, "i16": 8}[fmt] and d2h == h2d + 4 * P


@pytest.mark.parametrize("P,k", [(1, 1), (5000, 37), (300_000, 1024), (1_000_003, 7168), (70_000, 69_999)])
def test_select_set_is_the_topk_set_in_index_order(tf, P, k):
    torch, L, B = tf
    rng = np.random.default_rng(P + k)
    vals = (rng.integers(0, max(2, P // 40), size=P).astype(np.float32) / 3.0)
    fit = torch.from_numpy(vals).cuda()
    sel = B.SelectSet(P, k)
    for worst in (False, True):
        got = sel(fit, worst=worst).cpu().numpy()
        assert np.array_equal(got, O.select_set(vals, k, worst=worst))


def test_select_set_special_values_and_reuse(tf):
    """One cooperative launch (round 2): all-equal fitness, +-0, +-inf, negative values, k == P, and a workspace
    that is reused back to back (the kernel hands it back zeroed)."""
    torch, L, B = tf
    rng = np.random.default_rng(3)
    P = 40_000
    cases = [np.full(P, 0.25, np.float32),
             np.where(rng.random(P) < 0.5, 0.0, -0.0).astype(np.float32),
             rng.choice(np.array([-np.inf, -1.5, -0.0, 0.0, 1e-30, 2.0, np.inf], np.float32), size=P),
             rng.standard_normal(P).astype(np.float32)]
    for k in (1, 2, 257, 4096, P - 1, P):
        sel = B.SelectSet(P, k)
        for vals in cases:
            fit = torch.from_numpy(vals).cuda()
            for worst in (False, True, True, False):
                got = sel(fit, worst=worst).cpu().numpy()
                assert np.array_equal(got, O.select_set(vals, k, worst=worst)), (k, worst)


def test_pf_operators_replay_the_reference_tape(tf):
    """PF GeneticAlgorithm.crossover / .mutate (PF:196-208) recorded from the unmodified reference
    (tests/golden/ops_pf.npz): ragged parents, crossover points beyond the second parent, mutated endpoints.
    tf_ga_step / tf_ga_step_f64 in tape mode must return the reference's children."""
    torch, L, B = tf
    from tests.test_oracle_ops import check_children, load_ops_pf
    g, pop, lens, tape = load_ops_pf()
    ga = L.default_ga_params(n_elite=0, selection=L.SEL_TOURNAMENT, mutation=L.MUT_LATTICE_PF)
    for dtype in (torch.float64, torch.float32):
        dpop = torch.from_numpy(pop).to("cuda", dtype)
        fit = torch.zeros(len(pop), dtype=dtype, device="cuda")
        nxt, nlens, sel, cp = B.ga_step(dpop, fit, torch.zeros(2, dtype=torch.int32, device="cuda"), ga,
                                        tape=tape, lens=torch.from_numpy(lens).cuda())
        check_children(g, nxt.double().cpu().numpy(), nlens.cpu().numpy())
        assert np.array_equal(cp.cpu().numpy(), g["tape"][:, 2])
        onxt, onlens, _, _ = O.ga_step(pop.astype(np.float64 if dtype == torch.float64 else np.float32), np.zeros(len(pop)),
                                       np.zeros(2, np.int32), O.default_ga(n_elite=0, selection=1, mutation=1), tape=tape, lens=lens)
        assert np.array_equal(onxt, nxt.cpu().numpy()) and np.array_equal(onlens, nlens.cpu().numpy())


def test_top2_selection_mode(tf):
    """TF_SEL_TOP2 = PF:192-194 select_parents, fused and unfused, against the oracle."""
    torch, L, B = tf
    rng = np.random.default_rng(8)
    P, W = 3000, 32
    table = random_table(rng, 30, 48, sem_pf=True)
    env, prm = B.Environment(L.SEM_PF, table, 0.0), L.default_params(dist_zero_guard=0)
    pop = B.init_walks(3, P, W, start=(24, 24, 24), lo=(0, 0, 0), hi=(47, 47, 47))
    fit, _ = B.fitness_batch(pop, None, env, prm)
    top = B.rank_topk(fit, 2)
    ga = L.default_ga_params(n_elite=2, selection=L.SEL_TOP2, mutation=L.MUT_LATTICE_PF, mutation_rate=0.4)
    oga = O.default_ga(n_elite=2, selection=2, mutation=1, mutation_rate=0.4)
    onxt, _, osel, ocp = O.ga_step(pop.cpu().numpy(), fit.cpu().numpy(), top.cpu().numpy(), oga, seed=21, gen=4)
    nxt, _, sel, cp = B.ga_step(pop, fit, top, ga, seed=21, generation=4)
    assert np.array_equal(osel, sel.cpu().numpy()) and np.array_equal(ocp, cp.cpu().numpy())
    assert np.array_equal(onxt, nxt.cpu().numpy())
    fnxt, ffit, _, fsel, fcp = B.generation(env, pop, fit, top, ga, prm, seed=21, generation=4, want_parity=True)
    assert torch.equal(fnxt, nxt) and torch.equal(fsel, sel) and torch.equal(fcp, cp)
    assert (sel[:, 0] == top[0]).all() and (sel[:, 1] == top[1]).all()
    bad = L.default_ga_params(selection=3)
    with pytest.raises(ValueError):
        B.ga_step(pop, fit, top, bad)


@pytest.mark.parametrize("P", [2, 50, 4097])
def test_roulette_selection_kernels(tf, P):
    """tf_roulette_cumweights + tf_select_roulette (GAE:28-34) == the oracle == CPython's random.choices."""
    import random
    torch, L, B = tf
    rng = np.random.default_rng(P)
    score = -rng.integers(20, 300, size=P).astype(np.int32)
    cum = B.roulette_cumweights(torch.from_numpy(score).cuda())
    assert np.array_equal(cum.cpu().numpy(), O.roulette_cumweights(score))
    total = int(score.sum())
    random.seed(P)
    st = random.getstate()
    picks = random.choices(range(P), [int(v) / total for v in score], k=999)
    random.setstate(st)
    u = np.array([random.random() for _ in range(999)])
    got = B.select_roulette(cum, torch.from_numpy(u).cuda()).cpu().numpy()
    assert list(got) == picks and np.array_equal(got, O.select_roulette(O.roulette_cumweights(score), u))


def test_waypoint_codecs_on_device(tf):
    """tf_encode_waypoints == the unmodified reference functions (tests/golden/waypoint_codecs.npz) == the oracle, from
    float64 paths and from a float32 population; the host-side zlib of the device bytes is the reference's stream."""
    import json
    import os
    import zlib
    torch, L, B = tf
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waypoint_codecs.npz"))
    for case in json.loads(str(g["meta"]))["cases"]:
        w = g[f"in_{case}"]
        p4 = np.zeros((1, len(w), 4))
        p4[0, :, :3] = w
        q = B.encode_waypoints(torch.from_numpy(p4).cuda(), L.CODEC_QUANT_F32, 0.05).cpu().numpy()[0]
        assert np.array_equal(q.view(np.uint32), g[f"q32_{case}"].view(np.uint32)), case
        if json.loads(str(g["meta"]))["zlib"] == zlib.ZLIB_VERSION:
            assert np.array_equal(np.frombuffer(zlib.compress(q.tobytes()), np.uint8), g[f"zlib_{case}"])
        d = B.encode_waypoints(torch.from_numpy(p4).cuda(), L.CODEC_DELTA_ROUND, 3).cpu().numpy()[0]
        assert np.array_equal(d.view(np.uint64), g[f"dq_{case}"].view(np.uint64)), case
    pop = B.init_walks(2, 500, 64, start=(100, 100, 100), lo=(0, 0, 0), hi=(1023, 1023, 1023))
    for mode, prm in ((L.CODEC_QUANT_F32, 0.05), (L.CODEC_DELTA_ROUND, 3)):
        got = B.encode_waypoints(pop, mode, prm).cpu().numpy()
        assert np.array_equal(got, O.encode_waypoints(pop.cpu().numpy().astype(np.float64), mode, prm))
    with pytest.raises(ValueError):
        B.encode_waypoints(pop, L.CODEC_QUANT_F32, 0.0)


def test_malformed_tapes_are_clamped_not_trusted(tf):
    """A host tape with out-of-range ranks / indices / crossover points (evolve_batch(tapes=...) accepts user tapes)
    is clamped on the device — no out-of-bounds parent gather — and the oracle clamps the same way."""
    torch, L, B = tf
    rng = np.random.default_rng(2)
    P, W = 600, 32
    pop = B.init_walks(4, P, W, start=(9, 9, 9), lo=(0, 0, 0), hi=(30, 30, 30))
    fit = torch.from_numpy(rng.random(P).astype(np.float32)).cuda()
    npairs = (P - 2 + 1) // 2
    tape = np.zeros((npairs, 8), np.int32)
    tape[:, 0] = rng.integers(-5, 2000, size=npairs)
    tape[:, 1] = rng.integers(-2_000_000, 2_000_000, size=npairs)
    tape[:, 2] = rng.integers(-10, 700, size=npairs)
    tape[:, 3] = rng.integers(-1, 40, size=npairs)
    tape[:, 5] = rng.integers(-1, 40, size=npairs)
    tape[:, 4] = tape[:, 6] = 0b100100
    env, prm = B.Environment.empty(), L.default_params()
    for sel_mode, k in ((L.SEL_TRUNCATION, 10), (L.SEL_TOURNAMENT, 2), (L.SEL_TOP2, 2)):
        top = B.rank_topk(fit, k)
        ga = L.default_ga_params(n_elite=2, selection=sel_mode, n_top=10)
        oga = O.default_ga(n_elite=2, selection=int(sel_mode), n_top=10)
        onxt, _, osel, ocp = O.ga_step(pop.cpu().numpy(), fit.cpu().numpy(), top.cpu().numpy(), oga, tape=tape)
        nxt, _, sel, cp = B.ga_step(pop, fit, top, ga, tape=tape)
        assert np.array_equal(osel, sel.cpu().numpy()) and np.array_equal(ocp, cp.cpu().numpy())
        assert np.array_equal(onxt, nxt.cpu().numpy())
        fnxt, _, _, fsel, fcp = B.generation(env, pop, fit, top, ga, prm, tape=tape, want_parity=True)
        assert torch.equal(fnxt, nxt) and torch.equal(fsel, sel) and torch.equal(fcp, cp)
        assert int(sel.min()) >= 0 and int(sel.max()) < P and int(cp.min()) >= 0 and int(cp.max()) <= W


def test_generation_falls_back_to_the_unfused_pair_for_large_tables(tf):
    """More than 2048 analytic obstacles and no candidate index: tf_generation answers TF_ERR_UNSUPPORTED and
    batched.generation / evolve_batch take tf_ga_step + tf_fitness — same results as the oracle."""
    torch, L, B = tf
    rng = np.random.default_rng(6)
    table = random_table(rng, 2500, 64, sem_pf=True)
    env, prm = B.Environment(L.SEM_PF, table, 0.0), L.default_params(dist_zero_guard=0)
    oenv = oracle_env_from_table(O.SEM_PF, table, 0.0)
    P, W = 300, 32
    pop = B.init_walks(4, P, W, start=(30, 30, 30), lo=(0, 0, 0), hi=(63, 63, 63))
    fit, _ = B.fitness_batch(pop, None, env, prm)
    top = B.rank_topk(fit, 2)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.5)
    assert L.lib.tf_generation(env.handle, pop.data_ptr(), fit.data_ptr(), top.data_ptr(), None, 1, 2, ga, prm, P, W,
                               torch.empty_like(pop).data_ptr(), torch.empty_like(fit).data_ptr(), None, None, None,
                               None) == L.TF_ERR_UNSUPPORTED
    nxt, nfit, ncoll, sel, cp = B.generation(env, pop, fit, top, ga, prm, seed=1, generation=2, want_collisions=True, want_parity=True)
    oga = O.default_ga(selection=1, tournament_k=3, mutation_rate=0.5)
    onxt, _, osel, ocp = O.ga_step(pop.cpu().numpy(), fit.cpu().numpy(), top.cpu().numpy(), oga, seed=1, gen=2)
    onfit, oncoll = O.fitness_f32(oenv, onxt, None, O.default_params(guard=0))
    assert np.array_equal(onxt, nxt.cpu().numpy()) and np.array_equal(osel, sel.cpu().numpy())
    assert np.array_equal(onfit, nfit.cpu().numpy()) and np.array_equal(oncoll, as_u32(ncoll))
    res = B.evolve_batch(pop, env, 3, ga, prm, seed=1)
    assert res["generation"] == 3


def test_evolve_batch_resume_is_bit_identical(tf, tmp_path):
    """Checkpoint / resume: Philox is counter-based, so 12 generations == 5 + resume + 7."""
    torch, L, B = tf
    rng = np.random.default_rng(0)
    table = random_table(rng, 40, 64, sem_pf=True)
    env = B.Environment(L.SEM_PF, table, 0.0)
    pop = B.init_walks(2, 3000, 32, start=(32, 32, 32), lo=(0, 0, 0), hi=(63, 63, 63))
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, mutation_rate=0.3)
    full = B.evolve_batch(pop, env, 12, ga, seed=9, history=True)
    part = B.evolve_batch(pop, env, 5, ga, seed=9)
    state = B.checkpoint(part, str(tmp_path / "ckpt.pt"))
    rest = B.evolve_batch(None, env, 7, ga, resume=torch.load(str(tmp_path / "ckpt.pt")))
    assert rest["generation"] == 12 and state["generation"] == 5
    assert torch.equal(full["pop"], rest["pop"]) and torch.equal(full["fit"], rest["fit"])
    assert torch.equal(full["best_path"], rest["best_path"]) and torch.equal(full["best_fit"], rest["best_fit"])
    assert full["best_history"].shape == (12,) and bool((full["best_history"][1:] >= full["best_history"][:-1]).all())
    # emergency stop hook (HR:110-112)
    calls = []
    assert B.evolve_batch(pop, env, 50, ga, should_stop=lambda: len(calls) >= 3 or calls.append(1)) is None


def test_island_exchange_kernels_match_the_oracle(tf):
    """The device side of an elite migration (islands.CudaOps: tf_select_set -> tf_pack_elites -> [all-gather] ->
    tf_select_set -> tf_replace_worst) against the CPU restatement, with three islands emulated on one GPU (the
    all-gather is a stack); then a single-island IslandGA run against the plain rank + generation loop."""
    torch, L, B = tf
    from tracefly_b200.islands import CudaOps, IslandGA
    world, P, W, E = 3, 4001, 32, 37
    rng = np.random.default_rng(12)
    table = random_table(rng, 40, 48, sem_pf=True)
    env, prm = B.Environment(L.SEM_PF, table, 0.0), L.default_params(dist_zero_guard=0)
    pops = [B.init_walks(20 + r, P, W, start=(24, 24, 24), lo=(0, 0, 0), hi=(47, 47, 47)) for r in range(world)]
    fits = [B.fitness_batch(p, None, env, prm, want_collisions=False)[0] for p in pops]
    fits[1][::7] = fits[1][3]                                           # ties among elites and victims
    ops = CudaOps(env, prm, P, W, 2, E, (world - 1) * E)
    ref_pop = [p.cpu().numpy().copy() for p in pops]
    ref_fit = [f.cpu().numpy().copy() for f in fits]
    rows = torch.empty((world, E, W, 4), device="cuda")
    rfit = torch.empty((world, E), device="cuda")
    idx = torch.empty(E, dtype=torch.int32, device="cuda")
    for r in range(world):
        ops.elites(fits[r], idx)
        ops.pack(pops[r], fits[r], idx, rows[r], rfit[r])
    worst = torch.empty((world - 1) * E, dtype=torch.int32, device="cuda")
    for r in range(world):
        ops.worst(fits[r], worst)
        ops.replace(pops[r], fits[r], worst, rows, rfit, world, r, E)
    torch.cuda.synchronize()
    packs = [O.pack_elites(ref_pop[r], ref_fit[r], O.select_set(ref_fit[r], E, worst=False)) for r in range(world)]
    grows, gfit = np.stack([p[0] for p in packs]), np.stack([p[1] for p in packs])
    assert np.array_equal(grows, rows.cpu().numpy()) and np.array_equal(gfit, rfit.cpu().numpy())
    for r in range(world):
        O.replace_worst(ref_pop[r], ref_fit[r], O.select_set(ref_fit[r], (world - 1) * E, worst=True), grows, gfit, world, r, E)
        assert np.array_equal(ref_pop[r], pops[r].cpu().numpy()), f"island {r}: population differs"
        assert np.array_equal(ref_fit[r], fits[r].cpu().numpy())
    # one island: IslandGA.step == rank + fused generation
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.3)
    start = B.init_walks(5, P, W, start=(24, 24, 24), lo=(0, 0, 0), hi=(47, 47, 47))
    isl = IslandGA(start.clone(), ga, seed=9, env=env, params=prm)
    pop, fit = start.clone(), B.fitness_batch(start, None, env, prm, want_collisions=False)[0]
    for g in range(4):
        isl.step()
        pop, fit, *_ = B.generation(env, pop, fit, B.rank_topk(fit, 2), ga, prm, seed=9, generation=g)
    assert isl.migrations == 0 and torch.equal(isl.pop, pop) and torch.equal(isl.fit, fit)
