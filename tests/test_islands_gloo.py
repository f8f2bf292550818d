"""Island model host logic on CPU: world_size 2 over gloo.  The device operations
are swapped for the oracle's (tests may use the oracle; the product always uses CudaOps), so this
checks the scheduling, the all-gather layout and the deterministic replacement order:
the j-th foreign elite (source rank ascending, then elite index) overwrites the j-th worst."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

P, W, E, INTERVAL, GENS, SEED = 256, 20, 8, 3, 7, 42


class OracleOps:
    """Same interface as tracefly_b200.islands.CudaOps, CPU tensors, oracle arithmetic."""

    def __init__(self, oenv, k_rank, E, n_worst):
        from oracle import c_oracle as O
        self.O, self.oenv, self.k_rank, self.E, self.n_worst = O, oenv, k_rank, E, n_worst

    def fitness(self, pop):
        return torch.from_numpy(self.O.fitness_f32(self.oenv, pop.numpy())[0])

    def rank(self, fit, out):
        out.copy_(torch.from_numpy(self.O.topk(fit.numpy(), self.k_rank)))

    def generation(self, pop, fit, top, ga, seed, gen, out_pop, out_fit):
        oga = self.O.default_ga(n_elite=ga.n_elite, selection=ga.selection, n_top=ga.n_top, tournament_k=ga.tournament_k,
                                crossover=ga.crossover, mutation=ga.mutation, mutation_rate=ga.mutation_rate,
                                sigma=ga.sigma, stream_id=ga.stream_id)
        nxt, _, _, _ = self.O.ga_step(pop.numpy(), fit.numpy(), top.numpy(), oga, seed=seed, gen=gen)
        out_pop.copy_(torch.from_numpy(nxt))
        out_fit.copy_(torch.from_numpy(self.O.fitness_f32(self.oenv, nxt)[0]))

    def elites(self, fit, out):
        out.copy_(torch.from_numpy(self.O.select_set(fit.numpy(), self.E, worst=False)))
        return out

    def worst(self, fit, out):
        out.copy_(torch.from_numpy(self.O.select_set(fit.numpy(), self.n_worst, worst=True)))
        return out

    def pack(self, pop, fit, idx, rows, rfit):
        r, f = self.O.pack_elites(pop.numpy(), fit.numpy(), idx.numpy())
        rows.copy_(torch.from_numpy(r))
        rfit.copy_(torch.from_numpy(f))

    def replace(self, pop, fit, worst_idx, grows, gfit, world, rank, E):
        p, f = pop.numpy(), fit.numpy()   # share memory with the tensors
        self.O.replace_worst(p, f, worst_idx.numpy(), grows.numpy(), gfit.numpy(), world, rank, E)


def _world(rank):
    from oracle import c_oracle as O
    rng = np.random.default_rng(5)
    obs = O.obstacle_table(rng.integers(0, 32, size=(12, 3)), rng.integers(1, 4, size=12), rng.integers(0, 3, size=12))
    oenv = O.OracleEnv(O.SEM_PF, obs, 0.0)
    pop = O.init_walks(SEED + rank, P, W, (16, 16, 16), (0, 0, 0), (31, 31, 31))
    return O, oenv, pop


def _worker(rank, world, port, outdir, pipelined):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tracefly_b200 import _lib as L
    from tracefly_b200.islands import IslandGA
    O, oenv, pop = _world(rank)
    ga = L.default_ga_params(selection=L.SEL_TOURNAMENT, tournament_k=3, mutation_rate=0.3)
    ops = OracleOps(oenv, 2, E, (world - 1) * E)
    isl = IslandGA(torch.from_numpy(pop), ga, seed=SEED, elites=E, interval=INTERVAL, ops=ops, pipelined=pipelined)
    assert isl.E == E and isl.world == world and isl.ga.stream_id == rank
    isl.warm_up()               # a throw-away pipelined exchange: must leave the island as it was
    for g in range(GENS):
        isl.step()
        if g == 1:
            isl.warm_up()
    isl.flush()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), pop=isl.pop.numpy(), fit=isl.fit.numpy(), mig=isl.migrations)
    dist.destroy_process_group()


def _simulate(world, pipelined):
    """The single-process restatement of the island schedule (oracle/islands_sim.py) on this test's world."""
    from oracle import islands_sim
    return islands_sim.simulate(world, lambda r: (_world(r)[1], None, _world(r)[2]),
                                dict(selection=1, tournament_k=3, mutation_rate=0.3), SEED, GENS, INTERVAL, E, pipelined)


@pytest.mark.parametrize("pipelined", [True, False])
def test_two_islands_over_gloo(tmp_path, pipelined):
    import __graft_entry__ as g
    g.build()
    world = 2
    port = 29500 + (os.getpid() % 2000) + (7 if pipelined else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), pipelined), nprocs=world, join=True)
    ref = _simulate(world, pipelined)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert int(got["mig"]) == GENS // INTERVAL
        assert np.array_equal(got["pop"], ref[r]["pop"]), f"rank {r}: population differs"
        assert np.array_equal(got["fit"], ref[r]["fit"])
    # migration really moved chromosomes: rank 0's best elites appear in rank 1
    assert not np.array_equal(ref[0]["pop"], ref[1]["pop"])
