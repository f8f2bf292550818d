"""Island-model GA (BASELINE configs[4]): the population is split into islands,
one per GPU / process; islands evolve independently and exchange only their elites — an NCCL
all-gather over NVLink every `interval` generations.  Each island then overwrites its worst
(world-1)*E chromosomes with the foreign elites (source rank ascending, then elite rank).

The reference is single-population (HR:115 keeps 2 elites), so there is no reference behaviour
to match here: a 1-island run is exactly `evolve_batch`, and the N-island exchange is checked
against the CPU restatement (tests/test_islands_gloo.py, world_size 2 on gloo).

The data path never touches the host: pack (tf_pack_elites) -> all_gather_into_tensor ->
replace (tf_replace_worst) all run on the island's CUDA stream.  `ops` abstracts the four
device operations so the same scheduling / ordering logic runs on CPU tensors under gloo in
the tests (with the oracle's ops) — the product always uses `CudaOps`.
"""
import torch
import torch.distributed as dist

from . import _lib as L


class CudaOps:
    """The product path: every operation is a kernel of libtracefly_b200.so."""

    def __init__(self, env, params, P, W, k_rank, E, n_worst):
        import ctypes as C
        from . import batched as B
        self.B, self.C = B, C
        self.env, self.params = env, params
        self.topk_rank = B.TopK(P, k_rank)
        self.topk_elite = B.TopK(P, E) if E else None
        self.topk_worst = B.TopK(P, n_worst) if n_worst else None

    def fitness(self, pop):
        return self.B.fitness_batch(pop, None, self.env, self.params, want_collisions=False)[0]

    def rank(self, fit, out):
        return self.topk_rank(fit, out=out)

    def generation(self, pop, fit, top, ga, seed, gen, out_pop, out_fit):
        self.B.generation(self.env, pop, fit, top, ga, self.params, seed=seed, generation=gen, out_pop=out_pop,
                          out_fit=out_fit)

    def elites(self, fit):
        return self.topk_elite(fit)

    def worst(self, fit):
        return self.topk_worst(fit, worst=True)

    def pack(self, pop, fit, idx, rows, rfit):
        B, C = self.B, self.C
        L.check(L.lib.tf_pack_elites(B._ptr(pop), B._ptr(fit), B._ptr(idx), idx.numel(), pop.shape[1], B._ptr(rows),
                                     B._ptr(rfit), B._stream()), "tf_pack_elites")

    def replace(self, pop, fit, worst_idx, grows, gfit, world, rank, E):
        B = self.B
        L.check(L.lib.tf_replace_worst(B._ptr(pop), B._ptr(fit), B._ptr(worst_idx), pop.shape[1], B._ptr(grows),
                                       B._ptr(gfit), world, rank, E, B._stream()), "tf_replace_worst")


class IslandGA:
    """One island.  `pop` f32[P, W, 4] lives on this rank's device and is consumed (ping-pong)."""

    def __init__(self, pop, ga, seed=0, elites=1024, interval=10, group=None, ops=None, env=None, params=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        P, W, _ = pop.shape
        self.P, self.W = P, W
        self.E = min(elites, P // max(1, 2 * max(self.world - 1, 1))) if self.world > 1 else 0
        self.interval = interval
        self.ga = ga
        self.ga.stream_id = self.rank          # Philox streams are keyed by island
        self.seed = seed
        self.k_rank = max(min(ga.n_elite, P), min(ga.n_top, P) if ga.selection == L.SEL_TRUNCATION else 1, 1)
        n_worst = (self.world - 1) * self.E
        if ops is None:
            params = params or L.default_params()
            ops = CudaOps(env, params, P, W, self.k_rank, self.E, n_worst)
        self.ops = ops
        self.pop, self.nxt = pop, torch.empty_like(pop)
        self.fit = ops.fitness(pop)
        self.nfit = torch.empty_like(self.fit)
        self.top = torch.empty(self.k_rank, dtype=torch.int32, device=pop.device)
        self.gen = 0
        self.migrations = 0
        if self.world > 1:
            # one send / receive buffer: E rows of W*4 floats followed by E fitness values
            self.send = torch.empty(self.E * (W * 4 + 1), dtype=torch.float32, device=pop.device)
            self.recv = torch.empty(self.world * self.send.numel(), dtype=torch.float32, device=pop.device)

    def step(self):
        """One generation: rank -> select/crossover/mutate + fitness (fused); migrate when due."""
        self.ops.rank(self.fit, self.top)
        self.ops.generation(self.pop, self.fit, self.top, self.ga, self.seed, self.gen, self.nxt, self.nfit)
        self.pop, self.nxt = self.nxt, self.pop
        self.fit, self.nfit = self.nfit, self.fit
        self.gen += 1
        if self.world > 1 and self.interval > 0 and self.gen % self.interval == 0:
            self.migrate()

    def migrate(self):
        """pack elites -> all-gather (NCCL over NVLink in production) -> replace worst."""
        E, W = self.E, self.W
        if E == 0:
            return
        rows = self.send[:E * W * 4].view(E, W, 4)
        rfit = self.send[E * W * 4:]
        self.ops.pack(self.pop, self.fit, self.ops.elites(self.fit), rows, rfit)
        worst = self.ops.worst(self.fit)            # ranked BEFORE any replacement
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        per = self.send.numel()
        g = self.recv.view(self.world, per)
        if self.recv.is_cuda:
            # the kernels take [world][E][W][4] rows and [world][E] fitness: views into recv are
            # strided per rank, so hand them the two sub-blocks rank by rank-contiguous
            grows = g[:, :E * W * 4].contiguous()
            gfit = g[:, E * W * 4:].contiguous()
        else:
            grows, gfit = g[:, :E * W * 4].contiguous(), g[:, E * W * 4:].contiguous()
        self.ops.replace(self.pop, self.fit, worst, grows, gfit, self.world, self.rank, E)
        self.migrations += 1

    def best(self):
        f, i = torch.max(self.fit, dim=0)
        return float(f), self.pop[int(i)]
