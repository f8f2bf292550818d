"""Island-model GA (BASELINE configs[4]): the population is split into islands,
one per GPU / process; islands evolve independently and exchange only their elites — an NCCL
all-gather over NVLink every `interval` generations.  Each island then overwrites its worst
(world-1)*E chromosomes with the foreign elites (source rank ascending, then elite index; victims
taken in ascending index order).

Pipelined migration: the elites of generation g (the E best as a set, ascending index order) are
selected and packed on the main stream (one tf_select_set launch + one pack kernel, ~40 us at 1 M
chromosomes), all-gathered on a SIDE stream while the main stream already computes generation g+1,
and they replace the worst of generation g+1 as soon as it exists (tf_select_set + tf_replace_worst).
Round 1 also ran the elite selection on the side stream: it then shared the SMs with the persistent
generation kernel, finished late, and the step that applied the exchange waited 180-370 us for it
(profiles/README.md, round 2).  The schedule is deterministic (it does not depend on timing), and a
synchronous `migrate()` is kept for warm-up / final flush.

The reference is single-population (HR:115 keeps 2 elites), so there is no reference behaviour
to match here: a 1-island run is exactly `evolve_batch`, and the N-island exchange is checked
against the CPU restatement (tests/test_islands_gloo.py, world_size 2 on gloo).

The data path never touches the host: tf_select_set -> tf_pack_elites -> all_gather_into_tensor ->
tf_select_set -> tf_replace_worst.  `ops` abstracts the device operations so the same scheduling
logic runs on CPU tensors under gloo in the tests (with the oracle's ops) — the product always
uses `CudaOps`.
"""
import contextlib
import os

import torch
import torch.distributed as dist

from . import _lib as L


class CudaOps:
    """The product path: every operation is a kernel of libtracefly_b200.so."""

    def __init__(self, env, params, P, W, k_rank, E, n_worst):
        from . import batched as B
        self.B = B
        self.env, self.params = env, params
        self.topk_rank = B.TopK(P, k_rank)
        self.best_set = B.SelectSet(P, E) if E else None
        self.worst_set = B.SelectSet(P, n_worst) if n_worst else None
        self.side = torch.cuda.Stream() if n_worst else None
        self.reserve_sms = int(os.environ.get("TRACEFLY_RESERVE_SMS", 8))
        self._ga_beside = self._ga_beside_src = None

    def fitness(self, pop):
        return self.B.fitness_batch(pop, None, self.env, self.params, want_collisions=False)[0]

    def rank(self, fit, out):
        return self.topk_rank(fit, out=out)

    fused_rank = True   # generation(next_top=...) also ranks the children (tf_generation_ranked)

    def generation(self, pop, fit, top, ga, seed, gen, out_pop, out_fit, beside_exchange=False, next_top=None):
        # while an elite exchange is in flight on the side stream, the persistent generation kernel leaves a
        # few SMs free for it (it would otherwise hold every CTA slot until it ends).  The number travels in
        # the launch's own tf_ga_params copy: no library-wide state is touched.
        if beside_exchange and self.reserve_sms > 0:
            if self._ga_beside is None or self._ga_beside_src is not ga:
                self._ga_beside = L.GaParams.from_buffer_copy(ga)
                self._ga_beside_src = ga
            self._ga_beside.stream_id = ga.stream_id
            self._ga_beside.reserve_sms = self.reserve_sms
            ga = self._ga_beside
        self.B.generation(self.env, pop, fit, top, ga, self.params, seed=seed, generation=gen, out_pop=out_pop,
                          out_fit=out_fit, next_top=next_top, topk=self.topk_rank if next_top is not None else None)

    def elites(self, fit, out):
        return self.best_set(fit, worst=False, out=out)

    def worst(self, fit, out):
        return self.worst_set(fit, worst=True, out=out)

    def pack(self, pop, fit, idx, rows, rfit):
        B = self.B
        L.check(L.lib.tf_pack_elites(B._ptr(pop), B._ptr(fit), B._ptr(idx), idx.numel(), pop.shape[1], B._ptr(rows),
                                     B._ptr(rfit), B._stream()), "tf_pack_elites")

    def replace(self, pop, fit, worst_idx, grows, gfit, world, rank, E):
        B = self.B
        L.check(L.lib.tf_replace_worst(B._ptr(pop), B._ptr(fit), B._ptr(worst_idx), pop.shape[1], B._ptr(grows),
                                       B._ptr(gfit), world, rank, E, B._stream()), "tf_replace_worst")

    # ---- stream plumbing of the pipelined migration
    def on_side_stream(self):
        self.side.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.side)

    def join_side_stream(self):
        torch.cuda.current_stream().wait_stream(self.side)


class IslandGA:
    """One island.  `pop` f32[P, W, 4] lives on this rank's device and is consumed (ping-pong)."""

    def __init__(self, pop, ga, seed=0, elites=1024, interval=10, group=None, ops=None, env=None, params=None,
                 pipelined=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        P, W, _ = pop.shape
        self.P, self.W = P, W
        self.E = min(elites, P // max(1, 2 * max(self.world - 1, 1))) if self.world > 1 else 0
        self.interval = interval
        self.pipelined = pipelined
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
        dev = pop.device
        self.top = torch.empty(self.k_rank, dtype=torch.int32, device=dev)
        # the generation kernel ranks its own children (ops.fused_rank): `ntop` receives the next generation's top_idx
        # and `_top_valid` says that `top` already ranks the current population (false at the start and after foreign
        # elites were written into it)
        self.ntop = torch.empty_like(self.top) if getattr(ops, "fused_rank", False) else None
        self._top_valid = False
        self.gen = 0
        self.hook = None
        self.migrations = 0
        self._pending = False
        if self.world > 1 and self.E > 0:
            E = self.E
            self.elite_idx = torch.empty(E, dtype=torch.int32, device=dev)
            self.worst_idx = torch.empty(n_worst, dtype=torch.int32, device=dev)
            self.send_rows = torch.empty((E, W, 4), dtype=torch.float32, device=dev)
            self.send_fit = torch.empty(E, dtype=torch.float32, device=dev)
            self.recv_rows = torch.empty((self.world, E, W, 4), dtype=torch.float32, device=dev)
            self.recv_fit = torch.empty((self.world, E), dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ one generation
    def step(self):
        """rank -> select/crossover/mutate + fitness (fused); then the migration schedule.
        `self.hook`, if set, is called as hook("generation_begin" | "generation_end", self) right around the
        launch of the generation kernel (bench.py records its CUDA events there)."""
        if not self._top_valid:
            self.ops.rank(self.fit, self.top)
        if self.hook is not None:
            self.hook("generation_begin", self)
        self.launch_generation()
        if self.hook is not None:
            self.hook("generation_end", self)
        self.pop, self.nxt = self.nxt, self.pop
        self.fit, self.nfit = self.nfit, self.fit
        if self.ntop is not None:
            self.top, self.ntop = self.ntop, self.top
            self._top_valid = True
        self.gen += 1
        self.after_generation()

    def launch_generation(self):
        """The fused generation kernel of the current population into the spare buffers."""
        kw = {"next_top": self.ntop} if self.ntop is not None else {}
        if self._pending and isinstance(self.ops, CudaOps):
            kw["beside_exchange"] = True
        self.ops.generation(self.pop, self.fit, self.top, self.ga, self.seed, self.gen, self.nxt, self.nfit, **kw)

    def after_generation(self):
        if self.world <= 1 or self.E == 0 or self.interval <= 0:
            return
        if self._pending:                       # elites gathered during this generation arrive now
            self._finish_exchange()
        if self.gen % self.interval == 0:
            if self.pipelined:
                self._start_exchange()
            else:
                self.migrate()

    # ------------------------------------------------------------------ migration
    def _side(self):
        return self.ops.on_side_stream() if hasattr(self.ops, "on_side_stream") else contextlib.nullcontext()

    def _start_exchange(self):
        """elites of the CURRENT population -> packed (main stream) -> all-gathered (side stream on the GPU)."""
        self.ops.elites(self.fit, self.elite_idx)
        self.ops.pack(self.pop, self.fit, self.elite_idx, self.send_rows, self.send_fit)
        with self._side():
            # output = the inputs of all ranks concatenated along dim 0 (accepted by NCCL and gloo alike)
            dist.all_gather_into_tensor(self.recv_rows.view(self.world * self.E, self.W, 4), self.send_rows, group=self.group)
            dist.all_gather_into_tensor(self.recv_fit.view(self.world * self.E), self.send_fit, group=self.group)
        self._pending = True

    def _finish_exchange(self):
        """foreign elites overwrite the worst of the CURRENT population."""
        if hasattr(self.ops, "join_side_stream"):
            self.ops.join_side_stream()
        self.ops.worst(self.fit, self.worst_idx)
        self.ops.replace(self.pop, self.fit, self.worst_idx, self.recv_rows, self.recv_fit, self.world, self.rank, self.E)
        self._top_valid = False       # foreign elites may outrank the local best
        self._pending = False
        self.migrations += 1

    def warm_up(self):
        """One throw-away pass through the PIPELINED exchange — elite selection and packing, the all-gather on the side
        stream beside a generation kernel that leaves SMs free for it, the victim selection — without changing the
        island: the generation writes into the spare buffers (the next step overwrites them), the gathered rows are
        never applied.  NCCL connects and sizes its channels for this message on the first collective that runs
        beside a kernel, and that cost (several hundred microseconds at 8 ranks) otherwise lands in the first real
        exchange.  Callers that time generations call it once before they start."""
        if self.world <= 1 or self.E == 0 or self.interval <= 0 or not self.pipelined or self._pending:
            return
        if not self._top_valid:
            self.ops.rank(self.fit, self.top)
            self._top_valid = self.ntop is not None
        self.ops.elites(self.fit, self.elite_idx)
        self.ops.pack(self.pop, self.fit, self.elite_idx, self.send_rows, self.send_fit)
        with self._side():
            dist.all_gather_into_tensor(self.recv_rows.view(self.world * self.E, self.W, 4), self.send_rows, group=self.group)
            dist.all_gather_into_tensor(self.recv_fit.view(self.world * self.E), self.send_fit, group=self.group)
        self._pending = True
        self.launch_generation()          # into nxt / nfit / ntop: not swapped in
        self._pending = False
        if hasattr(self.ops, "join_side_stream"):
            self.ops.join_side_stream()
        self.ops.worst(self.fit, self.worst_idx)

    def migrate(self):
        """Synchronous exchange on the current population (warm-up, non-pipelined mode)."""
        if self.world <= 1 or self.E == 0:
            return
        if self._pending:
            self._finish_exchange()
        self._start_exchange()
        self._finish_exchange()

    def flush(self):
        """Apply an exchange that is still in flight (call before reading the final population)."""
        if self._pending:
            self._finish_exchange()

    def best(self):
        self.flush()
        f, i = torch.max(self.fit, dim=0)
        return float(f), self.pop[int(i)]
