"""Single-process CPU restatement of the island model — TEST INFRASTRUCTURE like the rest of
oracle/: the checker of tests/test_islands_gloo.py, of the multi-GPU `-m gpu` test and of bench.py's N > 1
correctness gate.  The reference is single-population (HR:115 keeps 2 elites), so this restatement is what
defines the N-island behaviour: every `interval` generations each island's E best (as a set under the stable
order, listed by ascending index) are gathered from all islands; island r overwrites its (world - 1) * E worst — victims in ascending index order —
with the foreign elites ordered by source rank, then elite index.  Pipelined schedule: the elites of generation
g replace the worst of generation g + 1."""
import numpy as np

from . import c_oracle as O


def simulate(world, make_island, ga_kw, seed, gens, interval, E, pipelined, k_rank=2):
    """make_island(rank) -> (OracleEnv, params or None, pop f32[P, W, 4]).  Returns a list of dict(pop, fit)."""
    st = []
    for r in range(world):
        env, prm, pop = make_island(r)
        st.append(dict(env=env, prm=prm, pop=pop, fit=O.fitness_f32(env, pop, None, prm)[0]))
    pending = None

    def insert(packs):
        grows = np.stack([p[0] for p in packs])
        gfit = np.stack([p[1] for p in packs])
        for r, s in enumerate(st):
            worst = O.select_set(s["fit"], (world - 1) * E, worst=True)
            O.replace_worst(s["pop"], s["fit"], worst, grows, gfit, world, r, E)

    for gen in range(gens):
        for r, s in enumerate(st):
            ga = O.default_ga(stream_id=r, **ga_kw)
            s["pop"], _, _, _ = O.ga_step(s["pop"], s["fit"], O.topk(s["fit"], k_rank), ga, seed=seed, gen=gen)
            s["fit"] = O.fitness_f32(s["env"], s["pop"], None, s["prm"])[0]
        if world > 1 and E > 0 and interval > 0:
            if pending is not None:
                insert(pending)
                pending = None
            if (gen + 1) % interval == 0:
                packs = [O.pack_elites(s["pop"], s["fit"], O.select_set(s["fit"], E, worst=False)) for s in st]
                if pipelined:
                    pending = packs
                else:
                    insert(packs)
    if pending is not None:
        insert(pending)
    return st
