"""HR:101-130 evolve() restated in array form on top of the C oracle (f64 reference-order model).
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py."""
import numpy as np

from . import c_oracle as O


def evolve_hr(env, start, end, P, G, tapes, n_elite=2, n_top=10, record=False):
    """tapes: int32[G, npairs, 8] draw tape (tf_draw records).  Returns (best_path [W,3] or None,
    pops [G+1,P,W,4], fits [G+1,P]) — pops/fits only when record=True."""
    init = O.init_diagonal(start, end)                     # HR:68-80
    W = len(init)
    pop = np.repeat(init[None], P, axis=0)                 # HR:82-83: P identical paths
    prm = O.default_params(guard=1)
    fit, _ = O.fitness_f64(env, pop, None, prm)            # HR:103-104
    pops, fits = [pop.copy()], [fit.copy()]
    best_fit, best = -1.0, None                            # HR:106-107
    ga = O.default_ga(n_elite=n_elite, n_top=n_top)
    for g in range(G):
        top = O.topk(fit, min(max(n_top, n_elite), P))      # HR:114 stable descending sort
        pop, _, _, _ = O.ga_step(pop, fit, top, ga, tape=tapes[g])  # HR:115-122
        fit, _ = O.fitness_f64(env, pop, None, prm)        # HR:123-124 (elites re-evaluated)
        if fit[0] > best_fit:                              # HR:126-128 (population[0] = last gen's elite)
            best_fit, best = fit[0], pop[0, :, :3].copy()
        if record:
            pops.append(pop.copy())
            fits.append(fit.copy())
    if record:
        return best, np.stack(pops), np.stack(fits)
    return best, None, None
