"""tracefly_b200 — B200-native (sm_100a) implementation of TraceFly's GA hot path:
batched fitness evaluation of a path population + selection / crossover / mutation.

Import name: `tracefly_b200` (the directory is named after the upstream repository,
`tracefly-ga-uav-pathplanning_b200/`, which is not a valid Python identifier; the
`tracefly_b200/` shim at the repo root loads this package under the importable name).

Sub-modules
    _lib                  ctypes binding of libtracefly_b200.so (C-ABI, include/tracefly_b200.h)
    batched               array-form API: fitness_batch, ga_step, generation, evolve_batch, ...
    path_planner_hrmaga   drop-in for UAV_Navigation/path_planning_GA/path_planner_hrmaga.py
    ga_complex_pf         drop-in for UAV_Navigation/path_planning_GA/GA_complex _PF.py
    ga_enhanced           drop-in for uav-backend/app/routers/GA_enhanced.py (GeneticAlgorithm.run)
    islands               island-model GA: one island per GPU, NCCL elite all-gather

There is no CPU fallback and no multi-backend dispatch: importing without the built
shared library raises ImportError; calling without a B200 raises.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (raises loudly when libtracefly_b200.so is missing)
