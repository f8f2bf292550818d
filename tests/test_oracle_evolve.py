"""Replays the reference's own evolve() runs (HR:101-130) through the oracle, driven by the
draw tape recorded from np.random while the unmodified reference ran (tests/golden/gen_golden.py):
every generation's population and float64 fitness must be BIT-IDENTICAL, and so must the
returned best path (including the one-generation lag quirk of HR:126-128)."""
import json

import numpy as np
import pytest

from oracle import c_oracle as O
from oracle.evolve import evolve_hr
from tests.helpers import load_golden


@pytest.mark.parametrize("name", ["tiny", "odd_hot", "default", "cfg1"])
def test_evolve_replay_bit_exact(name):
    g, meta = load_golden(f"evolve_hr_{name}.npz")
    obs64 = O.obstacle_table(g["obs_pos"], g["obs_size"], dtype=np.float64)
    env = O.OracleEnv(O.SEM_HR, obs64.astype(np.float32), 2.0).with_obs64(obs64)
    best, pops, fits = evolve_hr(env, meta["start"], meta["end"], meta["P"], meta["G"], g["tape"], record=True)
    assert pops.shape[2] == meta["W"]
    assert np.array_equal(pops[..., :3], g["pops"].astype(np.float64)), "population trajectory differs"
    assert np.array_equal(fits, g["fits"]), "float64 fitness trajectory differs"
    assert np.array_equal(best, g["best"])


def test_pf_default_scenario_fitness():
    """PF:255-269 at seed 0: ragged population (lens 16..21), shaped obstacles, best_path()."""
    g, meta = load_golden("evolve_pf_default.npz")
    obs64 = O.obstacle_table(g["obs_pos"], g["obs_size"], g["obs_kind"], dtype=np.float64)
    obs32 = O.obstacle_table(g["obs_pos"], g["obs_size"], g["obs_kind"], dtype=np.float32)
    env = O.OracleEnv(O.SEM_PF, obs32).with_obs64(obs64)
    P, W, _ = g["wp"].shape
    pop = np.zeros((P, W, 4))
    pop[:, :, :3] = g["wp"]
    fit, _ = O.fitness_f64(env, pop, g["lens"], O.default_params(guard=0))
    assert np.array_equal(fit, g["fitness"])
    b = int(np.argmax(fit))  # max() returns the first maximal element (PF:223)
    assert np.array_equal(pop[b, :g["lens"][b], :3], g["best"])


def _find_path_cases():
    g, meta = load_golden("find_path_pf.npz")
    for scen in meta["scenarios"]:
        obs64 = O.obstacle_table(g[f"pos_{scen}"], g[f"size_{scen}"], g[f"kind_{scen}"], dtype=np.float64)
        obs32 = O.obstacle_table(g[f"pos_{scen}"], g[f"size_{scen}"], g[f"kind_{scen}"], dtype=np.float32)
        rows = np.nonzero(g["scen"] == scen)[0]
        yield g, scen, obs64, obs32, rows


def test_pf_find_path_matches_reference_walks():
    """PF:93-132 greedy walk + dodge + loop guard: 180 paths recorded from the unmodified PathPlanner
    (6 environments x 6 dodge strategies x 5 height variations), waypoint for waypoint."""
    for g, scen, obs64, obs32, rows in _find_path_cases():
        env = O.OracleEnv(O.SEM_PF, obs32).with_obs64(obs64)
        for dodge in O.DODGE_MOVES:
            sel = [r for r in rows if str(g["dodge"][r]) == dodge]
            cap = int(g["lens"][sel].max())
            paths, lens, over = O.find_path_pf(env, g[f"src_{scen}"], g[f"dst_{scen}"], g["hv"][sel], dodge, cap=cap)
            assert over == 0
            assert np.array_equal(lens, g["lens"][sel])
            for i, r in enumerate(sel):
                assert np.array_equal(paths[i, :lens[i], :3], g["wp"][r, :lens[i]])
        # a capacity below the needed length is reported, never overrun
        longest = rows[int(np.argmax(g["lens"][rows]))]
        if g["lens"][longest] > 2:
            cap = int(g["lens"][longest]) - 1
            paths, lens, over = O.find_path_pf(env, g[f"src_{scen}"], g[f"dst_{scen}"], g["hv"][[longest]],
                                               str(g["dodge"][longest]), cap=cap)
            assert over == 1 and lens[0] == cap
            assert np.array_equal(paths[0, :, :3], g["wp"][longest, :cap])
