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
