"""Pins oracle/tf_oracle.c against vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, generator: tests/golden/gen_golden.py).

  - f64 reference-order model: BIT-EXACT fitness, collision flags and nearest distances.
  - f32 fixed-order model (what the CUDA throughput kernels compute): collision flags
    bit-exact, fitness within 1e-5 relative (north_star tolerance).
"""
import numpy as np
import pytest

from oracle import c_oracle as O
from tests.helpers import golden_case, load_golden, rel_err, unpack_bits, valid_mask

FIT_RTOL = 1e-5  # BASELINE.json north_star: "fitness must agree within 1e-5 relative"


def _cases(fam):
    _, meta = load_golden(f"fitness_{fam}.npz")
    return meta["cases"]


@pytest.mark.parametrize("fam,case", [("hr", c) for c in _cases("hr")] + [("pf", c) for c in _cases("pf")])
def test_fitness_models_vs_reference(fam, case):
    g, _ = load_golden(f"fitness_{fam}.npz")
    d = golden_case(g, fam, case)
    sem = O.SEM_HR if fam == "hr" else O.SEM_PF
    env = O.OracleEnv(sem, d["obs32"], 2.0).with_obs64(d["obs64"])
    prm = O.default_params(guard=1 if fam == "hr" else 0)  # HR:53 guard vs PF:163 none
    P, W, _ = d["pop64"].shape
    valid = valid_mask(d["lens"], W)

    f64, c64, dn = O.fitness_f64(env, d["pop64"], d["lens"], prm, want_dnear=True)
    assert np.array_equal(f64, d["fitness"]), "f64 reference-order model must be bit-identical to the reference"
    assert np.array_equal(unpack_bits(c64, W) & valid, d["collide"] & valid)
    assert np.array_equal(np.where(valid, dn, 0), np.where(valid, d["dnear"], 0))

    f32, c32 = O.fitness_f32(env, d["pop64"].astype(np.float32), d["lens"], prm)
    assert np.array_equal(unpack_bits(c32, W) & valid, d["collide"] & valid), "collision flags must be bit-exact"
    assert rel_err(f32, d["fitness"]) <= FIT_RTOL


def test_threads_do_not_change_results():
    g, _ = load_golden("fitness_hr.npz")
    d = golden_case(g, "hr", "lat_W32_N64")
    env = O.OracleEnv(O.SEM_HR, d["obs32"], 2.0)
    pop = d["pop64"].astype(np.float32)
    a = O.fitness_f32(env, pop)
    O.set_threads(4)
    try:
        b = O.fitness_f32(env, pop)
    finally:
        O.set_threads(1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    assert O.philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert O.philox(f, f, f, f, f, f) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
