"""Shared helpers for the oracle / parity tests."""
import json
import os

import numpy as np

from oracle import c_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    return g, json.loads(str(g["meta"]))


def golden_case(g, fam, case):
    """-> dict(pop64 [P,W,4], lens, obs64 [N,8], obs32, fitness, collide[P,W] bool, dnear[P,W])"""
    wp = g[f"{case}/wp"]
    P, W, _ = wp.shape
    kind = g[f"{case}/obs_kind"] if fam == "pf" else None
    pop64 = np.zeros((P, W, 4))
    pop64[:, :, :3] = np.nan_to_num(wp)
    return dict(
        pop64=pop64, lens=g[f"{case}/lens"],
        obs64=O.obstacle_table(g[f"{case}/obs_pos"], g[f"{case}/obs_size"], kind, dtype=np.float64),
        obs32=O.obstacle_table(g[f"{case}/obs_pos"], g[f"{case}/obs_size"], kind, dtype=np.float32),
        fitness=g[f"{case}/fitness"], collide=g[f"{case}/collide"], dnear=g[f"{case}/dnear"])


def unpack_bits(words, W):
    """u32[P, ceil(W/32)] -> bool[P, W]"""
    words = np.asarray(words, np.uint32)
    i = np.arange(W)
    return ((words[:, i >> 5] >> (i & 31).astype(np.uint32)) & 1).astype(bool)


def valid_mask(lens, W):
    return np.arange(W)[None, :] < np.asarray(lens)[:, None]


def rel_err(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(a), fin), "inf/nan pattern differs"
    assert np.array_equal(a[~fin], ref[~fin]) or np.all(np.isnan(ref[~fin]) == np.isnan(a[~fin]))
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - ref[fin]) / np.abs(ref[fin])))


def tape_from_golden(tape_g):
    """[npairs, 8] int32 golden tape rows are already in tf_draw layout."""
    return np.ascontiguousarray(tape_g, np.int32)


# ----------------------------------------------------------------------------- GPU-side helpers
def gpu_env_from_table(B, L, sem, table64, safety=2.0):
    return B.Environment(sem, table64, safety)


def random_table(rng, N, world, sem_pf=True, max_size=4, aabb=False):
    t = np.zeros((N, 8))
    t[:, :3] = rng.integers(0, world, size=(N, 3))
    t[:, 3] = rng.integers(1, max_size + 1, size=N)
    if sem_pf:
        t[:, 7] = rng.integers(0, 4 if aabb else 3, size=N)
        t[:, 4:7] = rng.integers(1, max_size + 1, size=(N, 3))
    return t


def oracle_env_from_table(sem, table64, safety=2.0, **kw):
    kind = table64[:, 7].astype(np.int32)
    obs32 = O.obstacle_table(table64[:, :3], table64[:, 3], kind, half=table64[:, 4:7])
    obs64 = table64.copy()
    return O.OracleEnv(sem, obs32, safety, **kw).with_obs64(obs64)


def as_u32(t):
    return t.cpu().numpy().view(np.uint32)
