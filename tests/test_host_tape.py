"""tf_np_hr_tape (host function of the C-ABI, no GPU work): the native restatement of NumPy's legacy generator
must hand out exactly the draws HRMAGAGenerator.evolve takes from np.random (HR:117-120) and leave the generator in
the same state — this is what lets the drop-in pre-draw whole blocks of generations."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from tracefly_b200 import _lib
    return _lib


def reference_order_draws(G, npairs, n_top, W, rate):
    """The calls of HR:118, HR:86, HR:94-96 in the reference's order, through np.random itself."""
    out = np.zeros((G, npairs, 8), np.int32)
    rnd = np.random
    for g in range(G):
        for j in range(npairs):
            pa, pb = rnd.choice(n_top, 2, replace=False)
            row = [pa, pb, rnd.randint(1, W - 1), -1, 0, -1, 0, 0]
            for c in range(2):
                if rnd.rand() < rate:
                    row[3 + 2 * c] = rnd.randint(1, W - 1)
                    d = rnd.choice([-1, 0, 1], 3)
                    row[4 + 2 * c] = int((d[0] + 1) | ((d[1] + 1) << 2) | ((d[2] + 1) << 4))
            out[g, j] = row
    return out


@pytest.mark.parametrize("seed,G,npairs,n_top,W,rate", [
    (0, 200, 49, 10, 20, 0.1),      # BASELINE configs[0]
    (1, 7, 3, 5, 3, 0.9),           # W = 3: randint(1, 2) consumes no draw
    (2, 50, 24, 10, 4, 0.5),
    (3, 30, 100, 10, 64, 0.3),
    (4, 3, 2, 2, 16, 1.0),          # two ranked parents only
    (5, 20, 10, 10, 250, 0.0),      # gate never passes
    (6, 400, 11, 10, 33, 0.25),     # crosses several MT19937 state regenerations
])
def test_native_tape_equals_numpy_legacy_stream(L, seed, G, npairs, n_top, W, rate):
    np.random.seed(seed)
    want = reference_order_draws(G, npairs, n_top, W, rate)
    probe_want = np.random.rand(3)
    np.random.seed(seed)
    st = np.random.get_state()
    key = np.ascontiguousarray(st[1], np.uint32).copy()
    pos = C.c_int32(int(st[2]))
    got = np.zeros((G, npairs, 8), np.int32)
    assert L.lib.tf_np_hr_tape(key.ctypes.data_as(C.c_void_p), C.byref(pos), G, npairs, n_top, W, rate,
                               got.ctypes.data_as(C.c_void_p)) == 0
    np.random.set_state((st[0], key, pos.value, st[3], st[4]))
    assert np.array_equal(got, want)
    assert np.array_equal(np.random.rand(3), probe_want)      # the generator continues where the reference's would


def test_native_tape_rejects_bad_arguments(L):
    key = np.zeros(624, np.uint32)
    pos = C.c_int32(0)
    out = np.zeros((1, 1, 8), np.int32)
    f = L.lib.tf_np_hr_tape
    kp, op = key.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)
    assert f(None, C.byref(pos), 1, 1, 10, 20, 0.1, op) == L.TF_ERR_INVALID
    assert f(kp, C.byref(pos), 1, 1, 1, 20, 0.1, op) == L.TF_ERR_SHAPE       # fewer than two ranked parents
    assert f(kp, C.byref(pos), 1, 1, 10, 2, 0.1, op) == L.TF_ERR_SHAPE       # W < 3: randint(1, 1) raises in the reference
    pos = C.c_int32(625)
    assert f(kp, C.byref(pos), 1, 1, 10, 20, 0.1, op) == L.TF_ERR_SHAPE


def test_dropin_probe_accepts_the_native_tape(L):
    from tracefly_b200 import path_planner_hrmaga as H
    assert H._native_tape_ok() is True
