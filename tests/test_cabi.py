"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/tracefly_b200.h declares, and fails loudly without a device (no fallback path)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from tracefly_b200 import _lib
    return _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "tracefly_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(L):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L.lib, n), f"{n} declared in include/tracefly_b200.h but not exported"
    assert set(L.EXPORTED) == set(names), "ctypes signature table out of sync with the header"


def test_struct_layouts_match_header(L):
    assert C.sizeof(L.Params) == 32 and C.sizeof(L.GaParams) == 40
    p = L.default_params()
    assert [round(v, 6) for v in (p.w_safety, p.w_energy, p.w_wind, p.w_time, p.w_turn)] == [0.4, 0.2, 0.2, 0.1, 0.1]
    g = L.default_ga_params()
    assert (g.n_elite, g.n_top, round(g.mutation_rate, 6)) == (2, 10, 0.1)   # HR:115, HR:118, HR:62


def test_host_side_entry_points(L):
    assert L.lib.tf_version() == 100
    assert L.philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert b"fallback" in L.lib.tf_strerror(L.TF_ERR_NODEVICE)
    assert L.lib.tf_voxel_grid_bytes(1024, 1024, 1024) == 128 * 1024 * 1024
    assert L.lib.tf_voxel_grid_bytes(10, 8, 8) < 0
    n = C.c_int(0)
    s, e = (C.c_double * 3)(0, 0, 0), (C.c_double * 3)(19, 19, 8)
    assert L.lib.tf_init_diagonal(s, e, 0, 0, None, C.byref(n), None) == 0 and n.value == 20   # cfg 1: W = 20


def test_no_device_means_loud_failure(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.lib.tf_env_create_f64(L.SEM_HR, None, 0, 2.0, C.byref(h)) == L.TF_ERR_NODEVICE
    from tracefly_b200 import batched as B
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.Environment.empty()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tracefly-ga-uav-pathplanning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "tf_oracle" not in text or f.endswith((".cu", ".cuh")), f
