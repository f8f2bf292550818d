"""Loader of the UNMODIFIED reference modules for bench.py's CPU legs (test / measurement infrastructure).

Looks in oracle/_ref/ (made by oracle/make_ref.py in the build container, travels to the GPU box) and then in
/root/reference.  HR imports as is; PF needs matplotlib stubbed (it is not installed and only plots); GAE runs a
whole GA and plt.show() at import (GAE:181-194), so only its class definitions are exec'd.
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
HR = "UAV_Navigation/path_planning_GA/path_planner_hrmaga.py"
PF = "UAV_Navigation/path_planning_GA/GA_complex _PF.py"
GAE = "uav-backend/app/routers/GA_enhanced.py"


def root():
    for r in (os.path.join(HERE, "_ref"), os.environ.get("TRACEFLY_REFERENCE", "/root/reference")):
        if os.path.exists(os.path.join(r, HR)) and os.path.exists(os.path.join(r, PF)):
            return r
    return None


def available():
    return root() is not None


def _stub_matplotlib():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_tf_stub", False):
        return
    for n in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m._tf_stub = True
            sys.modules[n] = m
    sys.modules["mpl_toolkits.mplot3d.art3d"].Poly3DCollection = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def _load(name, rel):
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(name, os.path.join(root(), rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_hr():
    return _load("_tracefly_ref_hr", HR)


def load_pf():
    _stub_matplotlib()
    return _load("_tracefly_ref_pf", PF)
