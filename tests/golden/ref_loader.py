"""Load the UNMODIFIED TraceFly reference modules read-only from /root/reference.

Only used by tests/golden/gen_golden.py (run in the build container, where
/root/reference exists) and by the optional `-m "not gpu"` live cross-check
tests, which skip when the reference tree is absent (e.g. on the GPU box).

HR = UAV_Navigation/path_planning_GA/path_planner_hrmaga.py
PF = UAV_Navigation/path_planning_GA/GA_complex _PF.py   (needs matplotlib stubs)
GAE = uav-backend/app/routers/GA_enhanced.py (class definitions only: the
      module runs a whole GA and plt.show() at import, GAE:181-194)
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("TRACEFLY_REFERENCE", "/root/reference")
HR_PATH = os.path.join(REF_ROOT, "UAV_Navigation/path_planning_GA/path_planner_hrmaga.py")
PF_PATH = os.path.join(REF_ROOT, "UAV_Navigation/path_planning_GA/GA_complex _PF.py")
GAE_PATH = os.path.join(REF_ROOT, "uav-backend/app/routers/GA_enhanced.py")


def available() -> bool:
    return os.path.exists(HR_PATH) and os.path.exists(PF_PATH)


def _stub_matplotlib():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_tf_stub", False):
        return
    names = ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
             "mpl_toolkits.mplot3d.art3d"]
    for n in names:
        if n not in sys.modules:
            m = types.ModuleType(n)
            m._tf_stub = True
            sys.modules[n] = m
    sys.modules["mpl_toolkits.mplot3d.art3d"].Poly3DCollection = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def _load(name, path):
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_hr():
    return _load("_tracefly_ref_hr", HR_PATH)


def load_pf():
    _stub_matplotlib()
    return _load("_tracefly_ref_pf", PF_PATH)


def load_gae_classes():
    """exec only the class definitions of GA_enhanced.py (source lines 1-179)."""
    _stub_matplotlib()
    with open(GAE_PATH, "r") as f:
        lines = f.readlines()
    cut = next(i for i, l in enumerate(lines) if l.startswith("# Sample Input"))
    ns = {"__name__": "_tracefly_ref_gae"}
    exec(compile("".join(lines[:cut]), GAE_PATH, "exec"), ns)
    return types.SimpleNamespace(**{k: v for k, v in ns.items() if not k.startswith("__")})
