"""Import shim: loads the package that lives in `tracefly-ga-uav-pathplanning_b200/`
(a directory name Python cannot import directly) under the name `tracefly_b200`."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "tracefly-ga-uav-pathplanning_b200")
_spec = _ilu.spec_from_file_location("tracefly_b200", _os.path.join(_real, "__init__.py"),
                                     submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["tracefly_b200"] = _mod
_spec.loader.exec_module(_mod)
