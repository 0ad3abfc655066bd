"""Import shim: the package directory is named `diffeqoperators.jl_b200` (with a dot), which Python
cannot import by name.  `import deo_b200` loads it from that directory under this module name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffeqoperators.jl_b200")
_spec = importlib.util.spec_from_file_location("deo_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["deo_b200"] = _mod
_spec.loader.exec_module(_mod)
