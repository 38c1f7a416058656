"""Import shim: the package directory is named `dlpoly2-mbpol-heat_b200` (not a valid Python identifier),
so `import mbpol_b200` loads it from there under this name."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_root, "dlpoly2-mbpol-heat_b200")
_spec = importlib.util.spec_from_file_location("mbpol_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mbpol_b200"] = _mod
_spec.loader.exec_module(_mod)
