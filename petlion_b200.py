"""Import shim: the package directory is literally named `petlion.jl_b200/` (not importable by that
name because of the dot), so `import petlion_b200` loads it under this module name."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "petlion.jl_b200")
_spec = importlib.util.spec_from_file_location("petlion_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["petlion_b200"] = _mod
_spec.loader.exec_module(_mod)
