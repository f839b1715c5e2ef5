"""Import shim: exposes the package in ``homotopycontinuation.jl_b200/`` (not importable by name
because of the dot) as ``hcb200``."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "homotopycontinuation.jl_b200")
_spec = _u.spec_from_file_location("hcb200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["hcb200"] = _mod
_spec.loader.exec_module(_mod)
