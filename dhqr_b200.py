"""Import shim: ``import dhqr_b200`` loads the package in ./distributedhouseholderqr.jl_b200/
(that directory name is fixed by the task and is not an importable identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "distributedhouseholderqr.jl_b200")
_spec = importlib.util.spec_from_file_location("dhqr_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dhqr_b200"] = _mod
_spec.loader.exec_module(_mod)
