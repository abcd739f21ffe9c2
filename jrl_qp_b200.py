"""Import shim: the package lives in ./jrl-qp_b200/ (a hyphen is not importable)."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jrl-qp_b200")
_spec = importlib.util.spec_from_file_location("jrl_qp_b200", os.path.join(_d, "__init__.py"),
                                               submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["jrl_qp_b200"] = _mod
_spec.loader.exec_module(_mod)
