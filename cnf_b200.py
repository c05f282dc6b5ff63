"""Import shim: the package directory is named ``continuousnormalizingflows.jl_b200``
(with a dot, after the reference repo), which Python's import statement cannot
spell.  ``import cnf_b200`` loads that directory as the module ``cnf_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "continuousnormalizingflows.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "cnf_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cnf_b200"] = _mod
_spec.loader.exec_module(_mod)
