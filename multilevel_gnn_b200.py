"""Import shim: the package directory is ``multilevel-gnn_b200/`` (hyphen fixed by the project
layout), which is not a valid Python identifier; this module loads it under the importable name
``multilevel_gnn_b200`` (sub-modules resolve normally: ``multilevel_gnn_b200.functional`` ...)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multilevel-gnn_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
