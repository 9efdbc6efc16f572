"""Import alias: ``import pamnet_b200`` loads the package that lives in the (non-identifier)
directory ``physics-aware-multiplex-gnn_b200/`` next to this file."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "physics-aware-multiplex-gnn_b200")
_spec = importlib.util.spec_from_file_location(
    "pamnet_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pamnet_b200"] = _mod
_spec.loader.exec_module(_mod)
