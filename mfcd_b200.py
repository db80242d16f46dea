"""Import shim: makes the package directory
``matrix-factorization-with-comparison-data_b200/`` importable as ``mfcd_b200``
(its on-disk name contains dashes)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "matrix-factorization-with-comparison-data_b200")
_spec = importlib.util.spec_from_file_location(
    "mfcd_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mfcd_b200"] = _mod
_spec.loader.exec_module(_mod)
