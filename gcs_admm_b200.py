"""Loader: makes the hyphenated package directory ``gcs-admm_b200/`` importable
as ``gcs_admm_b200`` (``import gcs_admm_b200.graph`` etc.)."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_root, "gcs-admm_b200")
_spec = importlib.util.spec_from_file_location(
    "gcs_admm_b200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gcs_admm_b200"] = _mod
_spec.loader.exec_module(_mod)
