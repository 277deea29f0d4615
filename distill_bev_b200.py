"""Import alias: ``import distill_bev_b200`` -> the package in ``distill-bev_b200/``.

The package directory carries the repository's name (with a hyphen, which
Python cannot import directly); this module loads it under an importable name
and replaces itself in ``sys.modules``.
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "distill-bev_b200")
_spec = importlib.util.spec_from_file_location(
    "distill_bev_b200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["distill_bev_b200"] = _mod
_spec.loader.exec_module(_mod)
