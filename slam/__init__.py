"""Import shim: the package directory is named ``slam.net_b200`` (with a dot), which the normal
import machinery reads as ``slam`` -> ``net_b200``.  This module makes ``import slam.net_b200`` load
the directory ``<repo>/slam.net_b200``."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PKG_DIR = os.path.join(_ROOT, "slam.net_b200")

if "slam.net_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        "slam.net_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["slam.net_b200"] = _mod
    _spec.loader.exec_module(_mod)
    net_b200 = _mod
else:
    net_b200 = sys.modules["slam.net_b200"]
