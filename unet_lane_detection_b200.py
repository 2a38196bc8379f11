"""Import shim: the package directory is named `unet-lane-detection_b200/` (with hyphens, as the
project layout prescribes), which Python cannot import by name. This module loads it under the
importable name `unet_lane_detection_b200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unet-lane-detection_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
