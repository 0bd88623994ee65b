"""Import alias: ``import fav`` loads the package that lives in ``failure-aware-vision_b200/``
(a directory name that is not a Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "failure-aware-vision_b200")
_spec = importlib.util.spec_from_file_location("fav", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fav"] = _mod
_spec.loader.exec_module(_mod)
