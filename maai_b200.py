"""Import shim: exposes the package in ``multimodal-active-ai_b200/`` (not a valid Python
identifier) as the module ``maai_b200``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimodal-active-ai_b200")
_spec = importlib.util.spec_from_file_location(
    "maai_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["maai_b200"] = _mod
_spec.loader.exec_module(_mod)
