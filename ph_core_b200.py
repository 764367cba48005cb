"""Import shim: the package directory is named `ph-core_b200/` (not a valid Python
identifier), so `import ph_core_b200` loads it from there under this name."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ph-core_b200")
_spec = importlib.util.spec_from_file_location(
    "ph_core_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ph_core_b200"] = _mod
_spec.loader.exec_module(_mod)
