"""Import shim: the package directory is named `zpaq-v_b200` (not a Python identifier); this module
registers it under the importable name `zpaq_v_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zpaq-v_b200")
_spec = importlib.util.spec_from_file_location("zpaq_v_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["zpaq_v_b200"] = _mod
_spec.loader.exec_module(_mod)
