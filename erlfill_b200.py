"""Import alias: the package directory is `erl-fill_b200/` (not a Python identifier), so
`import erlfill_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "erl-fill_b200")
_spec = importlib.util.spec_from_file_location("erlfill_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["erlfill_b200"] = _mod
_spec.loader.exec_module(_mod)
