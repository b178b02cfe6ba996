"""Import shim: the package directory is named `vr180-convert_b200/` (not a Python identifier), so this module
loads it under the importable name `vr180_convert_b200` and replaces itself in sys.modules."""
import importlib.util as _u
import sys as _s
from pathlib import Path as _P

_dir = _P(__file__).resolve().parent / "vr180-convert_b200"
_spec = _u.spec_from_file_location("vr180_convert_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_s.modules["vr180_convert_b200"] = _mod
_spec.loader.exec_module(_mod)
