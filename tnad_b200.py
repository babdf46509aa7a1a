"""Import alias: `import tnad_b200` loads the package directory `tensornetworkad.jl_b200/`.

The directory name is fixed by the project layout and contains a dot, which Python's import
statement cannot spell; this module registers that directory under the importable name `tnad_b200`.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "tensornetworkad.jl_b200")
_spec = _ilu.spec_from_file_location("tnad_b200", _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["tnad_b200"] = _mod
_spec.loader.exec_module(_mod)
