"""Import shim: the package directory `hssmatrices.jl_b200/` has a dot in its
name, so it cannot be imported with a plain `import` statement.  `import hssb200`
loads it from that directory and re-exports it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hssmatrices.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "hssmatrices_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hssmatrices_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
sys.modules[__name__] = _mod
