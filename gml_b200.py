"""Import alias: the package directory is named `graphicalmodellearning.jl_b200/` (a dot cannot be
imported directly), so `import gml_b200` loads it under this name."""
import importlib.util
import pathlib
import sys

_pkg_dir = pathlib.Path(__file__).resolve().parent / "graphicalmodellearning.jl_b200"
_spec = importlib.util.spec_from_file_location("gml_b200", _pkg_dir / "__init__.py",
                                               submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gml_b200"] = _mod
_spec.loader.exec_module(_mod)
