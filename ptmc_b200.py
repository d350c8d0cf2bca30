"""Import shim: the package directory is named after the reference repo
(`paralleltemperingmontecarlo.jl_b200/`, which is not a valid Python identifier), so this module loads
it under the importable name `ptmc_b200`."""
import importlib.util
import pathlib
import sys

_dir = pathlib.Path(__file__).resolve().parent / "paralleltemperingmontecarlo.jl_b200"
_spec = importlib.util.spec_from_file_location(
    "ptmc_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ptmc_b200"] = _mod
_spec.loader.exec_module(_mod)
