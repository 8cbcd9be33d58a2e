"""Import shim: the package directory is named `slam.jl_b200` (not a valid Python identifier), so it is
loaded by path and exposed as the module `slamklt`.  `import slamklt; slamklt.Context(0)`."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_root, "slam.jl_b200")
_spec = importlib.util.spec_from_file_location("slam_jl_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["slam_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
_sspec = importlib.util.spec_from_file_location("slam_jl_b200.synth", os.path.join(_pkg, "synth.py"))
synth = importlib.util.module_from_spec(_sspec)
sys.modules["slam_jl_b200.synth"] = synth
_sspec.loader.exec_module(synth)
_dspec = importlib.util.spec_from_file_location("slam_jl_b200.dist", os.path.join(_pkg, "dist.py"))
dist = importlib.util.module_from_spec(_dspec)
sys.modules["slam_jl_b200.dist"] = dist
_dspec.loader.exec_module(dist)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
