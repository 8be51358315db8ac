"""Import alias: the package directory is `video-caption-algorithm_b200/` (not a
valid Python identifier), so `import vcb200` loads it under this name."""
import importlib.util as _u
import pathlib as _p
import sys as _s

_dir = _p.Path(__file__).resolve().parent / "video-caption-algorithm_b200"
_spec = _u.spec_from_file_location("vcb200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_s.modules["vcb200"] = _mod
_spec.loader.exec_module(_mod)
