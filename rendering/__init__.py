"""Drop-in alias: `import rendering as ren` (what every reference tutorial does) resolves to
rendertoy_b200.rendering, including submodule imports such as `rendering._raster` / `rendering._core`."""
import sys as _sys

from rendertoy_b200 import rendering as _impl
from rendertoy_b200.rendering import *  # noqa: F401,F403
from rendertoy_b200.rendering import _core, _modeling, _loaders, _presentation, _raster, _raycaster, _dsl  # noqa: F401

for _name in ("_core", "_modeling", "_loaders", "_presentation", "_raster", "_raycaster", "_dsl"):
    _mod = _sys.modules.get(f"rendertoy_b200.rendering.{_name}")
    if _mod is not None:
        _sys.modules[f"{__name__}.{_name}"] = _mod
__all__ = [n for n in dir(_impl) if not n.startswith("__")]
