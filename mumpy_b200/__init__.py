"""Import shim: `mumpy_b200` is the importable name of the product package, whose directory
(`multilateral-temporal-view-pyramid-transformer-for-video-inpainting-detection_b200/`) is not a valid identifier."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "multilateral-temporal-view-pyramid-transformer-for-video-inpainting-detection_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
