"""Importable alias of the package directory `open-o3-video_b200/` (a hyphenated name
cannot be imported).  `import open_o3_video_b200 as o3v`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "open-o3-video_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
