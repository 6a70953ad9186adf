"""Import alias: the package directory is `sensorfusion-kalmanfilter_b200/` (the repo's name,
which is not a valid Python identifier); this module makes it importable as
`sensorfusion_kalmanfilter_b200` by pointing __path__ at that directory and running its
__init__.py in this module's namespace."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sensorfusion-kalmanfilter_b200")
__path__ = [_PKG_DIR]
__file__ = _os.path.join(_PKG_DIR, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
