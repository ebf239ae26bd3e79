"""Import alias: ``import morefusion`` resolves to the B200 implementation (morefusion_b200), so
reference-side code written against ``morefusion.functions.*`` / ``morefusion.contrib.*`` /
``morefusion.metrics.*`` / ``morefusion.geometry.*`` finds the same names.

Only the hot-path members exist (see DESIGN.md section 7 for what is out of scope); anything else
raises AttributeError / ImportError instead of silently doing something different."""

import importlib
import sys

import morefusion_b200 as _impl
from morefusion_b200 import InvalidType, config  # noqa: F401

__version__ = _impl.__version__

for _name in ("functions", "functions.geometry", "functions.loss", "contrib",
              "contrib.singleview_3d", "contrib.singleview_3d.models", "geometry", "metrics",
              "synthetic"):
    _mod = importlib.import_module("morefusion_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod
