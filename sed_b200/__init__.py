"""Importable alias of the product package, which lives in `sound-event-detection_b200/` (a name
Python cannot import directly).  `import sed_b200.models` resolves into that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "sound-event-detection_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
