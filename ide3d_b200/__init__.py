"""Importable alias of the package directory `ide-3d_b200/` (a hyphen cannot appear in a Python module name).
All code lives in ../ide-3d_b200; this shim points the package search path there and runs its __init__."""

import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'ide-3d_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _os, _f
