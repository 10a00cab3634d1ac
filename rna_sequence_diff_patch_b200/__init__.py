"""Importable alias of the package directory `rna-sequence-diff-patch_b200/` (a hyphen cannot appear
in a Python module name).  All code lives there; this shim only points the import system at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "rna-sequence-diff-patch_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
