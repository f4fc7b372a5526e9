"""Importable alias for the framework package.

The product package lives in the directory
``pytorch-distributeddataparallel-training-tricks_b200/`` (the name the build
spec mandates).  A hyphenated directory cannot be imported with a plain
``import`` statement, so this tiny alias package points its ``__path__`` at that
directory and executes its ``__init__``.  ``import ddpt_b200`` and
``import ddpt_b200.parallel`` therefore resolve to the real sources; there is
no second copy of any code.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(_ROOT, "pytorch-distributeddataparallel-training-tricks_b200")
__path__ = [PACKAGE_DIR]  # submodules are loaded from the real package directory

with open(_os.path.join(PACKAGE_DIR, "__init__.py"), "r") as _f:
    exec(compile(_f.read(), _os.path.join(PACKAGE_DIR, "__init__.py"), "exec"))
del _f
