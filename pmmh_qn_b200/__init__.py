"""Importable alias of the ``pmmh-qn_b200/`` package directory (a hyphen is not a valid
Python identifier).  ``import pmmh_qn_b200`` executes ``pmmh-qn_b200/__init__.py`` in this
module's namespace and points ``__path__`` there, so sub-modules resolve normally."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pmmh-qn_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
del _fh
