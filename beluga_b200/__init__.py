"""Import shim: the product package lives in ``beluga.jl_b200/`` (a directory name Python's
import statement cannot spell).  ``import beluga_b200`` loads that directory as this package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "beluga.jl_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _fh
