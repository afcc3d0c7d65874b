"""Import shim: the package directory is ``hc-mvs_b200/`` (not a valid Python identifier),
so ``import hcmvs_b200`` resolves here and re-exports it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hc-mvs_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
