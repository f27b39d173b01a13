"""Importable alias of the package directory ``parallel-tempering-neural-net_b200/`` (whose name,
fixed by the project layout, is not a valid Python identifier).  ``import ptnn_b200`` resolves the
sub-modules (capi, sampler, regression, classification, ...) from that directory."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "parallel-tempering-neural-net_b200")
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
