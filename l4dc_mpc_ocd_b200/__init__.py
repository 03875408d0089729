"""Import alias: the product package lives in ``l4dc-mpc-ocd_b200/`` (a directory name Python
cannot import directly).  ``import l4dc_mpc_ocd_b200`` executes that package's ``__init__`` with
``__path__`` pointing at it, so ``l4dc_mpc_ocd_b200.engine`` etc. resolve to the real files."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "l4dc-mpc-ocd_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
