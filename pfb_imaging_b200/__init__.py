"""Import shim: the package sources live in ``pfb-imaging_b200/`` (the name the
repo layout mandates, which is not a valid Python identifier).  Importing
``pfb_imaging_b200`` executes that directory's ``__init__.py`` with this
package's ``__path__`` pointing at it, so ``pfb_imaging_b200.plan`` etc. resolve
to ``pfb-imaging_b200/plan.py``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pfb-imaging_b200")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
