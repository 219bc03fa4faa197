"""Importable alias of the package directory `self-supervised-learning-for-aerial-image-segmentation_b200/`
(whose name, fixed by the project layout, is not a valid Python identifier).  Sub-modules are resolved
from that directory through the extended package __path__."""
import os as _os

_IMPL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "self-supervised-learning-for-aerial-image-segmentation_b200")
__path__.insert(0, _IMPL)
with open(_os.path.join(_IMPL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_IMPL, "__init__.py"), "exec"))
del _os, _f
