"""Import alias: ``import spef_b200`` -> the package in ``spacecraft-pose-estimation-framework_b200/``.

The build contract names the package directory after the reference repository; a hyphenated name
is not a Python identifier, so this stub extends its ``__path__`` to that directory and runs its
``__init__``.  All sub-modules therefore live under a single name (``spef_b200.*``).
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "spacecraft-pose-estimation-framework_b200")
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError(f"spef_b200: package directory not found: {_real}")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
