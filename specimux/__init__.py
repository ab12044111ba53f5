"""`specimux` -- drop-in alias of the B200 package under the reference's own import name.

`import specimux`, `python -m specimux.cli`, `specimux.core.TrimMode` ... resolve to specimux_b200, so the
reference's scripts and its own test-suite (tests/test_integration.py, tests/test_orientation_normalization.py)
run unchanged against the GPU path (reference: src/specimux/__init__.py:6-25).
"""
import importlib as _importlib
import sys as _sys

import specimux_b200 as _impl
from specimux_b200 import *                     # noqa: F401,F403
from specimux_b200 import __all__, __version__  # noqa: F401

# submodules the reference exposes and this package mirrors; `specimux.cli` and `specimux.core` are real files so
# that `python -m specimux.cli` works
for _name in ("constants", "databases", "models", "demultiplex", "io_utils", "orchestration", "trace", "specimine"):
    try:
        _sys.modules[__name__ + "." + _name] = _importlib.import_module("specimux_b200." + _name)
    except ImportError:                         # optional module not built in this tree
        pass
del _importlib, _sys, _name, _impl
