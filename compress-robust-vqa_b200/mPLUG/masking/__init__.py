"""Drop-in ``masking`` package of the reference's ``mPLUG/masking/``.

``sparsity_control.py`` is byte-identical in the reference's ``masking/`` and ``mPLUG/masking/``; here it (and the
shared masked-module core, ``_core.py``) resolve to the one implementation in ``<package root>/masking`` through
this package's ``__path__``, whether the package is imported as ``mPLUG.masking`` or, with ``mPLUG/`` first on
``sys.path``, as ``masking``.
"""
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_ROOT not in sys.path:          # `crvqa` (ctypes binding of libcrvqa.so) lives there
    sys.path.append(_PKG_ROOT)
__path__.append(os.path.join(_PKG_ROOT, "masking"))
