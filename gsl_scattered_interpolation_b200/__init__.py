"""gsl_scattered_interpolation_b200 -- B200-native batched scattered-data interpolation.

The product is the C-ABI library ``libgsl_scattered_cuda.so`` (``include/gsl_scattered_cuda.h``,
sources in ``csrc/``, built by ``make`` there); this Python package only binds it with ctypes
(``_lib.py``, ``api.py``) for the tests and the benchmark driver, and holds the deterministic
synthetic workload generators (``synth.py``).
"""
from .api import *  # noqa: F401,F403
from . import synth  # noqa: F401
