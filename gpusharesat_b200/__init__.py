"""gpusharesat_b200 -- B200-native (sm_100a) rewrite of GpuShareSat's clause-vs-assignment check.

The product is the C-ABI shared library ``libgpushare_b200.so`` (hand-written CUDA + C++ host,
``include/gpushare_b200.h``).  This package only holds the build recipe and a thin ctypes
mirror of the reference's ``GpuClauseSharer`` interface for tests and ``bench.py``.  There is
no CPU fallback: importing :mod:`gpusharesat_b200.api` fails loudly when the library is missing.
"""
from .api import (GpuClauseSharer, GpuClauseSharerOptions, GlobalStats, OneSolverStats,  # noqa: F401
                  load_library, library_path, mkLit)
