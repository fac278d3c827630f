"""flatdd_b200 — B200-native (sm_100a) implementation of FlatDD's array-phase hot path.

The product is the C-ABI shared library ``libflatdd_b200.so`` (CUDA kernels, built from
``flatdd_b200/csrc``; contract in ``include/flatdd_b200.h``) and the C++ host driver under
``flatdd_b200/host``.  This Python package is only the thin ctypes binding the tests and
``bench.py`` use; it contains no compute and no CPU fallback.
"""
from .flat import FlatDD, read_flat, read_trace, TraceRecord  # noqa: F401
from .capi import Library, Context, FlatDDError, load_library, library_path  # noqa: F401
