"""amuse_b200: a B200-native GRAPE-6/Sapporo direct-summation force library for AMUSE.

The product is the C-ABI shared library ``amuse_b200/csrc/libsapporo.so`` (alias ``libg6.so``),
see ``include/g6_b200.h``; ``g6lib`` is a thin ctypes binding used by tests and the benchmark,
``plummer`` generates synthetic inputs.
"""
__version__ = "0.1.0"
