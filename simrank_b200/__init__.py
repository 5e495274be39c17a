"""B200-native SimRank iteration engine.

Layout: ``csrc/`` hand-written sm_100a kernels + the C ABI (include/simrank_b200.h),
``_lib`` ctypes binding, ``graph`` host graph construction, ``engine`` device orchestration,
``drivers`` glue for the drop-in classes in the top-level ``SimRank`` package, ``dist``
row-sharded multi-GPU solver, ``synth`` seeded synthetic graphs of the BASELINE configs.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "graph", "engine", "drivers", "dist", "synth", "build"]
