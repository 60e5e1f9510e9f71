"""rayfinder_b200 — B200-native (sm_100a) drop-in for the render path of Nelarius/rayfinder.

The product is ``librayfinder_b200.so`` (C-ABI in ``include/rayfinder_b200.h``: CUDA wavefront path
tracer + BVH traversal kernels, and the host pieces the reference keeps on the CPU).  This package is
the thin Python host side used by the tests and ``bench.py``; it contains no compute of its own and no
CPU fallback.
"""
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all
from . import capi  # noqa: F401

__all__ = list(_api_all) + ["capi"]
