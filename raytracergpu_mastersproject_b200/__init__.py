"""raytracergpu_mastersproject_b200 -- B200 (sm_100a) compute path of the RaytracerGPU path tracer.

The product is librtb200.so (hand-written CUDA behind the C-ABI of include/rtb200.h); this package is the thin
Python host side used by tests and bench.py.  The C++20 host mirror of the reference API lives in host/.
Importing the package does not load CUDA; the first use of capi.lib() does and fails loudly if the library is
missing -- there is no CPU fallback.
"""
from . import capi  # noqa: F401
from .capi import RtbError  # noqa: F401
from .renderer import Buffer, Device, Raytracer, make_ubo  # noqa: F401

__all__ = ["capi", "RtbError", "Buffer", "Device", "Raytracer", "make_ubo"]
