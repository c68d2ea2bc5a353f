"""back2future_b200 -- B200-native (sm_100a) implementation of the Back2Future hot path.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of include/b2f.h, built into
``libb2f_cuda.so``), ``_lib`` (ctypes binding) and ``nn`` (host-side mirror of the reference's
Torch7 module / criterion surface).  Importing the package does not load the shared library;
the first call does, and raises if it is missing -- there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import nn  # noqa: F401

__version__ = "0.1.0"
