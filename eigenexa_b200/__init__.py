"""eigenexa_b200 -- B200-native (sm_100a) implementation of EigenExa's eigen_s hot path.

The product is the CUDA shared library built from ``csrc/`` (C ABI in
``include/eigenexa_b200.h``); this package is the thin host-side mirror of the reference's
public interface.  See DESIGN.md / INTEGRATION.md.
"""
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
