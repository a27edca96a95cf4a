"""superman_b200 -- B200-native drop-in for SUPerman's GPU permanent paths.

Python is only the test/bench driver: the product is libsuperman_b200.so (C host + sm_100a CUDA,
C-ABI in include/superman_b200.h) and the `perman` CLI.  `superman_b200.api` mirrors the
reference's host-wrapper names (gpu_perman64_*) on top of that C-ABI.
"""
from . import _ffi  # noqa: F401  (fails loudly when the shared object is missing)
from .api import *  # noqa: F401,F403
