"""cornerstone-b200: B200-native (sm_100a) implementation of the Cornerstone domain-sync hot path.

The product is libcstone_b200.so (hand-written CUDA behind the C ABI of include/cstone_b200.h); this package is the
thin host-side mirror used by the tests and the benchmark.  Nothing here falls back to a CPU path."""
from . import capi  # noqa: F401
from .capi import CstoneError, kernel_launch_count  # noqa: F401
