"""B200-native dense pyramidal Lucas-Kanade path behind the reference's `namespace gpu` interface.

The package holds only what that path needs: csrc/ (hand-written sm_100a kernels + the C ABI of
include/ofb200.h), api.py (host-side mirror of the reference's entry points) and dist.py (frame-batch
sharding and row-strip partitioning over torch.distributed).  There is no CPU implementation here.
"""
from ._lib import (LIB_PATH, MAX_LEVELS, MAX_WINDOW, SOLVE_EXACT, SOLVE_FAST, WARP_AS_WRITTEN, WARP_BILINEAR, WARP_NEAREST,
                   OfbError, OfbParams)
from .api import REFERENCE_WINDOW, Context, align_up, planar_to_device, write_flo

__all__ = ["Context", "OfbError", "OfbParams", "WARP_AS_WRITTEN", "WARP_NEAREST", "WARP_BILINEAR", "SOLVE_EXACT", "SOLVE_FAST", "MAX_LEVELS",
           "MAX_WINDOW", "REFERENCE_WINDOW", "LIB_PATH", "align_up", "planar_to_device", "write_flo"]
