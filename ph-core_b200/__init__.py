"""ph-core_b200: device-resident NArray backing for ph-core's data-parallel hot path
on NVIDIA B200 (sm_100a).  The product is csrc/ (hand-written CUDA behind the C-ABI of
include/ph_gpu.h); this Python package is the thin host mirror used by tests and bench.
There is no CPU fallback: importing works anywhere, computing needs the GPU library."""
from . import _lib
from ._lib import PhDesc, PhError, init, load, check, K
from .narray import (DeviceNArray, DeviceView, make_region, cover_region, ShapeError, DimensionError, CrIndexError, CrOverflowError,
                     CrDivisionByZeroError, CrArgumentError, CrEmptyError, DeviceBlockError)

from .region import R, Step, rng, ALL
from .narray import Stream, pinned_empty, pinned_from, sync
from . import heat
from . import io
from . import sharding
from . import pipeline

__all__ = ["Stream", "pinned_empty", "pinned_from", "sync", "pipeline", "DeviceNArray", "DeviceView", "make_region", "cover_region", "R", "Step", "rng", "ALL", "heat", "io", "sharding", "PhDesc", "PhError", "init", "load", "check", "K", "ShapeError", "DimensionError",
           "CrIndexError", "CrOverflowError", "CrDivisionByZeroError", "CrArgumentError", "CrEmptyError",
           "DeviceBlockError"]
