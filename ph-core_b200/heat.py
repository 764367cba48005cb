"""Heat-equation stencil entry points (examples/heat_equation.cr:26-51) on device arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import K, check
from .narray import DeviceNArray, dtype_code

FIXED = K["PH_HEAT_FIXED"]
EXAMPLE1D = K["PH_HEAT_EXAMPLE1D"]


def _ext(shape):
    return (C.c_int64 * len(shape))(*[int(s) for s in shape])


def update_temp(state: DeviceNArray, coeff, mode: int = FIXED) -> DeviceNArray:
    """One explicit step (update_temp, examples/heat_equation.cr:38-51): returns a new array."""
    if state.dtype.kind != "f":
        raise TypeError("the heat stencil is defined for Float32 / Float64")
    out = DeviceNArray(state.shape, state.dtype)
    c = np.array(coeff, dtype=state.dtype)
    check(_lib.load().ph_heat_step(dtype_code(state.dtype), len(state.shape), _ext(state.shape), c.ctypes.data,
                                   mode, state.ptr, out.ptr))
    return out


def simulate(state: DeviceNArray, coeff, steps: int, mode: int = FIXED) -> DeviceNArray:
    """`steps` steps (simulate, examples/heat_equation.cr:26-36), ping-ponging two buffers;
    the input array is consumed as one of them."""
    if state.dtype.kind != "f":
        raise TypeError("the heat stencil is defined for Float32 / Float64")
    other = DeviceNArray(state.shape, state.dtype)
    c = np.array(coeff, dtype=state.dtype)
    final_is_b = C.c_int32(0)
    check(_lib.load().ph_heat_run(dtype_code(state.dtype), len(state.shape), _ext(state.shape), c.ctypes.data,
                                  mode, state.ptr, other.ptr, int(steps), C.byref(final_is_b)))
    return other if final_is_b.value else state


def simulate_into(state: DeviceNArray, other: DeviceNArray, coeff, steps: int, mode: int = FIXED) -> DeviceNArray:
    """`simulate` between two caller-owned buffers (no allocation: a 2048^3 f32 grid is 34 GB per buffer);
    returns whichever of the two holds the final state."""
    if state.dtype.kind != "f" or other.dtype != state.dtype or list(other.shape) != list(state.shape):
        raise TypeError("simulate_into needs two Float32 / Float64 arrays of one shape")
    c = np.array(coeff, dtype=state.dtype)
    final_is_b = C.c_int32(0)
    check(_lib.load().ph_heat_run(dtype_code(state.dtype), len(state.shape), _ext(state.shape), c.ctypes.data,
                                  mode, state.ptr, other.ptr, int(steps), C.byref(final_is_b)))
    return other if final_is_b.value else state
