"""DeviceNArray: host-side mirror of ph-core's NArray API for the device path.

Mirrors the reference names and error behaviour (src/n_array.cr, src/multi_indexable.cr,
src/multi_writable.cr) so the parity tests read like the reference's specs.  All data
work is done by libphgpu kernels through the C-ABI; this file only validates shapes,
builds descriptors and converts flag words into the reference's exception classes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import PhDesc, K, check


# ---- the reference's exception classes (src/exceptions/exceptions.cr + Crystal stdlib)
class ShapeError(Exception):
    pass


class DimensionError(ShapeError):
    pass


class CrIndexError(Exception):
    pass


class CrOverflowError(Exception):
    pass


class CrDivisionByZeroError(Exception):
    pass


class CrArgumentError(Exception):
    pass


class CrEmptyError(Exception):
    pass


class DeviceBlockError(Exception):
    """Arbitrary blocks (map/each_with/apply/process/build) cannot run on the device path:
    they raise instead of silently running on the CPU (BASELINE.json north_star)."""


_DTYPES = {np.dtype(np.float32): K["PH_F32"], np.dtype(np.float64): K["PH_F64"],
           np.dtype(np.int32): K["PH_I32"], np.dtype(np.int64): K["PH_I64"],
           np.dtype(np.uint8): K["PH_U8"], np.dtype(np.bool_): K["PH_U8"],
           np.dtype(np.int8): K["PH_I8"], np.dtype(np.int16): K["PH_I16"],
           np.dtype(np.uint16): K["PH_U16"], np.dtype(np.uint32): K["PH_U32"],
           np.dtype(np.uint64): K["PH_U64"]}


def dtype_code(dt) -> int:
    try:
        return _DTYPES[np.dtype(dt)]
    except KeyError:
        raise TypeError(f"dtype {dt} has no device representation (primitive numerics and Bool only)")


def raise_for_flags(flags: int) -> None:
    """ph_take_arith_flags -> the exception the reference would have raised."""
    if flags & K["PH_FLAG_DIV0"]:
        raise CrDivisionByZeroError("Division by 0")
    if flags & K["PH_FLAG_OVERFLOW"]:
        raise CrOverflowError("Arithmetic overflow")
    if flags & K["PH_FLAG_ARGUMENT"]:
        raise CrArgumentError("invalid integer argument (MIN // -1 or negative exponent)")
    if flags & K["PH_FLAG_NAN"]:
        raise CrArgumentError("Comparison of NaN failed")


class _Buffer:
    """Ref-counted owner of one device allocation (reshape aliases the buffer,
    src/n_array.cr:429-433; views keep their source alive, src/view.cr:7)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(_lib.load().ph_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                _lib.load().ph_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


_BIN = {"+": "PH_ADD", "-": "PH_SUB", "*": "PH_MUL", "/": "PH_DIV", "//": "PH_FLOORDIV", "%": "PH_MOD",
        "**": "PH_POW", "&+": "PH_WADD", "&-": "PH_WSUB", "&*": "PH_WMUL", "&**": "PH_WPOW",
        "&": "PH_AND", "|": "PH_OR", "^": "PH_XOR"}
_CMP = {">": "PH_GT", "<": "PH_LT", ">=": "PH_GE", "<=": "PH_LE", "==": "PH_EQ", "!=": "PH_NE"}


class DeviceNArray:
    """Row-major N-D array resident in HBM (mirror of Phase::NArray, src/n_array.cr:15)."""

    def __init__(self, shape: Sequence[int], dtype, buf: Optional[_Buffer] = None):
        self.shape = [int(s) for s in shape]
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 0
        self._buf = buf if buf is not None else _Buffer(max(1, self.size) * self.dtype.itemsize)

    # ---- construction / transfer (explicit, never implicit) ------------------
    @classmethod
    def from_host(cls, arr: np.ndarray) -> "DeviceNArray":
        """NArray#to_device: explicit host -> device transfer."""
        _lib.init()
        arr = np.ascontiguousarray(arr)
        out = cls(arr.shape, arr.dtype)
        if arr.size:
            check(_lib.load().ph_h2d(out.ptr, arr.ctypes.data, arr.nbytes))
            check(_lib.load().ph_sync())   # pageable source: keep it alive until copied
        return out

    @classmethod
    def fill(cls, shape, value, dtype) -> "DeviceNArray":
        """NArray.fill (src/n_array.cr:230-232)."""
        _lib.init()
        out = cls(shape, dtype)
        if out.size:
            v = np.array(value, dtype=out.dtype)
            check(_lib.load().ph_fill_region(out.dtype.itemsize, out.ptr, C.byref(out.desc()), v.ctypes.data))
        return out

    def to_host(self) -> np.ndarray:
        """NArray#to_host / to_narr: explicit device -> host transfer."""
        out = np.empty(self.shape, dtype=self.dtype)
        if self.size:
            check(_lib.load().ph_d2h(out.ctypes.data, self.ptr, out.nbytes))
        return out

    @property
    def ptr(self) -> int:
        return self._buf.ptr

    @property
    def buffer(self):
        """Buffered#buffer (src/buffered/buffered.cr:10) on a device array raises: no silent D2H."""
        raise DeviceBlockError("a device NArray has no host buffer; call to_host explicitly")

    def desc(self) -> PhDesc:
        return PhDesc.contiguous(self.shape)

    def clone(self) -> "DeviceNArray":
        """NArray#clone deep-copies (src/n_array.cr:372-374)."""
        out = DeviceNArray(self.shape, self.dtype)
        if self.size:
            check(_lib.load().ph_d2d(out.ptr, self.ptr, self.size * self.dtype.itemsize))
        return out

    # ---- blocks are out of scope on the device path ------------------------------
    def _no_blocks(self, *a, **k):
        raise DeviceBlockError("arbitrary blocks cannot run on the device path")

    map = map_with = each_with = map_with_coord = apply = process = build = each = _no_blocks

    # ---- elementwise (src/multi_indexable.cr:931-985) -------------------------------
    def _result_dtype(self, op: str):
        if op == "/" and self.dtype.kind in "iu":
            return np.dtype(np.float64)
        return self.dtype

    def _binary(self, op: str, other, reflected: bool = False) -> "DeviceNArray":
        lib = _lib.load()
        code = K[_BIN[op]]
        dt = dtype_code(self.dtype)
        if isinstance(other, DeviceNArray):
            if other.dtype != self.dtype:
                raise TypeError("device path: operands must share a dtype")
            a, b = (other, self) if reflected else (self, other)
            if a.shape != b.shape:                          # multi_indexable.cr:935-940
                raise ShapeError(f"The shape of this MultiIndexable ({a.shape}) does not match the shape of "
                                 f"the one provided ({b.shape}), so '{op}' cannot be applied element-wise.")
            out = DeviceNArray(a.shape, self._result_dtype(op))
            d = a.desc()
            check(lib.ph_ewise_binary(code, dt, a.ptr, C.byref(d), b.ptr, C.byref(d), out.ptr, C.byref(d)))
            return out
        # scalar: array.map &.op(other) (:947-951) / Number#op(narr) (patches/number.cr:6-15)
        out = DeviceNArray(self.shape, self._result_dtype(op))
        d = self.desc()
        if op == "**" and self.dtype.kind == "f" and isinstance(other, (int, np.integer)) and not reflected:
            s = np.array(other, dtype=np.int32)
            code = K["PH_POWI"]
        else:
            s = np.array(other, dtype=self.dtype)
        check(lib.ph_ewise_scalar(code, dt, self.ptr, C.byref(d), s.ctypes.data, int(reflected), out.ptr, C.byref(d)))
        return out

    def broadcast_op(self, op: str, other: "DeviceNArray") -> "DeviceNArray":
        """NEW (ShapeUtil.broadcast_shapes): equal rank, size-1 axes stretch."""
        if len(self.shape) != len(other.shape):
            raise ShapeError("broadcast requires equal rank")
        shape = []
        for x, y in zip(self.shape, other.shape):
            if x == y or y == 1:
                shape.append(x)
            elif x == 1:
                shape.append(y)
            else:
                raise ShapeError(f"shapes {self.shape} and {other.shape} cannot be broadcast")
        out = DeviceNArray(shape, self._result_dtype(op))
        da, db, do = self.bcast_desc(shape), other.bcast_desc(shape), out.desc()
        check(_lib.load().ph_ewise_binary(K[_BIN[op]], dtype_code(self.dtype), self.ptr, C.byref(da),
                                          other.ptr, C.byref(db), out.ptr, C.byref(do)))
        return out

    def bcast_desc(self, shape) -> PhDesc:
        d = self.desc()
        for i, (mine, want) in enumerate(zip(self.shape, shape)):
            d.extent[i] = want
            if mine == 1 and want != 1:
                d.stride[i] = 0
        return d

    def mul_add(self, b: "DeviceNArray", c: "DeviceNArray") -> "DeviceNArray":
        """Fused (self * b) + c with two roundings; b may broadcast."""
        out = DeviceNArray(self.shape, self.dtype)
        da, db, dc, do = self.desc(), b.bcast_desc(self.shape), c.bcast_desc(self.shape), out.desc()
        check(_lib.load().ph_ewise_mul_add(dtype_code(self.dtype), self.ptr, C.byref(da), b.ptr, C.byref(db),
                                           c.ptr, C.byref(dc), out.ptr, C.byref(do)))
        return out

    def __add__(self, o): return self._binary("+", o)
    def __sub__(self, o): return self._binary("-", o)
    def __mul__(self, o): return self._binary("*", o)
    def __truediv__(self, o): return self._binary("/", o)
    def __floordiv__(self, o): return self._binary("//", o)
    def __mod__(self, o): return self._binary("%", o)
    def __pow__(self, o): return self._binary("**", o)
    def __and__(self, o): return self._binary("&", o)
    def __or__(self, o): return self._binary("|", o)
    def __xor__(self, o): return self._binary("^", o)
    def __radd__(self, o): return self._binary("+", o, True)
    def __rsub__(self, o): return self._binary("-", o, True)
    def __rmul__(self, o): return self._binary("*", o, True)
    def __rtruediv__(self, o): return self._binary("/", o, True)
    def __rfloordiv__(self, o): return self._binary("//", o, True)
    def __rmod__(self, o): return self._binary("%", o, True)
    def __rpow__(self, o): return self._binary("**", o, True)
    def wrapping_add(self, o): return self._binary("&+", o)       # Crystal &+
    def wrapping_sub(self, o): return self._binary("&-", o)       # Crystal &-
    def wrapping_mul(self, o): return self._binary("&*", o)       # Crystal &*
    def wrapping_pow(self, o): return self._binary("&**", o)      # Crystal &**

    def _unary(self, name: str) -> "DeviceNArray":
        out = DeviceNArray(self.shape, self.dtype)
        d = self.desc()
        check(_lib.load().ph_ewise_unary(K[name], dtype_code(self.dtype), self.ptr, C.byref(d), out.ptr, C.byref(d)))
        return out

    def __pos__(self): return self._unary("PH_POS")
    def __neg__(self): return self._unary("PH_NEG")
    def __invert__(self): return self._unary("PH_NOT")

    def _compare(self, op: str, other, reflected=False, eq_style=False) -> "DeviceNArray":
        lib = _lib.load()
        out = DeviceNArray(self.shape, np.bool_)
        d = self.desc()
        dt = dtype_code(self.dtype)
        if isinstance(other, DeviceNArray):
            if other.shape != self.shape:
                if eq_style:                                    # multi_indexable.cr:900-902
                    raise DimensionError("Cannot compute the element-wise equality: shapes differ")
                raise ShapeError("shapes differ")               # :935-940
            check(lib.ph_compare(K[_CMP[op]], dt, self.ptr, C.byref(d), other.ptr, C.byref(d), out.ptr, C.byref(d)))
        else:
            s = np.array(other, dtype=self.dtype)
            check(lib.ph_compare_scalar(K[_CMP[op]], dt, self.ptr, C.byref(d), s.ctypes.data, int(reflected),
                                        out.ptr, C.byref(d)))
        return out

    def __gt__(self, o): return self._compare(">", o)
    def __lt__(self, o): return self._compare("<", o)
    def __ge__(self, o): return self._compare(">=", o)
    def __le__(self, o): return self._compare("<=", o)
    def eq(self, o): return self._compare("==", o, eq_style=True)   # MultiIndexable#eq (:899-913)

    # ---- data-dependent errors -----------------------------------------------------
    @staticmethod
    def take_flags() -> int:
        f = C.c_uint32()
        check(_lib.load().ph_take_arith_flags(C.byref(f)))
        return f.value

    @staticmethod
    def raise_pending() -> None:
        """Synchronise and raise OverflowError / DivisionByZeroError / ArgumentError if any
        launched op hit one (SURVEY.md 8(b) error conventions)."""
        raise_for_flags(DeviceNArray.take_flags())
