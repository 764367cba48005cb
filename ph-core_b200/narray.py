"""DeviceNArray / DeviceView: host-side mirror of ph-core's NArray / View API for the device path.

Mirrors the reference names and error behaviour (src/n_array.cr, src/multi_indexable.cr,
src/multi_writable.cr, src/view.cr, src/mutable_view.cr) so the parity tests read like the
reference's specs.  This file is plumbing: index math is done by the C++ host layer
(include/ph_host.h), all data work by libphgpu kernels (include/ph_gpu.h).  It converts status
codes / flag words into the reference's exception classes and never touches array data on
the host, except through the explicit to_host / from_host transfers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import PhDesc, PhRegion, PhRangeLit, K, check
from . import region as _region


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


def host_check(status: int) -> None:
    """ph_host.h status -> the reference's exception class."""
    if status == 0:
        return
    msg = _lib.load().ph_host_last_error()
    msg = msg.decode() if msg else ""
    exc = {K["PH_HOST_INDEX_ERROR"]: CrIndexError, K["PH_HOST_DIMENSION_ERROR"]: DimensionError,
           K["PH_HOST_SHAPE_ERROR"]: ShapeError, K["PH_HOST_DIV0_ERROR"]: CrDivisionByZeroError}.get(status)
    if exc is None:
        raise _lib.PhError(f"ph_host status {status}: {msg}")
    raise exc(msg)


def _i64(vals):
    vals = [int(v) for v in vals]
    return (C.c_int64 * max(1, len(vals)))(*vals)


def make_region(literal: Sequence, bound_shape: Sequence[int], drop: bool = True) -> PhRegion:
    """IndexRegion.new(region_literal, bound_shape, drop) (src/index_region.cr:192-224)."""
    lits = [_region.marshal(l) for l in literal]
    arr = (PhRangeLit * max(1, len(lits)))(*lits)
    out = PhRegion()
    host_check(_lib.load().ph_region_new(arr, len(lits), _i64(bound_shape), len(bound_shape), int(drop), C.byref(out)))
    return out


def cover_region(bound_shape: Sequence[int], drop: bool = True) -> PhRegion:
    """IndexRegion.cover(bound_shape) (src/index_region.cr:232-238): the whole array, also when an axis is
    empty (a literal `..` on a zero-length axis raises IndexError in the reference, range_syntax.cr:120-122)."""
    out = PhRegion()
    host_check(_lib.load().ph_region_cover(_i64(bound_shape), len(bound_shape), int(drop), C.byref(out)))
    return out


def _checked_d2h(dst_host, src_dev, nbytes: int) -> None:
    """ph_d2h_flags: the copy and the arithmetic flag word in one synchronisation; raises the
    reference's exception class when any earlier launch flagged one."""
    flags = C.c_uint32(0)
    check(_lib.load().ph_d2h_flags(dst_host, src_dev, int(nbytes), C.byref(flags)))
    raise_for_flags(flags.value)


def sync() -> None:
    """Device.sync: wait for every launched operator; a guaranteed raise point for data-dependent errors."""
    flags = C.c_uint32(0)
    check(_lib.load().ph_d2h_flags(None, None, 0, C.byref(flags)))
    raise_for_flags(flags.value)


class _PinnedOwner:
    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(_lib.load().ph_host_alloc(max(1, int(nbytes)), C.byref(p)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().ph_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """An uninitialised host array in pinned memory; slices of it are pinned too (they keep it alive)."""
    dt = np.dtype(dtype)
    n = int(np.prod([int(s) for s in shape], dtype=np.int64)) if len(shape) else 1
    owner = _PinnedOwner(n * dt.itemsize)
    buf = (C.c_char * max(1, n * dt.itemsize)).from_address(owner.ptr)
    buf._ph_owner = owner            # the array (and every slice of it) references buf, buf keeps the allocation
    return np.frombuffer(buf, dtype=dt, count=n).reshape(shape)


def pinned_from(arr: np.ndarray) -> np.ndarray:
    out = pinned_empty(arr.shape, arr.dtype)
    out[...] = arr
    return out


class Stream:
    """A CUDA stream of the library (ph_stream_create).  `with Stream() as s:` makes every array operation
    inside the block launch on it; `s.wait(other)` orders it behind what `other` has queued so far."""

    def __init__(self):
        _lib.init()
        p = C.c_void_p()
        check(_lib.load().ph_stream_create(C.byref(p)))
        self.handle = p.value
        self._saved = []

    def __enter__(self):
        lib = _lib.load()
        self._saved.append(lib.ph_stream())
        check(lib.ph_set_stream(self.handle))
        return self

    def __exit__(self, *exc):
        check(_lib.load().ph_set_stream(self._saved.pop()))
        return False

    def wait(self, other: "Stream" = None) -> None:
        check(_lib.load().ph_stream_wait(self.handle, other.handle if other is not None else None))

    def synchronize(self) -> None:
        check(_lib.load().ph_stream_sync(self.handle))

    def close(self) -> None:
        if self.handle:
            check(_lib.load().ph_stream_destroy(self.handle))
            self.handle = None


def main_stream_wait(s: Stream) -> None:
    """The library's own stream waits for everything queued so far on `s`."""
    check(_lib.load().ph_stream_wait(None, s.handle))


def _scalar_of(value, dtype: np.dtype, what: str) -> np.ndarray:
    """A scalar operand as ONE element of the array's dtype.  The device path computes in the
    array's element type only (same-dtype operands): a scalar that the type cannot
    hold exactly (`int_arr * 2.5`, `u8_arr + 300`, `int_arr < 2.5`) would silently change the
    result, so it is rejected; the reference would promote (Int32 * Float64 -> Float64)."""
    if isinstance(value, np.generic) or (isinstance(value, np.ndarray) and value.ndim == 0):
        value = value.item()
    if isinstance(value, (bool, np.bool_)):
        value = int(value) if dtype.kind != "b" else bool(value)
    if dtype.kind == "f":
        if isinstance(value, (int, float, np.integer, np.floating)):
            return np.array(value, dtype=dtype)           # Float32 arrays take any real (rounded like a Float32 literal)
        raise TypeError(f"device path: {what} needs a real scalar for a {dtype} array, got {type(value).__name__}")
    if dtype.kind in "iub":
        if isinstance(value, (float, np.floating)):
            if not float(value).is_integer():
                raise TypeError(f"device path: scalar {value!r} is not representable in {dtype} ({what}); "
                                "mixed Int/Float arithmetic is off the device path -- convert the array first")
            value = int(value)
        if isinstance(value, (int, np.integer)):
            if dtype.kind == "b":
                if int(value) not in (0, 1):
                    raise TypeError(f"device path: scalar {value!r} is not a Bool ({what})")
                return np.array(bool(value), dtype=dtype)
            info = np.iinfo(dtype)
            if not info.min <= int(value) <= info.max:
                raise TypeError(f"device path: scalar {value!r} does not fit {dtype} ({what})")
            return np.array(int(value), dtype=dtype)
    raise TypeError(f"device path: {what} cannot take a {type(value).__name__} scalar for a {dtype} array")


class _Buffer:
    """Ref-counted owner of one device allocation (reshape aliases the buffer,
    src/n_array.cr:429-433; views keep their source alive, src/view.cr:7)."""

    def __init__(self, nbytes: int):
        _lib.init()
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        lib = _lib.load()
        check(lib.ph_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.stream = lib.ph_stream()        # the pool block is released on the stream it was handed out on

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                _lib.load().ph_free_on(self.ptr, self.stream)
                self.ptr = None
        except Exception:
            pass


class _SubBuffer:
    """A byte range of another buffer (one slice of a batched `slices` copy).  Keeps its parent
    alive and never frees: the parent releases the whole allocation when the last slice dies."""

    def __init__(self, parent, byte_offset: int, nbytes: int):
        self.parent = parent
        self.nbytes = int(nbytes)
        self.ptr = parent.ptr + int(byte_offset)


_BIN = {"+": "PH_ADD", "-": "PH_SUB", "*": "PH_MUL", "/": "PH_DIV", "//": "PH_FLOORDIV", "%": "PH_MOD",
        "**": "PH_POW", "&+": "PH_WADD", "&-": "PH_WSUB", "&*": "PH_WMUL", "&**": "PH_WPOW",
        "&": "PH_AND", "|": "PH_OR", "^": "PH_XOR"}
_CMP = {">": "PH_GT", "<": "PH_LT", ">=": "PH_GE", "<=": "PH_LE", "==": "PH_EQ", "!=": "PH_NE"}
_RED = {"sum": "PH_SUM", "min": "PH_MIN", "max": "PH_MAX", "argmax": "PH_ARGMAX", "argmin": "PH_ARGMIN"}


class _Indexable:
    """What DeviceNArray and DeviceView share: a buffer, a descriptor, a shape, a dtype
    (the device-side MultiIndexable, src/multi_indexable.cr:30-65)."""
    shape: List[int]
    dtype: np.dtype
    _buf: _Buffer

    def desc(self) -> PhDesc:
        raise NotImplementedError

    @property
    def ptr(self) -> int:
        return self._buf.ptr

    @property
    def size(self) -> int:
        """MultiIndexable#size; NB the empty shape [] has size 0 (shape_util.cr:41-50)."""
        if len(self.shape) == 0:
            return 0
        n = 1
        for s in self.shape:
            n *= int(s)
        return n

    @property
    def dimensions(self) -> int:
        return len(self.shape)

    def __repr__(self) -> str:
        """No transfer: printing the elements is `to_host()` first (the reference's Formatter walks elements)."""
        return f"<{type(self).__name__} {self.dtype} shape={self.shape}>"

    # ---- small host-side queries (src/multi_indexable.cr:100-237) -----------------------
    def empty(self) -> bool:
        """MultiIndexable#empty? (:107-109)."""
        return self.size == 0

    def scalar(self) -> bool:
        """MultiIndexable#scalar? (:119-121): exactly one element, in any number of dimensions."""
        return self.size == 1

    def to_scalar(self):
        """MultiIndexable#to_scalar (:131-137): the sole element, ShapeError otherwise.  This is how a
        fully indexed chunk (`state[1]`, a 1-element array) becomes a number in the reference."""
        if not self.scalar():
            raise ShapeError(f"Only single-element MultiIndexables can be converted to scalars, but this one has "
                             f"{self.size} elements (shape: {self.shape}).")
        return self.first()

    def to_scalar_or_none(self):
        """MultiIndexable#to_scalar? (:145-148)."""
        return self.first() if self.scalar() else None

    def to_f(self) -> float:
        """MultiIndexable#to_f (:160-162)."""
        return float(self.to_scalar())

    def first(self):
        """MultiIndexable#first (:176-182): the element at the zero coordinate."""
        if self.size == 0:
            raise ShapeError(f"This MultiIndexable has zero elements (shape: {self.shape}).")
        return self.get([0] * len(self.shape))

    def last(self):
        """MultiIndexable#last (:197-203): the element at the largest coordinate."""
        if self.size == 0:
            raise ShapeError(f"This MultiIndexable has zero elements (shape: {self.shape}).")
        return self.get([s - 1 for s in self.shape])

    def sample(self, random=None):
        """MultiIndexable#sample (:234-237): one element at a random coordinate (one tiny D2H)."""
        import random as _random
        if self.size == 0:
            raise ShapeError(f"Can't sample empty collection. (shape: {self.shape})")
        r = random or _random
        return self.get([r.randrange(s) for s in self.shape])

    # ---- blocks are out of scope on the device path --------------------------------
    def _no_blocks(self, *a, **k):
        raise DeviceBlockError("arbitrary blocks cannot run on the device path")

    map = map_with = each_with = map_with_coord = apply = process = build = each = each_coord = fast_each = _no_blocks

    @property
    def buffer(self):
        """Buffered#buffer (src/buffered/buffered.cr:10) on a device array raises: no silent D2H."""
        raise DeviceBlockError("a device NArray has no host buffer; call to_host explicitly")

    # ---- views (src/multi_indexable.cr:881-895, src/view.cr) --------------------------
    def view(self, *literal, drop: bool = True) -> "DeviceView":
        v = DeviceView(self._buf, self.desc(), self.shape, self.dtype)
        return v.view(*literal, drop=drop) if literal else v

    mutable_view = view

    # ---- gather: [] / get_chunk (src/multi_indexable.cr:338-356, 523-531) ---------------
    def get_chunk(self, literal: Sequence, drop=True) -> "DeviceNArray":
        """get_chunk(region_literal, drop) (:354-356) and, with a second list, get_chunk(coord, region_shape)
        (:369-395): the block of `region_shape` whose lowermost corner is `coord`."""
        if isinstance(drop, (list, tuple)):
            return self._get_chunk_at(list(literal), list(drop))
        reg = literal if isinstance(literal, PhRegion) else make_region(literal, self.shape, drop)
        return self.unsafe_fetch_chunk(reg)

    def _get_chunk_at(self, coord: List[int], region_shape: List[int]) -> "DeviceNArray":
        if len(coord) != len(region_shape):
            raise DimensionError(f"'coord' ({coord}) and 'region_shape' {region_shape} had a different number of dimensions. Note "
                                 "that you must fully specify your coordinate and region shape for this overload of get_chunk.")
        if len(coord) != len(self.shape):
            raise DimensionError(f"'coord' ({coord}) had a different number of dimensions than this MultiIndexable (must have "
                                 f"{len(self.shape)}, but has {len(coord)}).")
        for idx, (c, r) in enumerate(zip(coord, region_shape)):
            if c < 0:
                raise CrArgumentError(f"'coord' {coord} was negative on axis {idx}, but must be strictly nonnegative.")
            if r < 0:
                raise CrArgumentError(f"'region_shape' {region_shape} was negative on axis {idx}, but must be strictly nonnegative.")
            if c + r > self.shape[idx]:
                raise ShapeError(f"The region defined by shape {region_shape} and lowermost coordinate {coord} is not contained "
                                 f"within this MultiIndexable on axis {idx} (this MultiIndexable has {self.shape[idx]} elements "
                                 f"on axis {idx}).")
        reg = cover_region(region_shape)                              # IndexRegion.cover(region_shape).translate!(coord)
        host_check(_lib.load().ph_region_translate(C.byref(reg), _i64(coord), len(coord)))
        return self.unsafe_fetch_chunk(reg)

    def unsafe_fetch_chunk(self, reg: PhRegion) -> "DeviceNArray":
        """NArray#unsafe_fetch_chunk (src/n_array.cr:450-453): one gather launch."""
        src = PhDesc()
        host_check(_lib.load().ph_desc_region(C.byref(self.desc()), C.byref(reg), C.byref(src)))
        out = DeviceNArray(reg.shape, self.dtype)
        if out.size:
            check(_lib.load().ph_copy_strided(self.dtype.itemsize, self.ptr, C.byref(src), out.ptr, C.byref(out.desc())))
        return out

    def get_available(self, literal: Sequence, drop: bool = True):
        """MultiIndexable#get_available (:411-413): the literal is trimmed to fit, never raises
        for an over-long range."""
        lits = [_region.marshal(l) for l in literal]
        arr = (PhRangeLit * max(1, len(lits)))(*lits)
        reg = PhRegion()
        host_check(_lib.load().ph_region_new_trimmed(arr, len(lits), _i64(self.shape), _i64(self.shape),
                                                     len(self.shape), int(drop), C.byref(reg)))
        return self.unsafe_fetch_chunk(reg)

    def has_region(self, literal: Sequence, drop: bool = True) -> bool:
        """MultiIndexable#has_region? (:313-318)."""
        try:
            make_region(literal, self.shape, drop)
            return True
        except (CrIndexError, DimensionError):
            return False

    def get_chunk_or_none(self, literal: Sequence, drop: bool = True):
        """MultiIndexable#[]? (:540-546): nil instead of raising."""
        return self.get_chunk(literal, drop) if self.has_region(literal, drop) else None

    def __getitem__(self, key):
        if isinstance(key, _Indexable):                      # narr[mask] returns self (:479-481)
            return self
        if not isinstance(key, tuple):
            key = (key,)
        return self.get_chunk(list(key))

    def get(self, *coord):
        """MultiIndexable#get / get_element (:567-575): one element, canonicalising the coord."""
        if len(coord) == 1 and isinstance(coord[0], (list, tuple)):
            coord = tuple(coord[0])
        canon = (C.c_int64 * max(1, len(coord)))()
        host_check(_lib.load().ph_canonicalize_coord(_i64(coord), len(coord), _i64(self.shape), len(self.shape), canon))
        off = C.c_int64()
        host_check(_lib.load().ph_desc_offset_of(C.byref(self.desc()), canon, len(coord), C.byref(off)))
        out = np.empty(1, dtype=self.dtype)
        _checked_d2h(out.ctypes.data, self.ptr + off.value * self.dtype.itemsize, self.dtype.itemsize)
        return out[0]

    get_element = get

    # ---- scatter / fill / mask store (src/multi_writable.cr:55-84, 127-172) ----------------
    def set_element(self, coord, value) -> None:
        canon = (C.c_int64 * max(1, len(coord)))()
        host_check(_lib.load().ph_canonicalize_coord(_i64(coord), len(coord), _i64(self.shape), len(self.shape), canon))
        off = C.c_int64()
        host_check(_lib.load().ph_desc_offset_of(C.byref(self.desc()), canon, len(coord), C.byref(off)))
        v = _scalar_of(value, self.dtype, "set_element").reshape(1)
        check(_lib.load().ph_h2d(self.ptr + off.value * self.dtype.itemsize, v.ctypes.data, self.dtype.itemsize))
        check(_lib.load().ph_sync())

    def set_chunk(self, literal: Sequence, value) -> None:
        reg = literal if isinstance(literal, PhRegion) else make_region(literal, self.shape)
        if isinstance(value, _Indexable):
            ok = C.c_int32()
            host_check(_lib.load().ph_shapes_compatible(_i64(value.shape), len(value.shape), _i64(reg.shape),
                                                        len(reg.shape), C.byref(ok)))
            if not ok.value:                                 # multi_writable.cr:58-60
                raise ShapeError(f"Cannot substitute: the given array has shape {value.shape}, but the region has "
                                 f"shape {reg.shape}.")
        self.unsafe_set_chunk(reg, value)

    def unsafe_set_chunk(self, reg: PhRegion, value) -> None:
        """NArray#unsafe_set_chunk (src/n_array.cr:484-500): one scatter / fill launch.  An array
        source is streamed in ITS OWN lex order onto the region's lex order."""
        lib = _lib.load()
        dst = PhDesc()
        host_check(lib.ph_desc_region(C.byref(self.desc()), C.byref(reg), C.byref(dst)))
        nreg = 1
        for i in range(dst.rank):
            nreg *= dst.extent[i]
        if nreg == 0:
            return
        if isinstance(value, _Indexable):
            if value.dtype != self.dtype:
                raise TypeError("device path: source and destination must share a dtype")
            # compatible_shapes? allows trailing ones: view the source with the region's extents
            sdesc = PhDesc()
            reg_shape = [int(dst.extent[i]) for i in range(dst.rank)]
            st = lib.ph_desc_reshape(C.byref(value.desc()), _i64(reg_shape), len(reg_shape), C.byref(sdesc))
            if st == K["PH_HOST_NEEDS_COPY"]:
                value = value.to_narr()
                st = lib.ph_desc_reshape(C.byref(value.desc()), _i64(reg_shape), len(reg_shape), C.byref(sdesc))
            host_check(st)
            check(lib.ph_copy_strided(self.dtype.itemsize, value.ptr, C.byref(sdesc), self.ptr, C.byref(dst)))
        else:
            v = _scalar_of(value, self.dtype, "[]=")
            check(lib.ph_fill_region(self.dtype.itemsize, self.ptr, C.byref(dst), v.ctypes.data))

    def set_mask(self, mask: "_Indexable", value) -> None:
        """NArray#[]=(mask, value) (src/n_array.cr:510-551)."""
        lib = _lib.load()
        if list(mask.shape) != list(self.shape):             # :511-513 (DimensionError on NArray)
            raise DimensionError("Cannot perform masking: mask shape does not match array shape.")
        if mask.dtype not in (np.dtype(np.bool_), np.dtype(np.uint8)):
            raise TypeError("mask must be a Bool array")
        if isinstance(value, _Indexable):
            if list(value.shape) != list(self.shape):
                raise DimensionError("Cannot perform masking: value shape does not match array shape.")
            self._same_dtype(value, "masked store")
            check(lib.ph_mask_set_array(self.dtype.itemsize, self.ptr, C.byref(self.desc()), mask.ptr,
                                        C.byref(mask.desc()), value.ptr, C.byref(value.desc())))
        else:
            v = _scalar_of(value, self.dtype, "masked store")
            check(lib.ph_mask_set_scalar(self.dtype.itemsize, self.ptr, C.byref(self.desc()), mask.ptr,
                                         C.byref(mask.desc()), v.ctypes.data))

    def __setitem__(self, key, value):
        if isinstance(key, _Indexable):
            return self.set_mask(key, value)
        if not isinstance(key, tuple):
            key = (key,)
        self.set_chunk(list(key), value)

    # ---- copies through views: permute / reverse / reshape (src/multi_indexable.cr:795-803) --
    def to_narr(self) -> "DeviceNArray":
        """View#to_narr (src/view.cr:123-126) / MultiIndexable#to_narr (:852-856): one gather."""
        out = DeviceNArray(self.shape, self.dtype)
        if out.size:
            check(_lib.load().ph_copy_strided(self.dtype.itemsize, self.ptr, C.byref(self.desc()), out.ptr,
                                              C.byref(out.desc())))
        return out

    def to_host(self, check_flags: bool = True) -> np.ndarray:
        """Explicit device -> host transfer of the (materialised) contents.  A guaranteed raise point:
        OverflowError / DivisionByZeroError / ArgumentError of ANY operator launched before it is raised
        here, in the same synchronisation as the copy (the reference raises at the operator; on the device
        path the error surfaces at the first read of a result, INTEGRATION.md "raise points").
        `check_flags=False` is the raw read the parity tests use to compare wrapped values AND flags."""
        src = self if isinstance(self, DeviceNArray) else self.to_narr()
        out = np.empty(src.shape, dtype=src.dtype)
        if check_flags:
            _checked_d2h(out.ctypes.data if out.size else None, src.ptr, out.nbytes if out.size else 0)
        elif out.size:
            check(_lib.load().ph_d2h(out.ctypes.data, src.ptr, out.nbytes))
        return out

    # ---- slices / tile (src/multi_indexable.cr:742-786, 818-843) ----------------------------
    def slices(self, axis: int = 0) -> List["DeviceNArray"]:
        """MultiIndexable#slices (:776-786): the chunks `self[.., i, ..]` for every index i of `axis`,
        each an independent array without that axis.  The reference gathers them one by one
        (ChunkIterator -> unsafe_fetch_chunk per index); on the device that is one launch per slice
        and launch-bound for every axis but the leading one.  All slices along `axis` together ARE
        the array with `axis` moved to the front, so they are produced by ONE permuting copy and
        handed out as consecutive ranges of its buffer (disjoint, so still independent arrays)."""
        nd = len(self.shape)
        if not 0 <= axis < nd:
            raise CrIndexError(f"axis {axis} is not present in a {nd}-dimensional MultiIndexable")
        n = self.shape[axis]
        if n == 0 or self.size == 0:
            rest = [s for i, s in enumerate(self.shape) if i != axis] or [1]
            return [DeviceNArray(rest, self.dtype) for _ in range(n)]
        order = [axis] + [i for i in range(nd) if i != axis]
        moved = self.view().permute(*order).to_narr()                  # one launch
        rest = moved.shape[1:] or [1]
        step = self.size // n * self.dtype.itemsize
        return [DeviceNArray(rest, self.dtype, _SubBuffer(moved._buf, i * step, step)) for i in range(n)]

    def each_slice(self, axis: int = 0):
        """MultiIndexable#each_slice (:742-748): the slices `self[.., i, ..]` one after the other.  On the
        device each one is a DeviceView over the SOURCE buffer -- a descriptor, no copy and no launch -- so
        the reference's per-axis idiom (`each_slice(axis) { |s| acc = acc + s }`, :742-786) costs only the
        consumer's kernel, which reads the strided slice directly.  `slices` (independent arrays, one
        batched permuting copy) stays for callers that need owned copies."""
        nd = len(self.shape)
        if not 0 <= axis < nd:
            raise CrIndexError(f"axis {axis} is not present in a {nd}-dimensional MultiIndexable")
        base = self.view()
        d = base.desc()
        rest_ext = [int(d.extent[i]) for i in range(nd) if i != axis]
        rest_str = [int(d.stride[i]) for i in range(nd) if i != axis]
        if not rest_ext:                                         # slicing a vector: 1-element slices of shape [1]
            rest_ext, rest_str = [1], [1]
        step = int(d.stride[axis])
        for i in range(self.shape[axis]):
            yield DeviceView(self._buf, PhDesc.make(rest_ext, rest_str, int(d.offset) + i * step), rest_ext, self.dtype)

    def tile(self, counts: Sequence[int]) -> "DeviceNArray":
        """MultiIndexable#tile (:818-827): out[c] = self[c % shape]; as a descriptor every axis
        becomes (count, extent) with strides (0, stride) -- no modulo on the device."""
        if len(counts) != len(self.shape):
            raise DimensionError("tile counts have the wrong number of dimensions")
        if 2 * len(self.shape) > _lib.PH_MAX_RANK:
            raise ShapeError("tile supports rank <= 4 on the device path")
        d = self.desc()
        ext, strd = [], []
        for i, c in enumerate(counts):
            ext += [int(c), int(d.extent[i])]
            strd += [0, int(d.stride[i])]
        src = PhDesc.make(ext, strd, d.offset)
        out = DeviceNArray([s * int(c) for s, c in zip(self.shape, counts)], self.dtype)
        if out.size:
            check(_lib.load().ph_copy_strided(self.dtype.itemsize, self.ptr, C.byref(src), out.ptr,
                                              C.byref(PhDesc.contiguous(ext))))
        return out

    # ---- elementwise (src/multi_indexable.cr:931-985) -------------------------------------
    def _result_dtype(self, op: str):
        if op == "/" and self.dtype.kind in "iu":
            return np.dtype(np.float64)
        return self.dtype

    def _same_dtype(self, other: "_Indexable", what: str) -> None:
        """Kernels are typed by ONE element type: a second operand of another dtype would be read with
        the wrong element size (past its allocation).  The Crystal and C++ layers reject this at compile
        time (both operands are DeviceIndexable(T)); here it is a TypeError before any launch."""
        a = np.dtype(np.uint8) if self.dtype == np.dtype(np.bool_) else self.dtype
        b = np.dtype(np.uint8) if other.dtype == np.dtype(np.bool_) else other.dtype
        if a != b:
            raise TypeError(f"device path: {what} needs operands of one dtype, got {self.dtype} and {other.dtype}")

    def _binary(self, op: str, other, reflected: bool = False) -> "DeviceNArray":
        lib = _lib.load()
        code = K[_BIN[op]]
        dt = dtype_code(self.dtype)
        if isinstance(other, _Indexable):
            self._same_dtype(other, f"'{op}'")
            a, b = (other, self) if reflected else (self, other)
            if list(a.shape) != list(b.shape):                # multi_indexable.cr:935-940
                raise ShapeError(f"The shape of this MultiIndexable ({a.shape}) does not match the shape of "
                                 f"the one provided ({b.shape}), so '{op}' cannot be applied element-wise.")
            out = DeviceNArray(a.shape, self._result_dtype(op))
            check(lib.ph_ewise_binary(code, dt, a.ptr, C.byref(a.desc()), b.ptr, C.byref(b.desc()), out.ptr,
                                      C.byref(out.desc())))
            return out
        # scalar: array.map &.op(other) (:947-951) / Number#op(narr) (patches/number.cr:6-15)
        out = DeviceNArray(self.shape, self._result_dtype(op))
        if op == "**" and self.dtype.kind == "f" and isinstance(other, (int, np.integer)) and not reflected:
            s = np.array(other, dtype=np.int32)
            code = K["PH_POWI"]
        else:
            s = _scalar_of(other, self.dtype, f"'{op}'")
        check(lib.ph_ewise_scalar(code, dt, self.ptr, C.byref(self.desc()), s.ctypes.data, int(reflected), out.ptr,
                                  C.byref(out.desc())))
        return out

    def bcast_desc(self, shape) -> PhDesc:
        out = PhDesc()
        host_check(_lib.load().ph_desc_broadcast(C.byref(self.desc()), _i64(shape), len(shape), C.byref(out)))
        return out

    def broadcast_op(self, op: str, other: "_Indexable") -> "DeviceNArray":
        """NEW (ShapeUtil.broadcast_shapes): equal rank, size-1 axes stretch."""
        if len(self.shape) != len(other.shape):
            raise ShapeError("broadcast requires equal rank")
        self._same_dtype(other, f"broadcast '{op}'")
        shape = (C.c_int64 * max(1, len(self.shape)))()
        host_check(_lib.load().ph_broadcast_shapes(_i64(self.shape), _i64(other.shape), len(self.shape), shape))
        shape = [int(shape[i]) for i in range(len(self.shape))]
        out = DeviceNArray(shape, self._result_dtype(op))
        da, db = self.bcast_desc(shape), other.bcast_desc(shape)
        check(_lib.load().ph_ewise_binary(K[_BIN[op]], dtype_code(self.dtype), self.ptr, C.byref(da),
                                          other.ptr, C.byref(db), out.ptr, C.byref(out.desc())))
        return out

    def mul_add(self, b: "_Indexable", c: "_Indexable") -> "DeviceNArray":
        """Fused (self * b) + c with two roundings; b and c may broadcast."""
        self._same_dtype(b, "mul_add")
        self._same_dtype(c, "mul_add")
        out = DeviceNArray(self.shape, self.dtype)
        da, db, dc = self.desc(), b.bcast_desc(self.shape), c.bcast_desc(self.shape)
        check(_lib.load().ph_ewise_mul_add(dtype_code(self.dtype), self.ptr, C.byref(da), b.ptr, C.byref(db),
                                           c.ptr, C.byref(dc), out.ptr, C.byref(out.desc())))
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
        check(_lib.load().ph_ewise_unary(K[name], dtype_code(self.dtype), self.ptr, C.byref(self.desc()), out.ptr,
                                         C.byref(out.desc())))
        return out

    def __pos__(self): return self._unary("PH_POS")
    def __neg__(self): return self._unary("PH_NEG")
    def __invert__(self): return self._unary("PH_NOT")

    def _compare(self, op: str, other, reflected=False, eq_style=False) -> "DeviceNArray":
        lib = _lib.load()
        out = DeviceNArray(self.shape, np.bool_)
        dt = dtype_code(self.dtype)
        if isinstance(other, _Indexable):
            if list(other.shape) != list(self.shape):
                if eq_style:                                    # multi_indexable.cr:900-902
                    raise DimensionError("Cannot compute the element-wise equality: shapes differ")
                raise ShapeError("shapes differ")               # :935-940
            self._same_dtype(other, f"'{op}'")
            check(lib.ph_compare(K[_CMP[op]], dt, self.ptr, C.byref(self.desc()), other.ptr, C.byref(other.desc()),
                                 out.ptr, C.byref(out.desc())))
        else:
            s = _scalar_of(other, self.dtype, f"'{op}'")
            check(lib.ph_compare_scalar(K[_CMP[op]], dt, self.ptr, C.byref(self.desc()), s.ctypes.data, int(reflected),
                                        out.ptr, C.byref(out.desc())))
        return out

    def __gt__(self, o): return self._compare(">", o)
    def __lt__(self, o): return self._compare("<", o)
    def __ge__(self, o): return self._compare(">=", o)
    def __le__(self, o): return self._compare("<=", o)
    def eq(self, o): return self._compare("==", o, eq_style=True)   # MultiIndexable#eq (:899-913)
    def match(self, value): return self._compare("==", value)        # MultiIndexable#=~ (:916-920)

    def cmp(self, other) -> "DeviceNArray":
        """`<=>` of the operator list (src/multi_indexable.cr:960-981): -1 / 0 / 1 as an Int32 array.  Integer
        element types only: Float#<=> is Int32? (nil against NaN), which has no device representation."""
        lib = _lib.load()
        if self.dtype.kind not in "iu":
            raise TypeError("device path: <=> is defined for integer element types (Float#<=> is nilable)")
        out = DeviceNArray(self.shape, np.int32)
        dt = dtype_code(self.dtype)
        if isinstance(other, _Indexable):
            if list(other.shape) != list(self.shape):
                raise ShapeError(f"The shape of this MultiIndexable ({self.shape}) does not match the shape of "
                                 f"the one provided ({other.shape}), so '<=>' cannot be applied element-wise.")
            self._same_dtype(other, "'<=>'")
            check(lib.ph_compare3(dt, self.ptr, C.byref(self.desc()), other.ptr, C.byref(other.desc()), out.ptr,
                                  C.byref(out.desc())))
        else:
            s = _scalar_of(other, self.dtype, "'<=>'")
            check(lib.ph_compare3_scalar(dt, self.ptr, C.byref(self.desc()), s.ctypes.data, 0, out.ptr, C.byref(out.desc())))
        return out

    def checksum64(self, word_offset: int = 0) -> int:
        """Position-weighted 64-bit checksum of a contiguous array's bytes (ph_checksum64): a verification
        aid -- shards add up (pass the global 8-byte-word index of the shard's first word)."""
        if not isinstance(self, DeviceNArray):
            return self.to_narr().checksum64(word_offset)
        out = C.c_uint64(0)
        check(_lib.load().ph_checksum64(self.ptr, self.size * self.dtype.itemsize, int(word_offset), C.byref(out)))
        return out.value

    def to_host_async(self, out: np.ndarray) -> None:
        """Device -> host into a PINNED array (`pinned_empty`), asynchronous: the data is valid after the
        next `sync()` (which is also the raise point for data-dependent errors) or Stream.synchronize()."""
        src = self if isinstance(self, DeviceNArray) else self.to_narr()
        if list(out.shape) != list(src.shape) or out.dtype != src.dtype or not out.flags["C_CONTIGUOUS"]:
            raise ShapeError(f"to_host_async needs a contiguous {src.dtype} destination of shape {src.shape}")
        lib = _lib.load()
        if out.size:
            check(lib.ph_d2h_async(out.ctypes.data, src.ptr, out.nbytes))
        # `src` may be a temporary: its block is released on the stream it was allocated on (stream-ordered,
        # so behind this copy when that is the current stream; otherwise that stream is made to wait for it)
        home = getattr(src._buf, "stream", None)
        cur = lib.ph_stream()
        if home is not None and home != cur:
            check(lib.ph_stream_wait(home, cur))

    def equals(self, other: "_Indexable") -> bool:
        """NArray#== (src/n_array.cr:440-447) for two device arrays."""
        if not isinstance(other, _Indexable) or list(other.shape) != list(self.shape) or other.dtype != self.dtype:
            return False
        if self.size == 0:
            return True
        v = self.eq(other).min()
        return bool(v)

    # ---- reductions (Enumerable over NArray#each, src/n_array.cr:556-564) ---------------------
    def _reduce_full(self, name: str):
        lib = _lib.load()
        if self.size == 0:
            if name == "sum":
                return self.dtype.type(0)
            raise CrEmptyError("Empty enumerable")
        dt = np.dtype(np.uint8) if self.dtype == np.dtype(np.bool_) else self.dtype
        # record mode (ph_reduce_full_sharded on one process = no exchange): ONE launch whose finishing block
        # writes value, index and the pending arithmetic flags into a pinned host record the call polls --
        # no device-to-host copy, no stream synchronisation, no second read for the flags
        val, idx, flags = _RED_CELLS
        st = lib.ph_reduce_full_sharded(K[_RED[name]], dtype_code(dt), self.ptr, C.byref(self.desc()), 0, val,
                                        C.byref(idx), C.byref(flags))
        if st:
            check(st)
        if flags.value:
            raise_for_flags(flags.value)
        return np.frombuffer(val, dtype=dt, count=1)[0].copy(), idx.value

    def sum(self, axis: Optional[int] = None):
        if axis is not None:
            return self._reduce_axis("sum", axis)
        r = self._reduce_full("sum")
        return r[0] if isinstance(r, tuple) else r

    def max(self, axis: Optional[int] = None):
        return self._reduce_axis("max", axis) if axis is not None else self._reduce_full("max")[0]

    def min(self, axis: Optional[int] = None):
        return self._reduce_axis("min", axis) if axis is not None else self._reduce_full("min")[0]

    def argmax(self, axis: Optional[int] = None):
        """README.md:56-61 idiom: (max, coord of the FIRST maximum); per axis: Int64 indices."""
        if axis is not None:
            return self._reduce_axis("argmax", axis)
        v, i = self._reduce_full("argmax")
        return v, self.index_to_coord(i)

    def argmin(self, axis: Optional[int] = None):
        if axis is not None:
            return self._reduce_axis("argmin", axis)
        v, i = self._reduce_full("argmin")
        return v, self.index_to_coord(i)

    def index_to_coord(self, index: int) -> List[int]:
        """Buffered.index_to_coord (src/buffered/buffered.cr:66-77) on the lex index."""
        coord = []
        for length in reversed(self.shape):
            coord.append(index % length)
            index //= length
        return list(reversed(coord))

    def _reduce_axis(self, name: str, axis: int, raise_now: bool = True) -> "DeviceNArray":
        if axis < 0 or axis >= len(self.shape):
            raise CrIndexError(f"axis {axis} is not present in a {len(self.shape)}-dimensional MultiIndexable")
        if self.shape[axis] == 0 and name != "sum":
            raise CrEmptyError("Empty enumerable")
        out_shape = [s for i, s in enumerate(self.shape) if i != axis] or [1]
        out_dtype = np.dtype(np.int64) if name.startswith("arg") else self.dtype
        out = DeviceNArray(out_shape, out_dtype)
        if out.size:
            if self.shape[axis] == 0:
                out.set_chunk([], 0)
            else:
                check(_lib.load().ph_reduce_axis(K[_RED[name]], dtype_code(self.dtype), self.ptr, C.byref(self.desc()),
                                                 axis, out.ptr, C.byref(out.desc())))
        if raise_now:                               # (a sharded fold checks once, after the cross-rank combine)
            DeviceNArray.raise_pending()
        return out

    # ---- data-dependent errors -----------------------------------------------------------
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


def _concat_shape(shapes, axis: int):
    """ph_concat_shape (include/ph_host.h): the reference's `compatible?` + the concatenated shape."""
    n = len(shapes)
    packed = (C.c_int64 * (n * _lib.PH_MAX_RANK))()
    ranks = (C.c_int32 * n)()
    for k, sh in enumerate(shapes):
        if len(sh) > _lib.PH_MAX_RANK:
            raise ShapeError(f"the device path supports rank <= {_lib.PH_MAX_RANK}")
        ranks[k] = len(sh)
        for i, d in enumerate(sh):
            packed[k * _lib.PH_MAX_RANK + i] = int(d)
    out = (C.c_int64 * _lib.PH_MAX_RANK)()
    ax = C.c_int32(0)
    host_check(_lib.load().ph_concat_shape(packed, ranks, n, int(axis), out, C.byref(ax)))
    return [int(out[i]) for i in range(len(shapes[0]))], ax.value


def _concatenate(narrs, axis: int) -> "DeviceNArray":
    first = narrs[0]
    for a in narrs:
        if a.dtype != first.dtype:
            raise TypeError("device path: concatenate needs arrays of one dtype")
    shape, ax = _concat_shape([list(a.shape) for a in narrs], axis)
    out = DeviceNArray(shape, first.dtype)
    at = 0
    for a in narrs:
        n = int(a.shape[ax])
        if n and out.size:
            lit = [_region.ALL] * len(shape)
            lit[ax] = _region.rng(at, at + n - 1)
            out.unsafe_set_chunk(make_region(lit, shape, False), a)
        at += n
    return out


_RED_CELLS = ((C.c_uint64 * 2)(), C.c_int64(-1), C.c_uint32(0))      # reused ctypes cells of the full reductions


class DeviceNArray(_Indexable):
    """Row-major N-D array resident in HBM (mirror of Phase::NArray, src/n_array.cr:15)."""

    def __init__(self, shape: Sequence[int], dtype, buf: Optional[_Buffer] = None):
        self.shape = [int(s) for s in shape]
        self.dtype = np.dtype(dtype)
        self._buf = buf if buf is not None else _Buffer(max(1, self.size) * self.dtype.itemsize)

    def desc(self) -> PhDesc:
        # built once per shape (a Python loop over the axes), handed out as a copy: callers edit descriptors
        c = self.__dict__.get("_desc_cache")
        if c is None or c[0] != self.shape:
            c = (list(self.shape), PhDesc.contiguous(self.shape))
            self.__dict__["_desc_cache"] = c
        d = PhDesc()
        C.memmove(C.byref(d), C.byref(c[1]), C.sizeof(PhDesc))
        return d

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
    def from_host_async(cls, arr: np.ndarray) -> "DeviceNArray":
        """Host -> device from a PINNED, contiguous array (`pinned_empty`): returns at once, the copy is
        ordered on the current stream like any operator (a pageable source would make it synchronous)."""
        _lib.init()
        if not arr.flags["C_CONTIGUOUS"]:
            raise ShapeError("from_host_async needs a contiguous (pinned) source")
        out = cls(arr.shape, arr.dtype)
        if arr.size:
            check(_lib.load().ph_h2d(out.ptr, arr.ctypes.data, arr.nbytes))
        return out

    @classmethod
    def fill(cls, shape, value, dtype) -> "DeviceNArray":
        """NArray.fill (src/n_array.cr:230-232)."""
        out = cls(shape, dtype)
        if out.size:
            v = _scalar_of(value, out.dtype, "fill")
            check(_lib.load().ph_fill_region(out.dtype.itemsize, out.ptr, C.byref(out.desc()), v.ctypes.data))
        return out

    def clone(self) -> "DeviceNArray":
        """NArray#clone deep-copies (src/n_array.cr:372-374)."""
        out = DeviceNArray(self.shape, self.dtype)
        if self.size:
            check(_lib.load().ph_d2d(out.ptr, self.ptr, self.size * self.dtype.itemsize))
        return out

    dup = clone

    def reshape(self, *new_shape) -> "DeviceNArray":
        """NArray#reshape ALIASES the buffer (src/n_array.cr:429-433)."""
        if len(new_shape) == 1 and isinstance(new_shape[0], (list, tuple)):
            new_shape = tuple(new_shape[0])
        n = 1
        for s in new_shape:
            n *= int(s)
        if (n if new_shape else 0) != self.size:
            raise ShapeError(f"Cannot change shape from {self.shape} to {list(new_shape)}: reshape cannot add or "
                             "remove elements.")
        return DeviceNArray(new_shape, self.dtype, self._buf)

    def flatten(self) -> "DeviceNArray":
        return self.reshape(self.size)

    # ---- joins (src/n_array.cr:321-344, 666-750): one strided copy per input, nothing touches the host ---------
    def concatenate(self, *others, axis: int = 0) -> "DeviceNArray":
        """NArray#concatenate(*others, axis) and, called on the class with the arrays as arguments,
        NArray.concatenate(*narrs, axis) (src/n_array.cr:712-750): the inputs laid side by side along `axis`.
        The shape rule is the reference's `compatible?` (ph_concat_shape: a negative axis excludes nothing, so
        every dimension must then match).  Each input is ONE scatter into its range of the result."""
        return _concatenate((self,) + tuple(others), axis)

    def push(self, *others, axis: int = 0) -> "DeviceNArray":
        """NArray#push / << (src/n_array.cr:682-710), in place: the buffers are appended as they lie and only
        shape[0] grows, whatever `axis` says (`axis` merely relaxes the compatibility test, as in the reference)."""
        if not others:
            return self
        shapes = [self.shape] + [list(o.shape) for o in others]
        _concat_shape(shapes, axis)                                   # compatible? -> DimensionError
        for o in others:
            if o.dtype != self.dtype:
                raise TypeError("device path: push needs arrays of one dtype")
        total = self.size + sum(o.size for o in others)
        rows = self.shape[0] + sum(o.shape[0] for o in others)
        new_shape = [rows] + self.shape[1:]
        n = 1
        for d in new_shape:
            n *= d
        if n != total:                                                # an axis != 0 let through rows that do not fill the shape
            raise ShapeError(f"Cannot change shape from [{total}] to {new_shape}: reshape cannot add or remove elements.")
        lib = _lib.load()
        isz = self.dtype.itemsize
        buf = _Buffer(max(1, total) * isz)
        at = 0
        for a in (self,) + tuple(others):
            flat = a if isinstance(a, DeviceNArray) else a.to_narr()
            if flat.size:
                check(lib.ph_d2d(buf.ptr + at * isz, flat.ptr, flat.size * isz))
            at += flat.size
        self.shape = new_shape
        self._buf = buf
        self.__dict__.pop("_desc_cache", None)
        return self

    def __lshift__(self, other) -> "DeviceNArray":
        return self.push(other)

    @staticmethod
    def wrap(*narrs) -> "DeviceNArray":
        """NArray.wrap(*objects, pad: false) (src/n_array.cr:321-340): a new leading axis with one input per row;
        identical shapes or DimensionError (padding is NotImplementedError in the reference as well)."""
        if not narrs:
            raise DimensionError("Cannot wrap these arrays: nothing to wrap")
        container = list(narrs[0].shape)
        if any(list(a.shape) != container for a in narrs):
            raise DimensionError("Cannot wrap these arrays: shapes do not match. Pass argument pad:true if you want to "
                                 "reshape arrays as necessary.")
        rows = [(a if isinstance(a, DeviceNArray) else a.to_narr()).reshape([1] + container) for a in narrs]
        return _concatenate(rows, 0)

    def permute(self, *order) -> "DeviceNArray":
        """MultiIndexable#permute = view.permute + copy (src/multi_indexable.cr:795-803)."""
        return self.view().permute(*order).to_narr()

    def reverse(self) -> "DeviceNArray":
        return self.view().reverse().to_narr()


class DeviceView(_Indexable):
    """Lazy view: source buffer + ONE descriptor.  Region / Permute / Reverse transforms
    (src/view_util/transforms.cr) are affine in the coordinate, so a chain of them folds into
    (offset, extent[], stride[]) as it is built; Reshape folds when it is expressible in
    strides and otherwise materialises first.  Reads gather, writes scatter
    (MutableView, src/mutable_view.cr:16-18)."""

    def __init__(self, buf: _Buffer, desc: PhDesc, shape, dtype):
        self._buf = buf
        self._desc = desc
        self.shape = [int(s) for s in shape]
        self.dtype = np.dtype(dtype)

    def desc(self) -> PhDesc:
        d = PhDesc()
        C.memmove(C.byref(d), C.byref(self._desc), C.sizeof(PhDesc))
        return d

    def clone(self) -> "DeviceView":
        return DeviceView(self._buf, self.desc(), self.shape, self.dtype)

    def view(self, *literal, drop: bool = True) -> "DeviceView":
        """View#view / restrict_to (src/view.cr:36-56)."""
        if len(literal) == 1 and isinstance(literal[0], (list, PhRegion)):
            literal = literal[0]
        if isinstance(literal, PhRegion):
            reg = literal
        elif len(literal) == 0:
            return self.clone()
        else:
            reg = make_region(list(literal), self.shape, drop)
        out = PhDesc()
        host_check(_lib.load().ph_desc_region(C.byref(self._desc), C.byref(reg), C.byref(out)))
        return DeviceView(self._buf, out, reg.shape, self.dtype)

    mutable_view = view

    def unsafe_fetch_chunk(self, reg: PhRegion) -> "DeviceView":
        """View#unsafe_fetch_chunk returns a view (src/view.cr:105-107)."""
        return self.view(reg)

    def get_chunk(self, literal: Sequence, drop=True) -> "DeviceView":
        if isinstance(drop, (list, tuple)):                          # get_chunk(coord, region_shape): a view of the block
            return self._get_chunk_at(list(literal), list(drop))
        reg = literal if isinstance(literal, PhRegion) else make_region(literal, self.shape, drop)
        return self.view(reg)

    def permute(self, *order) -> "DeviceView":
        """View#permute! (src/view.cr:72-81); no argument = reversed axes."""
        if len(order) == 1 and isinstance(order[0], (list, tuple)):
            order = tuple(order[0])
        out = PhDesc()
        if order:
            pat = (C.c_int32 * len(order))(*[int(o) for o in order])
            host_check(_lib.load().ph_desc_permute(C.byref(self._desc), pat, len(order), C.byref(out)))
        else:
            host_check(_lib.load().ph_desc_permute(C.byref(self._desc), None, 0, C.byref(out)))
        return DeviceView(self._buf, out, [int(out.extent[i]) for i in range(out.rank)], self.dtype)

    def reverse(self) -> "DeviceView":
        """View#reverse! (src/view.cr:96-99): every axis flipped."""
        out = PhDesc()
        host_check(_lib.load().ph_desc_reverse(C.byref(self._desc), C.byref(out)))
        return DeviceView(self._buf, out, self.shape, self.dtype)

    def reshape(self, *new_shape) -> "DeviceView":
        """View#reshape! (src/view.cr:58-66)."""
        if len(new_shape) == 1 and isinstance(new_shape[0], (list, tuple)):
            new_shape = tuple(new_shape[0])
        out = PhDesc()
        st = _lib.load().ph_desc_reshape(C.byref(self._desc), _i64(new_shape), len(new_shape), C.byref(out))
        if st == K["PH_HOST_NEEDS_COPY"]:
            return self.to_narr().view().reshape(*new_shape)
        host_check(st)
        return DeviceView(self._buf, out, list(new_shape), self.dtype)
