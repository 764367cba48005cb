"""CPU oracle for ph-core's data-parallel hot path -- TEST INFRASTRUCTURE ONLY.

This file is a restatement, in numpy + small pure-Python loops, of what the
Crystal reference (in-phase/ph-core) computes on the path SURVEY.md section 8
names.  It exists so that the CUDA path can be checked against it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import it; nothing under ``ph-core_b200/`` does (the product path fails loudly
without its CUDA library -- it never falls back to this file).

Parity status (see DESIGN.md "Oracle"):
  * index math, region literals, lex order, gather / scatter / mask store, view
    transforms, integer elementwise ops: PINNED against every golden vector the
    reference's spec/ holds for the path (tests/test_oracle_goldens.py).
  * floating-point elementwise results, reductions, the heat stencil: the
    reference's own tests pin none of these and no Crystal compiler exists in
    this image, so these functions are "parity unpinned": they restate Crystal
    1.0.0 stdlib number semantics (not vendored under /root/reference;
    shard.yml:8 pins `crystal: 1.0.0`) by construction only.
  * broadcasting, per-axis reductions, the N-D stencil: the reference has no
    implementation; they are DEFINED here by composing reference operators
    (tile + op, each_slice + fold, slice arithmetic) as SURVEY.md 8(a) states.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np


# --------------------------------------------------------------------------
# Exceptions (src/exceptions/exceptions.cr:4-22 + Crystal stdlib classes)
# --------------------------------------------------------------------------
class ShapeError(Exception):
    """src/exceptions/exceptions.cr:4-12"""


class DimensionError(ShapeError):
    """src/exceptions/exceptions.cr:14-22 (DimensionError < ShapeError)"""


class CrIndexError(Exception):
    """Crystal stdlib IndexError (range_syntax.cr:120-122, coord_util.cr:44-46)."""


class CrOverflowError(Exception):
    """Crystal stdlib OverflowError (checked Int + - * and unary -)."""


class CrDivisionByZeroError(Exception):
    """Crystal stdlib DivisionByZeroError (Int // %, Float %)."""


class CrArgumentError(Exception):
    """Crystal stdlib ArgumentError (MIN // -1, negative int exponent, NaN in max/min)."""


class CrEmptyError(Exception):
    """Crystal stdlib Enumerable::EmptyError (max/min of an empty collection)."""


# --------------------------------------------------------------------------
# Region literals.  Crystal writes `a..b`, `a...b`, `a..s..b`; Python has no
# such syntax, so a literal is an int or an `R` (a Crystal Range whose begin or
# end may itself be an R, exactly the nesting parse_range accepts).
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class R:
    begin: Union[None, int, "R"] = None
    end: Union[None, int, "R"] = None
    exclusive: bool = False

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        dots = "..." if self.exclusive else ".."
        f = lambda v: "" if v is None else (f"({v!r})" if isinstance(v, R) else str(v))
        return f"{f(self.begin)}{dots}{f(self.end)}"


@dataclass(frozen=True)
class Step:
    """Crystal `a.step(by: s, to: b, exclusive: e)` (range_syntax.cr:62-64)."""
    current: int
    limit: Optional[int]
    step: int
    exclusive: bool = False


def rng(a=None, b=None, step=None, exclusive=False) -> R:
    """Convenience: rng(a, b) = a..b ; rng(a, b, s) = a..s..b ; exclusive -> `...`."""
    if step is None:
        return R(a, b, exclusive)
    return R(R(a, step), b, exclusive)


def _sign(x: int) -> int:
    return (x > 0) - (x < 0)


def parse_range(lit) -> dict:
    """src/range_syntax/range_syntax.cr:41-69."""
    if isinstance(lit, Step):                                   # :62-64
        return dict(first=lit.current, last=lit.limit, step=lit.step, exclusive=lit.exclusive)
    if isinstance(lit, (int, np.integer)):                      # :66-69
        return dict(first=int(lit), last=int(lit), step=1, exclusive=False)
    first, last = lit.begin, lit.end
    if isinstance(first, R):                                    # (a..b)..c   :43-49
        if last is None or isinstance(last, (int, np.integer)):
            return dict(first=first.begin, last=last, step=first.end, exclusive=lit.exclusive)
    elif first is None or isinstance(first, (int, np.integer)):
        if isinstance(last, R):                                 # a..(b..c)   :52-53
            return dict(first=first, last=last.end, step=last.begin, exclusive=last.exclusive)
        if last is None or isinstance(last, (int, np.integer)):  # a..c       :54-55
            return dict(first=first, last=last, step=None, exclusive=lit.exclusive)
    raise ValueError("poorly formatted range")                  # :59


def canonicalize_index_unsafe(index: int, size: int) -> int:
    """src/coord_util.cr:65-71."""
    return size + index if index < 0 else index


def has_index(index: int, size: int) -> bool:
    """src/coord_util.cr:19-21."""
    return index < size and index >= -size


def canonicalize_index(index: int, size: int) -> int:
    """src/coord_util.cr:43-48."""
    if not has_index(index, size):
        raise CrIndexError(f"{index} is not a valid index for an axis of length {size}")
    return canonicalize_index_unsafe(index, size)


def canonicalize_coord(coord: Sequence[int], shape: Sequence[int]) -> List[int]:
    """src/coord_util.cr:76-82."""
    if len(coord) != len(shape):
        raise DimensionError("coord / shape dimension mismatch")
    return [canonicalize_index(c, s) for c, s in zip(coord, shape)]


def has_coord(coord, shape) -> bool:
    """src/coord_util.cr:30-33."""
    if len(coord) != len(shape):
        return False
    return all(has_index(c, s) for c, s in zip(coord, shape))


def get_size(first: int, last: int, step: int) -> int:
    """src/range_syntax/range_syntax.cr:7-17."""
    cmp = (last > first) - (last < first)
    if last != first and _sign(step) != cmp:
        raise CrIndexError("step direction disagrees with first and last")
    if last >= first:
        if step == 0:
            raise CrDivisionByZeroError()
        return (last - first) // step + 1
    return (first - last) // (-step) + 1


def infer_range(lit, bound: int) -> dict:
    """src/range_syntax/range_syntax.cr:84-136."""
    if isinstance(lit, (int, np.integer)):                      # :84-87
        c = canonicalize_index_unsafe(int(lit), bound)
        return dict(first=c, step=1, last=c, size=1)
    vals = parse_range(lit)
    f, l, step = vals["first"], vals["last"], vals["step"]
    if step is None:                                            # :96-101
        first = 0 if f is None else canonicalize_index_unsafe(f, bound)
        temp_last = bound - 1 if l is None else canonicalize_index_unsafe(l, bound)
        step = 1 if temp_last >= first else -1
    else:                                                       # :102-105
        first = (0 if step > 0 else bound - 1) if f is None else canonicalize_index_unsafe(f, bound)
        temp_last = (bound - 1 if step > 0 else 0) if l is None else canonicalize_index_unsafe(l, bound)
    if l is not None and vals["exclusive"]:                     # :108-114
        if temp_last == first:
            return dict(first=0, step=0, last=0, size=0)
        temp_last -= _sign(step)
    if first < 0 or temp_last < 0:                              # :120-122
        raise CrIndexError("endpoint negative after canonicalization")
    size = get_size(first, temp_last, step)                     # :128 (IndexError re-raised :131-134)
    last = first + step * (size - 1)
    return dict(first=first, step=int(step), last=last, size=size)


def canonicalize_range(lit, bound: int) -> dict:
    """src/range_syntax/range_syntax.cr:138-153."""
    r = infer_range(lit, bound)
    if not (0 <= r["last"] < bound and 0 <= r["first"] < bound):
        raise CrIndexError(f"{lit} is not a sensible index range for axis of length {bound}")
    return r


# --------------------------------------------------------------------------
# ShapeUtil / Buffered
# --------------------------------------------------------------------------
def compatible_shapes(shape1: Sequence[int], shape2: Sequence[int]) -> bool:
    """src/shape_util.cr:6-32 -- equal up to trailing ones."""
    shape1, shape2 = list(shape1), list(shape2)
    if len(shape1) == 0 or len(shape2) == 0:
        return shape1 == shape2
    if len(shape1) > len(shape2):
        larger, shared = shape1, len(shape2)
    else:
        larger, shared = shape2, len(shape1)
    for i in range(shared):
        if shape1[i] != shape2[i]:
            return False
    return all(larger[i] == 1 for i in range(shared, len(larger)))


def shape_to_size(shape: Sequence[int]) -> int:
    """src/shape_util.cr:41-50 -- NB: the empty shape [] has size 0."""
    if len(shape) == 0:
        return 0
    return int(np.prod([int(s) for s in shape], dtype=object))


def axis_strides(shape: Sequence[int]) -> List[int]:
    """src/buffered/buffered.cr:15-24."""
    ret = list(shape)
    ret[-1] = 1
    for idx in range(len(ret) - 2, -1, -1):
        ret[idx] = ret[idx + 1] * shape[idx + 1]
    return ret


def coord_to_index_fast(coord, strides) -> int:
    """src/buffered/buffered.cr:44-52."""
    return sum(int(c) * int(s) for c, s in zip(coord, strides))


def index_to_coord(index: int, shape: Sequence[int]) -> List[int]:
    """src/buffered/buffered.cr:66-77."""
    coord = [0] * len(shape)
    for dim, length in enumerate(reversed(shape)):
        coord[dim] = index % length
        index //= length
    return list(reversed(coord))


# --------------------------------------------------------------------------
# IndexRegion (src/index_region.cr:40-705)
# --------------------------------------------------------------------------
class IndexRegion:
    def __init__(self, first, step, last, proper_shape, drop: bool, degeneracy=None):
        """src/index_region.cr:298-302."""
        self.first = list(first)
        self.step = list(step)
        self.last = list(last)
        self.proper_shape = list(proper_shape)
        self.drop = drop
        self.degeneracy = list(degeneracy) if degeneracy is not None else [False] * len(self.proper_shape)
        self.reduced_shape = self.compute_reduced_shape(self.proper_shape, self.degeneracy, self.drop)

    # -- constructors ------------------------------------------------------
    @classmethod
    def new(cls, literal: Sequence, bound_shape: Sequence[int], drop: bool = True) -> "IndexRegion":
        """src/index_region.cr:192-224 (region literal + bound shape)."""
        n = len(bound_shape)
        first, step, last, shape = [0] * n, [1] * n, [0] * n, [0] * n
        degeneracy = [False] * n
        if len(literal) > n:                                    # :199-201
            raise DimensionError("region literal has more dimensions than its bound shape")
        for i, lit in enumerate(literal):                       # :203-213
            r = canonicalize_range(lit, bound_shape[i])
            first[i], step[i], last[i], shape[i] = r["first"], r["step"], r["last"], r["size"]
            if isinstance(lit, (int, np.integer)):
                degeneracy[i] = drop
        for axis in range(len(literal), n):                     # :217-221
            last[axis] = bound_shape[axis] - 1
            shape[axis] = bound_shape[axis]
        return cls(first, step, last, shape, drop, degeneracy)

    @classmethod
    def new_trimmed(cls, literal: Sequence, trim_to: Sequence[int], bound_shape=None, drop: bool = True):
        """src/index_region.cr:133-168 (trim_to: constructor)."""
        n = len(trim_to)
        first, step, last, shape = [0] * n, [1] * n, [0] * n, [0] * n
        degeneracy = [False] * n
        allow_relative = bound_shape is not None
        bound_shape = bound_shape if bound_shape is not None else trim_to
        for i, lit in enumerate(literal):
            if not allow_relative:
                ensure_nonnegative(lit)
            r = infer_range(lit, bound_shape[i])
            first[i], step[i], last[i], shape[i] = r["first"], r["step"], r["last"], r["size"]
            if isinstance(lit, (int, np.integer)):
                degeneracy[i] = drop
        for axis in range(len(literal), len(bound_shape)):
            last[axis] = bound_shape[axis] - 1
            shape[axis] = bound_shape[axis]
        return cls(first, step, last, shape, drop, degeneracy).trim(trim_to, inplace=True)

    @classmethod
    def cover(cls, bound_shape: Sequence[int], drop: bool = True, degeneracy=None) -> "IndexRegion":
        """src/index_region.cr:232-238."""
        first = [0] * len(bound_shape)
        step = [0 if x == 0 else 1 for x in bound_shape]
        last = [max(0, x - 1) for x in bound_shape]
        return cls(first, step, last, list(bound_shape), drop, degeneracy)

    @classmethod
    def absolute(cls, literal: Sequence, drop: bool = True) -> "IndexRegion":
        """src/index_region.cr:248-275 (absolute literal, no bound shape)."""
        n = len(literal)
        first, step, last, shape = [0] * n, [0] * n, [0] * n, [0] * n
        degeneracy = [False] * n
        for i, lit in enumerate(literal):
            ensure_nonnegative(lit)
            if not bounded(lit):
                raise ValueError("cannot create IndexRegion without an explicit upper bound")
            if isinstance(lit, (int, np.integer)):
                degeneracy[i] = True
            r = infer_range(lit, 0)
            first[i], step[i], last[i], shape[i] = r["first"], r["step"], r["last"], r["size"]
        return cls(first, step, last, shape, drop, degeneracy)

    @staticmethod
    def compute_reduced_shape(proper_shape, degeneracy, drop) -> List[int]:
        """src/index_region.cr:323-335 (+ drop_degenerate :304-321)."""
        if not drop:
            return list(proper_shape)
        kept = [v for v, d in zip(proper_shape, degeneracy) if not d]
        if not kept:
            return [shape_to_size(proper_shape)]
        return kept

    # -- queries -------------------------------------------------------------
    @property
    def shape(self) -> List[int]:
        """src/index_region.cr:339-341 (shape_internal = reduced shape)."""
        return list(self.reduced_shape)

    @property
    def proper_dimensions(self) -> int:
        return len(self.proper_shape)

    def clone(self) -> "IndexRegion":
        return IndexRegion(self.first, self.step, self.last, self.proper_shape, self.drop, self.degeneracy)

    def key(self):
        """def_equals_and_hash @first, @step, @last, @degeneracy, @drop (:98)."""
        return (tuple(self.first), tuple(self.step), tuple(self.last), tuple(self.degeneracy), self.drop)

    def __eq__(self, other):
        return isinstance(other, IndexRegion) and self.key() == other.key()

    def includes(self, coord) -> bool:
        """src/index_region.cr:445-459."""
        if len(coord) != self.proper_dimensions:
            return False
        for i, ordn in enumerate(coord):
            lo, hi = (self.first[i], self.last[i]) if self.step[i] > 0 else (self.last[i], self.first[i])
            if not (lo <= ordn <= hi):
                return False
            if (ordn - self.first[i]) % self.step[i] != 0:
                return False
        return True

    def fits_in(self, bound_shape) -> bool:
        """src/index_region.cr:468-478."""
        if len(bound_shape) != self.proper_dimensions:
            raise DimensionError("fits_in? with a different number of dimensions")
        return all(b > max(f, l) for b, f, l in zip(bound_shape, self.first, self.last))

    # -- mutation --------------------------------------------------------------
    @staticmethod
    def trim_axis(new_bound, first, step, last, size):
        """src/index_region.cr:684-703."""
        if first >= new_bound:
            if last >= new_bound:
                return 0, 0, 0, 0
            elif step < 0:
                span = (new_bound - 1) - last
                size = span // abs(step) + 1
                span -= span % abs(step)
                return last + span, step, last, size
        elif step > 0 and last >= new_bound:
            span = (new_bound - 1) - first
            span -= span % abs(step)
            size = span // abs(step) + 1
            return first, step, first + span, size
        return first, step, last, size

    def trim(self, bound_shape, inplace=False) -> "IndexRegion":
        """src/index_region.cr:502-515 (trim!) / :545-547 (trim)."""
        tgt = self if inplace else self.clone()
        if len(bound_shape) != tgt.proper_dimensions:
            raise DimensionError("trim! with a different number of dimensions")
        for axis, size in enumerate(bound_shape):
            tgt.first[axis], tgt.step[axis], tgt.last[axis], tgt.proper_shape[axis] = IndexRegion.trim_axis(
                size, tgt.first[axis], tgt.step[axis], tgt.last[axis], tgt.proper_shape[axis])
        tgt.reduced_shape = IndexRegion.compute_reduced_shape(tgt.proper_shape, tgt.degeneracy, tgt.drop)
        return tgt

    def reverse(self, inplace=False) -> "IndexRegion":
        """src/index_region.cr:533-537."""
        tgt = self if inplace else self.clone()
        tgt.first, tgt.last = tgt.last, tgt.first
        tgt.step = [-s for s in tgt.step]
        return tgt

    def translate(self, offset, inplace=False) -> "IndexRegion":
        """src/index_region.cr:577-585 (+ unsafe_translate! :555-561)."""
        tgt = self if inplace else self.clone()
        for axis, amount in enumerate(offset):
            if amount < 0 and (tgt.first[axis] < -amount or tgt.last[axis] < -amount):
                raise CrIndexError("Can't translate to negative indices")
        for axis, amount in enumerate(offset):
            tgt.first[axis] += amount
            tgt.last[axis] += amount
        return tgt

    # -- coordinate maps ----------------------------------------------------------
    def local_to_absolute_unsafe(self, coord) -> List[int]:
        """src/index_region.cr:621-637."""
        if self.drop:
            out, local_axis = [], 0
            for i, degenerate in enumerate(self.degeneracy):
                if degenerate:
                    out.append(self.first[i])
                else:
                    local_axis += 1
                    out.append(self.first[i] + coord[local_axis - 1] * self.step[i])
            return out
        return [self.first[i] + o * self.step[i] for i, o in enumerate(coord)]

    def absolute_to_local_unsafe(self, coord) -> List[int]:
        """src/index_region.cr:655-664."""
        local = [(o - self.first[i]) // self.step[i] for i, o in enumerate(coord)]
        if self.drop:
            kept = [v for v, d in zip(local, self.degeneracy) if not d]
            return kept if kept else [0]
        return local

    def absolute_to_local(self, coord) -> List[int]:
        """src/index_region.cr:646-651."""
        if not self.includes(coord):
            raise CrIndexError("coordinate does not exist in region")
        return self.absolute_to_local_unsafe(coord)

    def each(self):
        """src/index_region.cr:667-670 -- LexIterator over the absolute coords."""
        return lex_coords(self.first, self.step, self.last)


def ensure_nonnegative(lit) -> None:
    """src/range_syntax/range_syntax.cr:71-82."""
    if isinstance(lit, (int, np.integer)):
        if lit < 0:
            raise CrIndexError("negative index without a bounding shape")
        return
    vals = parse_range(lit)
    for v in (vals["first"], vals["last"]):
        if v is not None and v < 0:
            raise CrIndexError("negative index without a bounding shape")


def bounded(lit) -> bool:
    """src/range_syntax/range_syntax.cr:19-35."""
    if isinstance(lit, (int, np.integer)):
        return True
    vals = parse_range(lit)
    step = vals["step"]
    if step is None or step >= 0:
        return vals["last"] is not None
    return vals["first"] is not None


# --------------------------------------------------------------------------
# Coordinate iteration (src/iterators/stride_iterator.cr:110-122,
# lex_iterator.cr:13-25, colex_iterator.cr:13-25)
# --------------------------------------------------------------------------
def lex_coords(first, step, last):
    """Yields coords in lexicographic order: the last axis varies fastest."""
    if len(first) == 0:
        return
    if any(s == 0 for s in step):                              # stride_iterator.cr:114
        return
    coord = list(first)
    yield list(coord)
    n = len(coord)
    while True:
        i = n - 1
        while True:                                             # lex_iterator.cr:14-23
            if coord[i] == last[i]:
                coord[i] = first[i]
                if i == 0:
                    return
                i -= 1
            else:
                coord[i] += step[i]
                break
        yield list(coord)


def colex_coords(first, step, last):
    """src/iterators/colex_iterator.cr:13-25 -- the first axis varies fastest."""
    if len(first) == 0 or any(s == 0 for s in step):
        return
    coord = list(first)
    yield list(coord)
    n = len(coord)
    while True:
        i = 0
        while True:
            if coord[i] == last[i]:
                coord[i] = first[i]
                if i == n - 1:
                    return
                i += 1
            else:
                coord[i] += step[i]
                break
        yield list(coord)


def lex_buffer_indices(region: IndexRegion, shape: Sequence[int]) -> List[int]:
    """Indexed::LexIterator.new(region, shape): the buffer index visited at each step
    (src/buffered/indexed/stride_iterator.cr:10-22, lex_iterator.cr:7-21)."""
    if len(region.reduced_shape) == 0:
        raise DimensionError('cannot iterate over empty shape "[]"')
    strides = axis_strides(shape)
    return [coord_to_index_fast(c, strides) for c in lex_coords(region.first, region.step, region.last)]


# --------------------------------------------------------------------------
# NArray chunk get / set / mask (src/n_array.cr:450-551) on a numpy array that
# plays the row-major buffer.  These are the per-element restatements; the
# vectorised twins used for big parity cases live further below.
# --------------------------------------------------------------------------
def fetch_chunk(narr: np.ndarray, region: IndexRegion) -> np.ndarray:
    """NArray#unsafe_fetch_chunk, src/n_array.cr:450-453."""
    buf = narr.reshape(-1)
    out_shape = region.shape
    n = shape_to_size(out_shape)
    idx = lex_buffer_indices(region, narr.shape)[:n] if n else []
    out = np.empty(n, dtype=narr.dtype)
    for k, i in enumerate(idx):
        out[k] = buf[i]
    return out.reshape(out_shape)


def get_chunk(narr: np.ndarray, literal: Sequence, drop: bool = True) -> np.ndarray:
    """MultiIndexable#[] / get_chunk(region_literal), src/multi_indexable.cr:354-356, 523-525."""
    return fetch_chunk(narr, IndexRegion.new(literal, list(narr.shape), drop))


def get_chunk_at(narr: np.ndarray, coord: Sequence[int], region_shape: Sequence[int]) -> np.ndarray:
    """MultiIndexable#get_chunk(coord, region_shape) (src/multi_indexable.cr:369-395): the block of `region_shape`
    whose lowermost corner is `coord`; both fully specified, nonnegative, contained."""
    if len(coord) != len(region_shape):
        raise DimensionError("'coord' and 'region_shape' had a different number of dimensions")
    if len(coord) != narr.ndim:
        raise DimensionError("'coord' had a different number of dimensions than this MultiIndexable")
    for idx, (c, r) in enumerate(zip(coord, region_shape)):
        if c < 0 or r < 0:
            raise CrArgumentError(f"negative on axis {idx}, but must be strictly nonnegative")
        if c + r > narr.shape[idx]:
            raise ShapeError(f"The region is not contained within this MultiIndexable on axis {idx}")
    out = np.empty(tuple(region_shape), dtype=narr.dtype)
    for local in np.ndindex(*region_shape):                          # cover(region_shape).translate!(coord), lex order
        out[local] = narr[tuple(c + l for c, l in zip(coord, local))]
    return out


def set_chunk_array(narr: np.ndarray, region: IndexRegion, src: np.ndarray) -> None:
    """NArray#unsafe_set_chunk(region, src), src/n_array.cr:484-492: src is streamed
    in ITS OWN lex order onto the region's lex order."""
    buf = narr.reshape(-1)
    idx = lex_buffer_indices(region, narr.shape)
    flat = src.reshape(-1)
    for k in range(flat.size):
        buf[idx[k]] = flat[k]


def set_chunk_scalar(narr: np.ndarray, region: IndexRegion, value) -> None:
    """NArray#unsafe_set_chunk(region, value), src/n_array.cr:495-500."""
    buf = narr.reshape(-1)
    for i in lex_buffer_indices(region, narr.shape):
        buf[i] = value


def set_chunk(narr: np.ndarray, literal: Sequence, value) -> None:
    """MultiWritable#set_chunk / []=, src/multi_writable.cr:55-70, 77-84."""
    region = IndexRegion.new(literal, list(narr.shape))
    if isinstance(value, np.ndarray):
        if not compatible_shapes(list(value.shape), region.shape):      # :58-60
            raise ShapeError("cannot substitute array of this shape into region")
        set_chunk_array(narr, region, value)
    else:
        set_chunk_scalar(narr, region, value)


def mask_set(narr: np.ndarray, mask: np.ndarray, value) -> None:
    """NArray#[]=(mask, value), src/n_array.cr:510-551."""
    if list(mask.shape) != list(narr.shape):
        raise DimensionError("mask shape does not match array shape")
    buf, m = narr.reshape(-1), mask.reshape(-1)
    if isinstance(value, np.ndarray):
        v = value.reshape(-1)
        for i in range(buf.size):
            if m[i]:
                buf[i] = v[i]
    else:
        for i in range(buf.size):
            if m[i]:
                buf[i] = value


# --------------------------------------------------------------------------
# Views: transform chain (src/view_util/transforms.cr, src/view.cr)
# --------------------------------------------------------------------------
class RegionTransform:
    """src/view_util/transforms.cr:191-222."""
    def __init__(self, region: IndexRegion):
        self.region = region

    def apply(self, coord):
        return self.region.local_to_absolute_unsafe(list(coord))


class PermuteTransform:
    """src/view_util/transforms.cr:224-271."""
    def __init__(self, pattern_or_size):
        if isinstance(pattern_or_size, (int, np.integer)):      # :236-238 default = reversed axes
            size = int(pattern_or_size)
            self.pattern = [size - i - 1 for i in range(size)]
        else:
            self.pattern = list(pattern_or_size)

    def permute(self, src_coord):                               # :249-253
        return [src_coord[self.pattern[i]] for i in range(len(self.pattern))]

    def apply(self, coord):                                     # unpermute :260-270
        out = list(self.pattern)
        for idx, el in enumerate(self.pattern):
            out[el] = coord[idx]
        return out


class ReverseTransform:
    """src/view_util/transforms.cr:273-305."""
    def __init__(self, shape):
        self.shape = list(shape)

    def apply(self, coord):                                     # :298-304
        return [self.shape[i] - 1 - el for i, el in enumerate(coord)]


class ReshapeTransform:
    """src/view_util/transforms.cr:118-189."""
    def __init__(self, src_shape, new_shape):
        self.src_shape = list(src_shape)
        self.new_shape = list(new_shape)
        self.view_axis_strides = axis_strides(self.new_shape)

    def apply(self, coord):                                     # :184-188
        index = coord_to_index_fast(coord, self.view_axis_strides)
        return index_to_coord(index, self.src_shape)


class View:
    """src/view.cr:2-127 over a numpy array (the source NArray)."""
    def __init__(self, src: np.ndarray, shape=None, transforms=None):
        self.src = src
        self.shape = list(src.shape) if shape is None else list(shape)
        self.transforms = list(transforms) if transforms else []     # newest first (compose! = unshift, :46-48)

    def clone(self) -> "View":
        return View(self.src, self.shape, self.transforms)

    def view(self, literal=None) -> "View":                     # view.cr:36-40, 43-49
        v = self.clone()
        if literal is not None:
            region = literal if isinstance(literal, IndexRegion) else IndexRegion.new(literal, v.shape)
            v.shape = region.shape
            v.transforms.insert(0, RegionTransform(region))
        return v

    def permute(self, order=None) -> "View":                    # view.cr:72-81
        v = self.clone()
        if order is not None:
            for axis in order:
                if axis < 0 or axis >= len(v.shape):
                    raise CrIndexError("axis not present")
        pt = PermuteTransform(order if order is not None else len(v.shape))
        v.shape = pt.permute(v.shape)
        v.transforms.insert(0, pt)
        return v

    def reverse(self) -> "View":                                # view.cr:96-99
        v = self.clone()
        v.transforms.insert(0, ReverseTransform(v.shape))
        return v

    def reshape(self, new_shape) -> "View":                     # view.cr:58-66
        v = self.clone()
        if shape_to_size(new_shape) != shape_to_size(v.shape):
            raise ShapeError("reshape cannot add or remove elements")
        v.transforms.insert(0, ReshapeTransform(v.shape, new_shape))
        v.shape = list(new_shape)
        return v

    def src_coord(self, coord) -> List[int]:                    # transforms.cr:77-91
        c = list(coord)
        for t in self.transforms:
            c = t.apply(c)
        return c

    def fetch_element(self, coord):                             # view.cr:109-117
        return self.src[tuple(self.src_coord(coord))]

    def set_element(self, coord, value) -> None:                # mutable_view.cr:16-18
        self.src[tuple(self.src_coord(coord))] = value

    def to_narr(self) -> np.ndarray:                            # view.cr:123-126
        n = shape_to_size(self.shape)
        out = np.empty(n, dtype=self.src.dtype)
        cover = IndexRegion.cover(self.shape)
        for k, c in enumerate(lex_coords(cover.first, cover.step, cover.last)):
            if k >= n:
                break
            out[k] = self.fetch_element(c)
        return out.reshape(self.shape)

    def set_chunk(self, literal, value) -> None:
        """MultiWritable#unsafe_set_chunk default (src/multi_writable.cr:29-44) through
        MutableView#unsafe_set_element."""
        region = IndexRegion.new(literal, self.shape)
        coords = lex_coords(region.first, region.step, region.last)
        if isinstance(value, np.ndarray):
            if not compatible_shapes(list(value.shape), region.shape):
                raise ShapeError("cannot substitute array of this shape into region")
            flat = value.reshape(-1)
            for k, c in zip(range(flat.size), coords):
                self.set_element(c, flat[k])
        else:
            for c in coords:
                self.set_element(c, value)


# --------------------------------------------------------------------------
# tile / each_slice (src/multi_indexable.cr:742-748, 818-827)
# --------------------------------------------------------------------------
def tile(narr: np.ndarray, counts: Sequence[int]) -> np.ndarray:
    """MultiIndexable#tile: new_shape[i] = shape[i]*counts[i]; element at coord c is
    src[c % shape] (tiling_lex_iterator.cr:27-41)."""
    if len(counts) != narr.ndim:
        raise DimensionError("tile counts have the wrong number of dimensions")
    return np.tile(narr, tuple(int(c) for c in counts))


def each_slice(narr: np.ndarray, axis: int = 0):
    """MultiIndexable#each_slice(axis): chunk_shape[axis] = 1 with that axis dropped
    (multi_indexable.cr:742-748; region_iterator.cr:123-132)."""
    for i in range(narr.shape[axis]):
        yield np.take(narr, i, axis=axis)


# --------------------------------------------------------------------------
# concatenate / push / wrap (src/n_array.cr:321-344, 666-750): the joins one step above the gather / scatter path
# --------------------------------------------------------------------------
def shapes_compatible_except(first_shape: Sequence[int], other_shapes: Sequence[Sequence[int]], axis: int = -1) -> bool:
    """NArray#compatible?(*others, axis) (n_array.cr:666-673): every dimension of `first` equals the same dimension
    of every other array, except at index `axis`.  The comparison is `idx != axis` on the RAW argument: a negative
    axis excludes nothing (so `concatenate(axis: -1)` demands identical shapes -- kept, it is the reference's rule)."""
    for idx, dim in enumerate(first_shape):
        for sh in other_shapes:
            if idx >= len(sh):
                raise CrIndexError("Index out of bounds")            # Array#[] on the shorter shape
            if dim != sh[idx] and idx != axis:
                return False
    return True


def concatenate(narrs: Sequence[np.ndarray], axis: int = 0) -> np.ndarray:
    """NArray.concatenate / concatenate_to_slice (n_array.cr:722-750), loop for loop: for every index over the axes
    in FRONT of `axis` (`num_chunks`), every array in turn contributes `shape[axis] * axis_strides[axis]` elements
    from its own lexicographic iterator.  The stride is taken from the FIRST array (`narrs[0].axis_strides[axis]`)."""
    narrs = [np.asarray(a) for a in narrs]
    first = narrs[0]
    if not shapes_compatible_except(first.shape, [a.shape for a in narrs], axis):
        raise DimensionError(f"Cannot concatenate these arrays along axis {axis}: shapes do not match")
    if axis >= first.ndim or axis < -first.ndim:
        raise CrIndexError("Index out of bounds")
    concat_shape = list(first.shape)
    concat_shape[axis] = sum(a.shape[axis] for a in narrs)
    partial_chunk_size = axis_strides(first.shape)[axis]
    chunk_sizes = [a.shape[axis] * partial_chunk_size for a in narrs]
    lead = concat_shape[:axis] if axis >= 0 else concat_shape[:first.ndim + axis]     # concat_shape[...axis]
    num_chunks = shape_to_size(lead) if lead else 1                                    # ([] of Int32).product == 1
    flats = [a.reshape(-1) for a in narrs]                                             # BufferedECIterator: lex order
    pos = [0] * len(narrs)
    values = []
    for _ in range(num_chunks):
        for i, flat in enumerate(flats):
            values.append(flat[pos[i]:pos[i] + chunk_sizes[i]])
            pos[i] += chunk_sizes[i]
    out = np.concatenate(values) if values else np.zeros(0, first.dtype)
    return out.astype(first.dtype, copy=False).reshape(concat_shape)


def push(narr: np.ndarray, others: Sequence[np.ndarray], axis: int = 0) -> np.ndarray:
    """NArray#push / << (n_array.cr:682-710): the buffers are appended as they lie and ONLY shape[0] grows, whatever
    `axis` says (the reference's own TODO: "axis = 0 should not be a user modifiable parameter"); `axis` only
    relaxes the compatibility test."""
    others = [np.asarray(o) for o in others]
    if not shapes_compatible_except(narr.shape, [o.shape for o in others], axis):
        raise DimensionError(f"Cannot concatenate these arrays along axis {axis}: shapes do not match")
    flat = np.concatenate([narr.reshape(-1)] + [o.reshape(-1) for o in others])
    shape = list(narr.shape)
    shape[0] += sum(o.shape[0] for o in others)
    return flat.reshape(shape)          # raises when an `axis` != 0 let sizes through that do not fill the rows


def wrap(narrs: Sequence[np.ndarray]) -> np.ndarray:
    """NArray.wrap(*objects, pad: false) (n_array.cr:321-340): a new leading axis, one input per row; shapes must
    be identical (DimensionError)."""
    narrs = [np.asarray(a) for a in narrs]
    container = list(narrs[0].shape)
    if any(list(a.shape) != container for a in narrs):
        raise DimensionError("Cannot wrap these arrays: shapes do not match. Pass argument pad:true if you want to "
                             "reshape arrays as necessary.")
    flat = np.concatenate([a.reshape(-1) for a in narrs]) if narrs else np.zeros(0)
    return flat.reshape([len(narrs)] + container)


# --------------------------------------------------------------------------
# Elementwise number semantics (Crystal 1.0.0 stdlib).
# Every function returns (result ndarray, flags) where flags is a set of
# {"overflow", "div0", "argument"} naming the exception the reference would raise.
# --------------------------------------------------------------------------
_INT_INFO = {np.dtype(t): np.iinfo(t) for t in (np.int8, np.int16, np.int32, np.int64,
                                                 np.uint8, np.uint16, np.uint32, np.uint64)}


def _is_int(dt) -> bool:
    return np.dtype(dt).kind in "iu"


def _wrap(vals_obj: np.ndarray, dt) -> np.ndarray:
    info = _INT_INFO[np.dtype(dt)]
    span = int(info.max) - int(info.min) + 1
    lo = int(info.min)
    f = np.vectorize(lambda v: (int(v) - lo) % span + lo, otypes=[object])
    return f(vals_obj).astype(dt) if vals_obj.size else vals_obj.astype(dt)


def _obj(a) -> np.ndarray:
    return np.asarray(a).astype(object)


def _checked(vals_obj: np.ndarray, dt, flags: set) -> np.ndarray:
    info = _INT_INFO[np.dtype(dt)]
    if vals_obj.size:
        bad = np.vectorize(lambda v: int(v) < info.min or int(v) > info.max, otypes=[bool])(vals_obj)
        if bad.any():
            flags.add("overflow")
    return _wrap(vals_obj, dt)


def _int_pow(base: int, exp: int, info, checked: bool, flags: set) -> int:
    """Int#** (checked) / Int#&** (wrapping): repeated squaring, Crystal 1.0.0 int.cr."""
    if exp < 0:
        flags.add("argument")
        return 0
    span, lo = int(info.max) - int(info.min) + 1, int(info.min)

    def fit(v):
        if v < info.min or v > info.max:
            if checked:
                flags.add("overflow")
            return (v - lo) % span + lo
        return v

    result, k = 1, int(base)
    while exp > 0:
        if exp & 1:
            result = fit(result * k)
        exp >>= 1
        if exp > 0:
            k = fit(k * k)
    return result


def _powi(a: np.ndarray, n: int) -> np.ndarray:
    """Float ** Int = llvm.powi -> compiler-rt __powisf2/__powidf2: square-and-multiply
    in the operand type, reciprocal at the end for negative n."""
    dt = a.dtype
    recip = n < 0
    b = int(n)
    r = np.ones_like(a)
    a = a.copy()
    with np.errstate(all="ignore"):
        while True:
            if b & 1:
                r = (r * a).astype(dt)
            b = int(b / 2)          # C division truncates toward zero
            if b == 0:
                break
            a = (a * a).astype(dt)
        return (np.array(1, dtype=dt) / r).astype(dt) if recip else r


def result_dtype(op: str, dt) -> np.dtype:
    dt = np.dtype(dt)
    if op == "/" and _is_int(dt):
        return np.dtype(np.float64)         # Int / Int -> Float64
    return dt


def ewise(op: str, a, b) -> Tuple[np.ndarray, set]:
    """`a op b` where a, b are same-dtype arrays of identical shape, or one is a
    same-dtype scalar (numpy 0-d).  multi_indexable.cr:931-985, n_array.cr:589-595,
    patches/number.cr:6-15.  Operand order is preserved (a is the receiver)."""
    a, b = np.asarray(a), np.asarray(b)
    dt = a.dtype if a.ndim or not b.ndim else b.dtype
    if op == "**" and not _is_int(dt) and _is_int(b.dtype) and b.ndim == 0:
        return _powi(a.astype(dt), int(b)), set()
    a, b = a.astype(dt), b.astype(dt)
    flags: set = set()
    with np.errstate(all="ignore"):
        if _is_int(dt):
            info = _INT_INFO[dt]
            A, B = _obj(a), _obj(b)
            if op in ("+", "-", "*"):
                raw = {"+": A + B, "-": A - B, "*": A * B}[op]
                return _checked(np.asarray(raw, dtype=object), dt, flags), flags
            if op in ("&+", "&-", "&*"):
                raw = {"&+": A + B, "&-": A - B, "&*": A * B}[op]
                return _wrap(np.asarray(raw, dtype=object), dt), flags
            if op == "/":
                return (a.astype(np.float64) / b.astype(np.float64)), flags
            if op in ("//", "%"):
                Bb = np.broadcast_to(B, np.broadcast(A, B).shape)
                Ab = np.broadcast_to(A, Bb.shape)
                zero = np.vectorize(lambda v: int(v) == 0, otypes=[bool])(Bb) if Bb.size else np.zeros(Bb.shape, bool)
                if zero.any():
                    flags.add("div0")
                if op == "//" and dt.kind == "i" and Bb.size:
                    ovf = np.vectorize(lambda x, y: int(x) == info.min and int(y) == -1, otypes=[bool])(Ab, Bb)
                    if ovf.any():
                        flags.add("argument")          # "Overflow: MIN / -1"
                fn = (lambda x, y: 0 if int(y) == 0 else int(x) // int(y)) if op == "//" else \
                     (lambda x, y: 0 if int(y) == 0 else int(x) % int(y))
                raw = np.vectorize(fn, otypes=[object])(Ab, Bb) if Bb.size else Ab
                return _wrap(np.asarray(raw, dtype=object), dt), flags
            if op in ("**", "&**"):
                Bb = np.broadcast_to(B, np.broadcast(A, B).shape)
                Ab = np.broadcast_to(A, Bb.shape)
                fn = lambda x, y: _int_pow(int(x), int(y), info, op == "**", flags)
                raw = np.vectorize(fn, otypes=[object])(Ab, Bb) if Bb.size else Ab
                return _wrap(np.asarray(raw, dtype=object), dt), flags
            if op in ("&", "|", "^"):
                return {"&": a & b, "|": a | b, "^": a ^ b}[op].astype(dt), flags
        else:
            if op == "+":
                return (a + b).astype(dt), flags
            if op == "-":
                return (a - b).astype(dt), flags
            if op == "*":
                return (a * b).astype(dt), flags
            if op == "/":
                return (a / b).astype(dt), flags
            if op == "//":                                      # Float#// = (a / b).floor
                return np.floor((a / b).astype(dt)).astype(dt), flags
            if op == "%":                                       # Float#modulo = a - b * (a / b).floor
                if np.any(b == 0):
                    flags.add("div0")
                q = np.floor((a / b).astype(dt)).astype(dt)
                return (a - (b * q).astype(dt)).astype(dt), flags
            if op == "**":                                      # libm pow: tolerance only
                return np.power(a, b).astype(dt), flags
    raise ValueError(f"op {op!r} is not defined for dtype {dt}")


def unary(op: str, a) -> Tuple[np.ndarray, set]:
    """Unary + - ~ (multi_indexable.cr:954-958, 983-985)."""
    a = np.asarray(a)
    flags: set = set()
    if op == "+":
        return a.copy(), flags
    if op == "-":
        if _is_int(a.dtype):
            return _checked(np.asarray(-_obj(a), dtype=object), a.dtype, flags), flags
        return (-a).astype(a.dtype), flags
    if op == "~":
        if not _is_int(a.dtype):
            raise ValueError("~ is defined for integers only")
        return (~a).astype(a.dtype), flags
    raise ValueError(op)


def compare(op: str, a, b) -> np.ndarray:
    """> < >= <= (multi_indexable.cr:977-980) and eq (:899-913): NArray(Bool), IEEE for floats."""
    a, b = np.asarray(a), np.asarray(b)
    with np.errstate(all="ignore"):
        res = {">": a > b, "<": a < b, ">=": a >= b, "<=": a <= b, "==": a == b}[op]
    return np.asarray(res, dtype=np.bool_)


def check_same_shape(a_shape, b_shape, what="arith") -> None:
    """Arithmetic mismatch -> ShapeError (multi_indexable.cr:935-940); eq mismatch ->
    DimensionError (:900-902)."""
    if list(a_shape) != list(b_shape):
        raise (DimensionError if what == "eq" else ShapeError)("shape mismatch")


def broadcast_shapes(a_shape, b_shape) -> List[int]:
    """NEW rule (no reference implementation): equal ranks, each axis
    equal or one of them 1.  Its oracle is tile() then the same-shape op."""
    if len(a_shape) != len(b_shape):
        raise ShapeError("broadcast requires equal rank")
    out = []
    for x, y in zip(a_shape, b_shape):
        if x == y or y == 1:
            out.append(x)
        elif x == 1:
            out.append(y)
        else:
            raise ShapeError("shapes cannot be broadcast")
    return out


def ewise_broadcast(op: str, a: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, set]:
    """Broadcast op DEFINED as tile-to-common-shape then ewise."""
    shape = broadcast_shapes(a.shape, b.shape)
    ta = tile(a, [s // x if x else 1 for s, x in zip(shape, a.shape)])
    tb = tile(b, [s // x if x else 1 for s, x in zip(shape, b.shape)])
    return ewise(op, ta, tb)


# --------------------------------------------------------------------------
# Reductions (Crystal Enumerable over NArray#each, n_array.cr:556-564)
# --------------------------------------------------------------------------
def reduce_sum_sequential(a: np.ndarray):
    """Enumerable#sum = reduce(T.zero) { acc + e }: strict left fold IN T; ints
    overflow-check.  Pure loop: small inputs only."""
    flat = a.reshape(-1)
    if _is_int(a.dtype):
        info, acc = _INT_INFO[a.dtype], 0
        for v in flat:
            acc += int(v)
            if acc < info.min or acc > info.max:
                raise CrOverflowError()
        return a.dtype.type(acc)
    acc = a.dtype.type(0)
    with np.errstate(all="ignore"):
        for v in flat:
            acc = a.dtype.type(acc + v)
    return acc


def reduce_sum_fast(a: np.ndarray):
    """Order-free twin for big inputs: exact for ints (python ints via int64 chunks),
    f64-accumulated truth for floats (the tolerance anchor of SURVEY.md 7.4-1)."""
    if _is_int(a.dtype):
        return int(a.astype(np.int64).sum()) if a.dtype != np.uint64 else int(_obj(a).sum())
    return float(np.sum(a, dtype=np.float64))


def reduce_minmax(a: np.ndarray, which: str):
    """Enumerable#max/min: first extremum wins (strict > / < from the left); NaN ->
    ArgumentError; empty -> EmptyError."""
    flat = a.reshape(-1)
    if flat.size == 0:
        raise CrEmptyError()
    if a.dtype.kind == "f" and np.isnan(flat).any():
        raise CrArgumentError("comparison failed (NaN)")
    return flat.max() if which == "max" else flat.min()


def reduce_argmax(a: np.ndarray, which: str = "max") -> Tuple[object, int]:
    """README.md:56-61 idiom: each_with_coord + `if el > max` => FIRST extremum;
    returns (value, flat lex index); coord = index_to_coord(index, shape)."""
    flat = a.reshape(-1)
    if flat.size == 0:
        raise CrEmptyError()
    if a.dtype.kind == "f" and np.isnan(flat).any():
        raise CrArgumentError("comparison failed (NaN)")
    idx = int(np.argmax(flat) if which == "max" else np.argmin(flat))   # numpy: first occurrence
    return flat[idx], idx


def reduce_axis(a: np.ndarray, axis: int, which: str) -> np.ndarray:
    """Per-axis reduction DEFINED as the fold over each_slice(axis) in increasing index
    (SURVEY.md 8(a) a-10): out = slice_0 (op) slice_1 (op) ... with the elementwise
    reference operator, every intermediate rounded in T."""
    slices = list(each_slice(a, axis))
    if not slices:
        raise CrEmptyError()
    if which == "sum":
        acc = np.zeros_like(slices[0])                          # T.zero + s0 + s1 ...
        for s in slices:
            acc, fl = ewise("+", acc, s)
            if "overflow" in fl:
                raise CrOverflowError()
        return acc
    if which in ("max", "min"):
        if a.dtype.kind == "f" and np.isnan(a).any():
            raise CrArgumentError("comparison failed (NaN)")
        acc = slices[0].copy()
        for s in slices[1:]:
            better = (s > acc) if which == "max" else (s < acc)
            acc = np.where(better, s, acc)
        return acc
    if which in ("argmax", "argmin"):
        if a.dtype.kind == "f" and np.isnan(a).any():
            raise CrArgumentError("comparison failed (NaN)")
        acc = slices[0].copy()
        arg = np.zeros(acc.shape, dtype=np.int64)
        for i, s in enumerate(slices[1:], start=1):
            better = (s > acc) if which == "argmax" else (s < acc)
            acc = np.where(better, s, acc)
            arg = np.where(better, i, arg)
        return arg
    raise ValueError(which)


# --------------------------------------------------------------------------
# Heat equation (examples/heat_equation.cr)
# --------------------------------------------------------------------------
def heat_example_coeff() -> float:
    """examples/heat_equation.cr:5-20: (237 * 0.01) / (2700 * 900 * (0.05 ** 2))."""
    spacing_sq = float(_powi(np.array(0.05, dtype=np.float64), 2))
    return (237 * 0.01) / ((2700 * 900) * spacing_sq)


def heat_step_1d_example(state: np.ndarray, coeff) -> np.ndarray:
    """update_temp, examples/heat_equation.cr:38-51: one-sided (zero-flux) ends,
    d[i] = ((s[i-1] - 2*s[i]) + s[i+1]) * C, s' = s + d; every op rounded in T."""
    T = state.dtype.type
    c = T(coeff)
    two = T(2)
    d = np.zeros_like(state)
    with np.errstate(all="ignore"):
        d[0] = T(T(state[1] - state[0]) * c)                    # :43
        d[-1] = T(T(state[-2] - state[-1]) * c)                 # :44
        inner = state[1:-1]
        t = (state[:-2] - (two * inner).astype(state.dtype)).astype(state.dtype)
        t = (t + state[2:]).astype(state.dtype)
        d[1:-1] = (t * c).astype(state.dtype)                   # :46-48
        return (state + d).astype(state.dtype)                  # :50


def heat_simulate_1d_example(n: int = 21, steps: int = 10001, dtype=np.float64) -> np.ndarray:
    """simulate, examples/heat_equation.cr:22-36: init 20, s[0]=0, s[-1]=100."""
    state = np.full(n, 20.0, dtype=dtype)
    state[0], state[-1] = 0.0, 100.0
    c = heat_example_coeff()
    for _ in range(steps):
        state = heat_step_1d_example(state, c)
    return state


def heat_step_nd(state: np.ndarray, coeff) -> np.ndarray:
    """N-D generalisation DEFINED in SURVEY.md 8(a) a-9, in reference operators:
        c = s[1...-1, ...]; d_k = (s[lo_k] - 2*c) + s[hi_k];
        lap = (d_0 + d_1) + d_2; nxt = s.clone; nxt[1...-1, ...] = c + lap * C
    Boundary cells are held fixed; each operator is a separately rounded array op."""
    T = state.dtype.type
    C = T(coeff)
    two = T(2)
    nd = state.ndim
    nxt = state.copy()
    if any(s < 3 for s in state.shape):
        return nxt
    inner = tuple(slice(1, -1) for _ in range(nd))
    c = state[inner]
    with np.errstate(all="ignore"):
        two_c = (two * c).astype(state.dtype)
        lap = None
        for k in range(nd):
            lo = tuple(slice(0, -2) if j == k else slice(1, -1) for j in range(nd))
            hi = tuple(slice(2, None) if j == k else slice(1, -1) for j in range(nd))
            dk = (state[lo] - two_c).astype(state.dtype)
            dk = (dk + state[hi]).astype(state.dtype)
            lap = dk if lap is None else (lap + dk).astype(state.dtype)
        nxt[inner] = (c + (lap * C).astype(state.dtype)).astype(state.dtype)
    return nxt


# --------------------------------------------------------------------------
# Vectorised twins of gather / scatter for big parity cases: enumerate buffer
# offsets of a region or a view chain with numpy instead of a Python loop.
# Cross-checked against the per-element restatements in tests/.
# --------------------------------------------------------------------------
def region_buffer_offsets(region: IndexRegion, shape: Sequence[int]) -> np.ndarray:
    strides = axis_strides(shape)
    if any(s == 0 for s in region.step):
        return np.zeros(0, dtype=np.int64)
    grids = []
    for i in range(len(shape)):
        n = region.proper_shape[i]
        grids.append((region.first[i] + region.step[i] * np.arange(n, dtype=np.int64)) * strides[i])
    total = np.zeros((), dtype=np.int64)
    for i, g in enumerate(grids):
        shp = [1] * len(grids)
        shp[i] = g.size
        total = total + g.reshape(shp)
    return total.reshape(-1)


def fetch_chunk_fast(narr: np.ndarray, region: IndexRegion) -> np.ndarray:
    off = region_buffer_offsets(region, narr.shape)
    return narr.reshape(-1)[off].reshape(region.shape)


def set_chunk_fast(narr: np.ndarray, region: IndexRegion, value) -> None:
    off = region_buffer_offsets(region, narr.shape)
    if isinstance(value, np.ndarray):
        narr.reshape(-1)[off] = value.reshape(-1)
    else:
        narr.reshape(-1)[off] = value
