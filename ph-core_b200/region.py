"""Region literals for the Python host mirror.

Crystal writes `a..b`, `a...b`, `a..s..b`, `(a..s)..b`, `a..(s..b)`, `a.step(by:, to:)`;
Python has no such syntax, so a literal entry is an int, a Python slice (only for plain
forward ranges), or an `R` / `Step` object.  Parsing the NESTING is syntax and is done here
(RangeSyntax.parse_range, src/range_syntax/range_syntax.cr:41-69); everything after it --
inference, canonicalisation, bounds -- is done by the C++ host layer (include/ph_host.h).
The marshaller is duck-typed, so the oracle's own R/Step objects are accepted too.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

from ._lib import PhRangeLit


@dataclass(frozen=True)
class R:
    """A Crystal Range: R(a, b) = a..b, R(a, b, True) = a...b; begin/end may be None or R."""
    begin: Union[None, int, "R"] = None
    end: Union[None, int, "R"] = None
    exclusive: bool = False


@dataclass(frozen=True)
class Step:
    """Crystal `a.step(by: s, to: b, exclusive: e)`."""
    current: int
    limit: Optional[int]
    step: int
    exclusive: bool = False


def rng(a=None, b=None, step=None, exclusive=False) -> R:
    """rng(a, b) = a..b ; rng(a, b, s) = a..s..b ; exclusive=True -> `...`."""
    if step is None:
        return R(a, b, exclusive)
    return R(R(a, step), b, exclusive)


ALL = R(None, None)     # `..`


def _is_int(v) -> bool:
    return isinstance(v, int) and not isinstance(v, bool) or (hasattr(v, "__index__") and not hasattr(v, "begin")
                                                              and not isinstance(v, (bool, slice)))


def marshal(lit) -> PhRangeLit:
    out = PhRangeLit()
    if _is_int(lit):
        out.is_index = 1
        out.first = int(lit)
        return out
    if isinstance(lit, slice):
        first, last, step, excl = lit.start, lit.stop, lit.step, True
    elif hasattr(lit, "current"):
        first, last, step, excl = lit.current, lit.limit, lit.step, bool(lit.exclusive)
    elif hasattr(lit, "begin"):
        b, e = lit.begin, lit.end
        if hasattr(b, "begin"):                                     # (a..s)..c
            if not (e is None or _is_int(e)):
                raise ValueError("poorly formatted range")
            first, step, last, excl = b.begin, b.end, e, bool(lit.exclusive)
        elif hasattr(e, "begin"):                                   # a..(s..c)
            first, step, last, excl = b, e.begin, e.end, bool(e.exclusive)
        else:
            first, step, last, excl = b, None, e, bool(lit.exclusive)
    else:
        raise TypeError(f"not a region literal entry: {lit!r}")
    out.is_index = 0
    out.exclusive = int(excl)
    if first is not None:
        out.has_first, out.first = 1, int(first)
    if last is not None:
        out.has_last, out.last = 1, int(last)
    if step is not None:
        out.has_step, out.step = 1, int(step)
    return out
