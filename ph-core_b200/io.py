"""Host <-> device I/O formats either side of the hot path.

The reference serialises an NArray as `{"shape": [...], "elements": [...]}` with the elements in
flat lexicographic order (src/n_array.cr:807-912; goldens spec/n_array_spec.cr:520-558).  A
device array is brought to the host explicitly and then written in exactly that form, so files
are interchangeable with ph-core's own.  For arrays too large for text there is a raw binary
dump: a one-line JSON header (shape, dtype, byte order) followed by the row-major bytes.
"""
from __future__ import annotations

import json
from typing import Sequence

import numpy as np

from .narray import DeviceNArray, ShapeError

_MAGIC = b"PHNARR1\n"


def format_float(v) -> str:
    """Crystal's Float#to_s for a Float32 / Float64 element, which is what `NArray#to_json` / `#to_yaml`
    write (JSON::Builder#number / YAML scalar -> Float::Printer, Crystal 1.0.0 stdlib -- not vendored in
    the reference, restated: the SHORTEST digits that round-trip in the element's own width, positional
    for decimal points in [-3, 15], otherwise `d.ddde+X` with at least one fraction digit and an
    unpadded exponent: 0.1, 100000000000000.0, 1.0e+15, 1.0e-5).  include/ph_narray_io.hpp's
    format_element applies the same rule, so the two hosts write identical text.  NaN and the
    infinities are not JSON numbers: both hosts raise, like JSON::Builder does."""
    v = v if isinstance(v, np.floating) else np.float64(v)
    if not np.isfinite(v):
        raise ValueError("NaN and Infinity cannot be written as JSON / YAML numbers")
    sci = np.format_float_scientific(v, unique=True, trim="-")           # shortest round trip in v's own width
    mant, exp10 = sci.split("e")
    sign = "-" if mant.startswith("-") else ""
    digits = mant.lstrip("-").replace(".", "")
    if digits.strip("0") == "":
        return sign + "0.0"
    digits = digits.rstrip("0") or "0"
    point = int(exp10) + 1                                               # value = 0.DIGITS x 10^point
    n = len(digits)
    if point > 15 or point < -3:                                         # exponent form
        frac = digits[1:] or "0"
        e = point - 1
        return f"{sign}{digits[0]}.{frac}e{'+' if e > 0 else ''}{e}"
    if point <= 0:
        return f"{sign}0.{'0' * (-point)}{digits}"
    if point >= n:
        return f"{sign}{digits}{'0' * (point - n)}.0"
    return f"{sign}{digits[:point]}.{digits[point:]}"


def _element_text(host: np.ndarray) -> list:
    flat = host.reshape(-1)
    if host.dtype == np.bool_:
        return ["true" if v else "false" for v in flat]
    if host.dtype.kind in "iu":
        return [str(int(v)) for v in flat]
    return [format_float(v) for v in flat]


def host_to_json(host: np.ndarray) -> str:
    """The text half of `to_json` on a host array (what the C++ layer's IO::host::to_json writes)."""
    shape = ",".join(str(int(s)) for s in host.shape)
    return '{"shape":[' + shape + '],"elements":[' + ",".join(_element_text(host)) + "]}"


def host_to_yaml(host: np.ndarray) -> str:
    shape = ", ".join(str(int(s)) for s in host.shape)
    return f"---\nshape: [{shape}]\nelements: [{', '.join(_element_text(host))}]\n"


def to_json(arr: DeviceNArray) -> str:
    """NArray#to_json (src/n_array.cr:807-818): compact separators, flat lex order."""
    return host_to_json(arr.to_host())


def _typed_elements(elements, dtype: np.dtype, what: str) -> np.ndarray:
    """Elements of a parsed document as `dtype`, WITHOUT numpy's silent coercions: the reference's
    `T.new(pull)` (src/n_array.cr:838-841) raises when a JSON value is not a T."""
    dtype = np.dtype(dtype)
    for v in elements:
        if dtype == np.bool_:
            ok = isinstance(v, bool)
        elif dtype.kind in "iu":
            ok = isinstance(v, int) and not isinstance(v, bool)
        else:
            ok = isinstance(v, (int, float)) and not isinstance(v, bool)
        if not ok:
            raise ValueError(f"Could not read NArray from {what}: element {v!r} is not a {dtype}")
    if dtype.kind in "iu" and len(elements):
        info = np.iinfo(dtype)
        if min(elements) < info.min or max(elements) > info.max:
            raise ValueError(f"Could not read NArray from {what}: an element does not fit {dtype}")
    return np.array(elements, dtype=dtype)


def from_json(text: str, dtype) -> DeviceNArray:
    """NArray(T).from_json (src/n_array.cr:820-851): both keys required, element count must
    match the shape (the reference raises JSON::Error; here ValueError / ShapeError)."""
    obj = json.loads(text)
    if "shape" not in obj or "elements" not in obj:
        raise ValueError("Could not read NArray from JSON: 'shape' and/or 'elements' were missing.")
    shape, elements = [int(s) for s in obj["shape"]], obj["elements"]
    size = int(np.prod(shape, dtype=np.int64)) if shape else 0
    if size != len(elements):
        raise ShapeError(f"Could not read NArray from JSON: wrong number of elements for shape {shape}")
    return DeviceNArray.from_host(_typed_elements(elements, dtype, "JSON").reshape(shape))


def to_yaml(arr: DeviceNArray) -> str:
    """NArray#to_yaml (src/n_array.cr:853-869): flow sequences, document start marker."""
    return host_to_yaml(arr.to_host())


def from_yaml(text: str, dtype) -> DeviceNArray:
    """NArray(T).from_yaml (src/n_array.cr:871-912)."""
    import yaml
    obj = yaml.safe_load(text)
    if not isinstance(obj, dict) or "shape" not in obj or "elements" not in obj:
        raise ValueError("Could not read NArray from YAML: 'shape' and/or 'elements' were missing.")
    return from_json(json.dumps(obj), dtype)


def write_dump(host: np.ndarray, path: str) -> None:
    """Host half of `dump`: magic, JSON header line (shape, numpy dtype string), raw row-major bytes.  The C++
    host layer reads and writes the same file (include/ph_narray_io.hpp)."""
    host = np.ascontiguousarray(host)
    header = json.dumps({"shape": [int(s) for s in host.shape], "dtype": host.dtype.str}).encode() + b"\n"
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(header)
        if host.size:
            f.write(memoryview(host).cast("B"))                         # no second host copy


def read_dump(path: str) -> np.ndarray:
    """Host half of `load`."""
    with open(path, "rb") as f:
        if f.read(len(_MAGIC)) != _MAGIC:
            raise ValueError(f"{path} is not a ph-core binary dump")
        header = json.loads(f.readline())
        dtype = np.dtype(header["dtype"])
        shape: Sequence[int] = header["shape"]
        data = np.fromfile(f, dtype=dtype)                              # straight into the array's own buffer
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 0
    if data.size != size:
        raise ShapeError(f"binary dump holds {data.size} elements, the header's shape {shape} needs {size}")
    return data.reshape(shape)


def dump(arr: DeviceNArray, path: str) -> None:
    """Binary checkpoint of a device array (explicit D2H, then `write_dump`)."""
    write_dump(arr.to_host(), path)


def load(path: str) -> DeviceNArray:
    return DeviceNArray.from_host(read_dump(path))
