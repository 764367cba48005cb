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


def _elements(host: np.ndarray) -> list:
    flat = host.reshape(-1)
    if host.dtype == np.bool_:
        return [bool(v) for v in flat]
    if host.dtype.kind in "iu":
        return [int(v) for v in flat]
    return [float(v) for v in flat]


def to_json(arr: DeviceNArray) -> str:
    """NArray#to_json (src/n_array.cr:807-818): compact separators, flat lex order."""
    return json.dumps({"shape": [int(s) for s in arr.shape], "elements": _elements(arr.to_host())},
                      separators=(",", ":"))


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
    return DeviceNArray.from_host(np.array(elements, dtype=dtype).reshape(shape))


def to_yaml(arr: DeviceNArray) -> str:
    """NArray#to_yaml (src/n_array.cr:853-869): flow sequences, document start marker."""
    fmt = lambda v: ("true" if v else "false") if isinstance(v, bool) else repr(v) if isinstance(v, float) else str(v)
    shape = ", ".join(str(int(s)) for s in arr.shape)
    elems = ", ".join(fmt(v) for v in _elements(arr.to_host()))
    return f"---\nshape: [{shape}]\nelements: [{elems}]\n"


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
