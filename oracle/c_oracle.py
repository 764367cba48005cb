"""ctypes loader for oracle/libph_oracle.so (the C restatement).  Test infrastructure:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libph_oracle.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib = C.CDLL(_PATH)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _shape(shape):
    return (C.c_int64 * len(shape))(*[int(s) for s in shape])


def use_all_cores() -> int:
    """Use every host core the process may run on (overrides torchrun's OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    load().oracle_set_threads(C.c_int(n))
    return num_threads()


def num_threads() -> int:
    return int(load().oracle_num_threads())


def ref_map_with(op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Reference-structured `a op b` (op: 0 + 1 - 2 * 3 /), one thread."""
    assert a.shape == b.shape and a.dtype == b.dtype and a.flags.c_contiguous and b.flags.c_contiguous
    out = np.empty_like(a)
    fn = {np.dtype(np.float32): load().ref_map_with_f32, np.dtype(np.float64): load().ref_map_with_f64}[a.dtype]
    fn(C.c_int(op), _p(a), _p(b), _p(out), _shape(a.shape), C.c_int(a.ndim))
    return out


def ref_tile(src: np.ndarray, counts) -> np.ndarray:
    out = np.empty([s * c for s, c in zip(src.shape, counts)], dtype=src.dtype)
    fn = {np.dtype(np.float32): load().ref_tile_f32, np.dtype(np.float64): load().ref_tile_f64}[src.dtype]
    fn(_p(src), _shape(src.shape), _shape(counts), C.c_int(src.ndim), _p(out))
    return out


def ref_mul_rowvec_add_f32(a, b, c) -> np.ndarray:
    out = np.empty_like(a)
    load().ref_mul_rowvec_add_f32(_p(a), _p(b), _p(c), _p(out), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]))
    return out


def flat_mul_rowvec_add_f32(a, b, c, out=None, tmp=None) -> np.ndarray:
    out = np.empty_like(a) if out is None else out
    tmp = np.empty_like(a) if tmp is None else tmp
    load().flat_mul_rowvec_add_f32(_p(a), _p(b), _p(c), _p(out), _p(tmp), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]))
    return out


def flat_binary(op: int, a, b, out=None) -> np.ndarray:
    out = np.empty_like(a) if out is None else out
    fn = {np.dtype(np.float32): load().flat_binary_f32, np.dtype(np.float64): load().flat_binary_f64}[a.dtype]
    fn(C.c_int(op), _p(a), _p(b), _p(out), C.c_int64(a.size))
    return out


def ref_heat_step_3d_f32(s: np.ndarray, coeff) -> np.ndarray:
    """One 3-D heat step through the reference's operator structure (slices fetched with the
    lexicographic iterator, one materialised array per operator), one thread."""
    assert s.dtype == np.float32 and s.ndim == 3 and s.flags.c_contiguous
    out = np.empty_like(s)
    load().ref_heat_step_3d_f32(_p(s), _p(out), _shape(s.shape), C.c_float(float(coeff)))
    return out


def flat_heat_step_3d_f32(s: np.ndarray, coeff, out=None) -> np.ndarray:
    assert s.dtype == np.float32 and s.ndim == 3 and s.flags.c_contiguous
    out = np.empty_like(s) if out is None else out
    load().flat_heat_step_3d_f32(_p(s), _p(out), _shape(s.shape), C.c_float(float(coeff)))
    return out


def ref_sum_f32(x: np.ndarray) -> np.float32:
    """Enumerable#sum: sequential left fold in f32, one thread."""
    lib = load()
    lib.ref_sum_f32.restype = C.c_float
    return np.float32(lib.ref_sum_f32(_p(x), C.c_int64(x.size)))


def flat_sum_f32(x: np.ndarray) -> float:
    lib = load()
    lib.flat_sum_f32.restype = C.c_double
    return float(lib.flat_sum_f32(_p(x), C.c_int64(x.size)))
