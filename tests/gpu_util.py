"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle through the C-ABI)."""
import ctypes as C

import numpy as np

import ph_core_b200 as ph
from ph_core_b200 import _lib
from ph_core_b200.narray import _Buffer


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-exact comparison (NaN payloads and signed zeros included)."""
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return a.tobytes() == b.tobytes()


def _canon_nan(x: np.ndarray) -> np.ndarray:
    """IEEE-754 leaves the sign/payload of a generated NaN unspecified: x86 (where the
    reference runs) yields the negative "real indefinite" 0xFFC00000 and propagates operand
    payloads, sm_100 yields 0x7FFFFFFF.  NaN-ness must match exactly; the payload cannot."""
    if x.dtype.kind != "f":
        return x
    y = x.copy()
    y[np.isnan(y)] = np.nan
    return y


def assert_bits(got: np.ndarray, want: np.ndarray, what=""):
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    got, want = _canon_nan(got), _canon_nan(want)
    if got.tobytes() != want.tobytes():
        ga, wa = got.reshape(-1), want.reshape(-1)
        gv = ga.view(np.uint8).reshape(ga.size, -1)
        wv = wa.view(np.uint8).reshape(wa.size, -1)
        bad = np.nonzero((gv != wv).any(axis=1))[0]
        i = int(bad[0])
        raise AssertionError(f"{what}: {bad.size} of {ga.size} elements differ; first at flat {i}: "
                             f"got {ga[i]!r} want {wa[i]!r}")


def special_values(dtype, n, seed=0):
    """Random values salted with the float edge cases BASELINE.md section 4 lists."""
    rs = np.random.RandomState(seed)
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        x = (rs.rand(n) * 2 - 1).astype(dtype)
        fi = np.finfo(dtype)
        specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, fi.tiny, -fi.tiny, fi.tiny / 4, fi.max, -fi.max,
                             fi.max / 3, 1.0, -1.0, 1 + fi.eps, 3.0000002, 1e-20, 7.5, -2.5], dtype=dtype)
        idx = rs.randint(0, n, size=min(n, 4 * specials.size))
        x[idx] = specials[rs.randint(0, specials.size, size=idx.size)]
        return x
    info = np.iinfo(dtype)
    x = rs.randint(-1000, 1000, size=n).astype(dtype)
    return x


class Dev:
    """A raw device buffer holding a numpy array (for tests that drive the C-ABI directly
    with hand-built descriptors)."""

    def __init__(self, arr: np.ndarray):
        _lib.init()
        self.host = np.ascontiguousarray(arr)
        self.buf = _Buffer(max(1, self.host.nbytes))
        if self.host.nbytes:
            ph.check(_lib.load().ph_h2d(self.buf.ptr, self.host.ctypes.data, self.host.nbytes))
            ph.check(_lib.load().ph_sync())

    @property
    def ptr(self):
        return self.buf.ptr

    def read(self, shape=None, dtype=None) -> np.ndarray:
        dtype = np.dtype(dtype or self.host.dtype)
        shape = self.host.shape if shape is None else shape
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            ph.check(_lib.load().ph_d2h(out.ctypes.data, self.buf.ptr, out.nbytes))
        return out


def desc_of_view(base: np.ndarray, view: np.ndarray) -> ph.PhDesc:
    """Descriptor (element units) of a numpy view into `base`."""
    isz = base.dtype.itemsize
    off = (view.__array_interface__["data"][0] - base.__array_interface__["data"][0]) // isz
    return ph.PhDesc.make(list(view.shape), [s // isz for s in view.strides], off)


def take_flags() -> set:
    f = ph.DeviceNArray.take_flags()
    names = set()
    if f & ph.K["PH_FLAG_OVERFLOW"]:
        names.add("overflow")
    if f & ph.K["PH_FLAG_DIV0"]:
        names.add("div0")
    if f & ph.K["PH_FLAG_ARGUMENT"]:
        names.add("argument")
    if f & ph.K["PH_FLAG_NAN"]:
        names.add("nan")
    return names
