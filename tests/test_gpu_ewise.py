"""T2 parity (GPU): elementwise / broadcast / compare kernels vs the oracle, bit-exact.
Reference: def_elementwise_binary src/multi_indexable.cr:931-985, NArray#map
src/n_array.cr:589-595, patches/number.cr:6-15."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, _lib
from oracle import ph_oracle as O
from gpu_util import Dev, assert_bits, desc_of_view, special_values, take_flags

FLOAT_OPS = ["+", "-", "*", "/", "//", "%"]
INT_OPS = ["+", "-", "*", "/", "//", "%", "&+", "&-", "&*", "&", "|", "^"]
SIZES = [1, 7, 255, 4096, 8 * 256 * 2, 8 * 256 * 2 + 3, 100003]


def dev_op(op, a, b):
    f = {"+": lambda x, y: x + y, "-": lambda x, y: x - y, "*": lambda x, y: x * y, "/": lambda x, y: x / y,
         "//": lambda x, y: x // y, "%": lambda x, y: x % y, "**": lambda x, y: x ** y,
         "&": lambda x, y: x & y, "|": lambda x, y: x | y, "^": lambda x, y: x ^ y,
         "&+": lambda x, y: x.wrapping_add(y), "&-": lambda x, y: x.wrapping_sub(y),
         "&*": lambda x, y: x.wrapping_mul(y), "&**": lambda x, y: x.wrapping_pow(y)}[op]
    return f(a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("op", FLOAT_OPS)
def test_float_binary_bit_exact(dtype, op):
    for n in SIZES:
        a, b = special_values(dtype, n, 1), special_values(dtype, n, 2)
        want, wflags = O.ewise(op, a, b)
        got = dev_op(op, D.from_host(a), D.from_host(b)).to_host(False)
        assert_bits(got, want, f"{op} {np.dtype(dtype)} n={n}")
        assert take_flags() == wflags


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("op", INT_OPS)
def test_int_binary_bit_exact(dtype, op):
    rs = np.random.RandomState(5)
    for n in SIZES:
        a = rs.randint(-50000, 50000, size=n).astype(dtype)
        b = rs.randint(-50000, 50000, size=n).astype(dtype)
        if op in ("/", "//", "%"):
            b[b == 0] = 7
        want, wflags = O.ewise(op, a, b)
        got = dev_op(op, D.from_host(a), D.from_host(b)).to_host(False)
        assert_bits(got, want, f"{op} {np.dtype(dtype)} n={n}")
        assert take_flags() == wflags


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_int_flags_and_edges(dtype):
    info = np.iinfo(dtype)
    a = np.array([info.max, info.min, 5, -7, info.min, 0, 3], dtype)
    b = np.array([1, -1, 0, 2, -1, 9, -2], dtype)
    for op in ["+", "-", "*", "//", "%", "&+", "&-", "&*"]:
        want, wflags = O.ewise(op, a, b)
        got = dev_op(op, D.from_host(a), D.from_host(b)).to_host(False)
        assert take_flags() == wflags, op
        assert_bits(got, want, op)
    with pytest.raises(ph.CrDivisionByZeroError):
        _ = D.from_host(a) // D.from_host(b)
        D.raise_pending()
    with pytest.raises(ph.CrOverflowError):
        _ = D.from_host(a) + D.from_host(np.full(a.shape, info.max, dtype))
        D.raise_pending()
    # integer power: checked, wrapping, negative exponent
    base = np.array([2, 3, -2, 7, 0, 1, 10], dtype)
    exp = np.array([10, 5, 3, 2, 0, 60, 9], dtype)
    for op in ["**", "&**"]:
        want, wflags = O.ewise(op, base, exp)
        got = dev_op(op, D.from_host(base), D.from_host(exp)).to_host(False)
        assert take_flags() == wflags
        assert_bits(got, want, op)
    _ = D.from_host(base) ** D.from_host(np.full(base.shape, 70, dtype))
    assert "overflow" in take_flags()
    _ = D.from_host(base) ** D.from_host(np.full(base.shape, -1, dtype))
    assert "argument" in take_flags()
    got = (-D.from_host(a)).to_host(False)
    want, wflags = O.unary("-", a)
    assert take_flags() == wflags
    assert_bits(got, want, "neg")
    assert_bits((~D.from_host(a)).to_host(False), O.unary("~", a)[0], "not")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_scalar_both_sides(dtype):
    n = 5000
    a = special_values(dtype, n, 3)
    if np.dtype(dtype).kind == "i":
        a[a == 0] = 3
    s = np.dtype(dtype).type(3)
    d = D.from_host(a)
    for op in ["+", "-", "*", "/", "//", "%"]:
        want, _ = O.ewise(op, a, s)
        assert_bits(dev_op(op, d, s).to_host(False), want, f"a {op} s")
        want, _ = O.ewise(op, s, a)                           # scalar is the LEFT operand (number.cr:9-13)
        got = {"+": lambda: s + d, "-": lambda: s - d, "*": lambda: s * d, "/": lambda: s / d,
               "//": lambda: s // d, "%": lambda: s % d}[op]
        # numpy scalars would hijack the reflected op: use plain Python numbers
        pys = float(s) if np.dtype(dtype).kind == "f" else int(s)
        got = {"+": lambda: pys + d, "-": lambda: pys - d, "*": lambda: pys * d, "/": lambda: pys / d,
               "//": lambda: pys // d, "%": lambda: pys % d}[op]().to_host(False)
        assert_bits(got, want, f"s {op} a")
        take_flags()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_powi_and_pow(dtype):
    a = special_values(dtype, 3000, 9)
    d = D.from_host(a)
    for n in [0, 1, 2, 3, 5, 10, -1, -2, -7]:
        want, _ = O.ewise("**", a, np.int32(n))
        assert_bits((d ** n).to_host(False), want, f"powi {n}")
    # Float ** Float goes to libm pow: tolerance only
    x = (np.random.RandomState(1).rand(2000) * 4 + 0.1).astype(dtype)
    y = (np.random.RandomState(2).rand(2000) * 3 - 1.5).astype(dtype)
    got = (D.from_host(x) ** D.from_host(y)).to_host(False)
    np.testing.assert_allclose(got, np.power(x, y), rtol=1e-6 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_compare_and_unary(dtype):
    n = 70001
    a, b = special_values(dtype, n, 4), special_values(dtype, n, 5)
    b[::3] = a[::3]
    da, db = D.from_host(a), D.from_host(b)
    for op, fn in [(">", lambda: da > db), ("<", lambda: da < db), (">=", lambda: da >= db),
                   ("<=", lambda: da <= db), ("==", lambda: da.eq(db))]:
        got = fn().to_host(False)
        assert got.dtype == np.bool_
        assert_bits(got, O.compare(op, a, b), op)
    s = np.dtype(dtype).type(0)
    assert_bits((da > (0.0 if np.dtype(dtype).kind == "f" else 0)).to_host(False), O.compare(">", a, s), "> scalar")
    assert_bits(da.eq(float(a[5]) if np.dtype(dtype).kind == "f" else int(a[5])).to_host(False),
                O.compare("==", a, a[5]), "eq scalar")
    assert_bits((+da).to_host(False), a, "pos")
    if np.dtype(dtype).kind == "f":
        assert_bits((-da).to_host(False), -a, "neg")


def test_shape_errors():
    a = D.from_host(np.zeros((2, 3), np.float32))
    b = D.from_host(np.zeros((3, 2), np.float32))
    with pytest.raises(ph.ShapeError):
        a + b
    with pytest.raises(ph.DimensionError):
        a.eq(b)
    with pytest.raises(ph.DeviceBlockError):
        a.map(lambda x: x)
    with pytest.raises(ph.DeviceBlockError):
        a.buffer


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
@pytest.mark.parametrize("shape", [(37, 64), (5, 3, 24), (128, 8192), (3, 1000), (2, 2, 2, 16)])
def test_broadcast_rows(dtype, shape):
    """K2: row-vector / column-vector / scalar-axis broadcasting; oracle = tile + op."""
    rs = np.random.RandomState(11)
    a = special_values(dtype, int(np.prod(shape)), 6).reshape(shape)
    for bshape in [(1,) * (len(shape) - 1) + (shape[-1],),          # row vector
                   shape[:-1] + (1,),                                # column vector
                   (1,) * len(shape),                                # single element
                   (shape[0],) + (1,) * (len(shape) - 2) + (shape[-1],)]:
        b = special_values(dtype, int(np.prod(bshape)), 7).reshape(bshape)
        for op in ["*", "+", "-"]:
            want, wf = O.ewise_broadcast(op, a, b)
            got = D.from_host(a).broadcast_op(op, D.from_host(b)).to_host(False)
            assert_bits(got, want, f"{shape} {op} {bshape}")
            want2, _ = O.ewise_broadcast(op, b, a)
            got2 = D.from_host(b).broadcast_op(op, D.from_host(a)).to_host(False)
            assert_bits(got2, want2, f"{bshape} {op} {shape}")
            take_flags()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mul_add_two_roundings(dtype):
    """a*b+c fused == the reference's two separately rounded operators (never an FMA)."""
    shape = (257, 1024)
    a = special_values(dtype, shape[0] * shape[1], 1).reshape(shape)
    b = special_values(dtype, shape[1], 2).reshape(1, shape[1])
    c = special_values(dtype, shape[0] * shape[1], 3).reshape(shape)
    t, _ = O.ewise_broadcast("*", a, b)
    want, _ = O.ewise("+", t, c)
    da, db, dc = D.from_host(a), D.from_host(b), D.from_host(c)
    two_step = da.broadcast_op("*", db) + dc
    assert_bits(two_step.to_host(False), want, "two kernels")
    assert_bits(da.mul_add(db, dc).to_host(False), want, "fused")
    fma = (a.astype(np.longdouble) * b + c).astype(dtype)      # what an FMA would have produced
    # the data can tell the two apart: a contracted multiply-add would NOT reproduce `want`
    fin = np.isfinite(fma) & np.isfinite(want)
    assert not np.array_equal(fma[fin], want[fin])


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_strided_operands_through_cabi(dtype):
    """Operands described by arbitrary descriptors: offsets that break 32/16-byte alignment,
    negative and non-unit inner strides, permuted axes (the `any` kernel)."""
    lib = _lib.load()
    rs = np.random.RandomState(2)
    base_a = special_values(dtype, 40 * 50, 1).reshape(40, 50)
    base_b = special_values(dtype, 60 * 70, 2).reshape(60, 70)
    cases = [
        (base_a[1:31, 3:43], base_b[2:32, 1:41]),               # unaligned rows
        (base_a[::2, ::3][:15, :12], base_b[5:20, 7:19]),       # strided inner
        (base_a[::-1, ::-1][:20, :30], base_b[10:30, 20:50]),   # negative strides
        (base_a.T[:30, :25], base_b[:30, :25]),                 # permuted
        (base_a[3, 5:45], base_b[7:47, 9]),                     # 1-D row vs column
    ]
    da, db = Dev(base_a), Dev(base_b)
    for va, vb in cases:
        want, _ = O.ewise("-", np.ascontiguousarray(va), np.ascontiguousarray(vb))
        out = Dev(np.zeros(va.shape, dtype))
        desc_a, desc_b = desc_of_view(base_a, va), desc_of_view(base_b, vb)
        desc_o = ph.PhDesc.contiguous(va.shape)
        ph.check(lib.ph_ewise_binary(ph.K["PH_SUB"], ph.narray.dtype_code(dtype), da.ptr, C.byref(desc_a),
                                     db.ptr, C.byref(desc_b), out.ptr, C.byref(desc_o)))
        assert_bits(out.read(), want, f"strided {va.shape} {va.strides}")
        take_flags()
    # strided OUTPUT: write into a sub-block of a larger zeroed array
    big = np.zeros((50, 64), dtype)
    dout = Dev(big)
    va, vb = base_a[:30, :40], base_b[:30, :40]
    target = big[7:37, 8:48]
    ph.check(lib.ph_ewise_binary(ph.K["PH_ADD"], ph.narray.dtype_code(dtype), da.ptr, C.byref(desc_of_view(base_a, va)),
                                 db.ptr, C.byref(desc_of_view(base_b, vb)), dout.ptr, C.byref(desc_of_view(big, target))))
    want = big.copy()
    want[7:37, 8:48] = O.ewise("+", np.ascontiguousarray(va), np.ascontiguousarray(vb))[0]
    assert_bits(dout.read(), want, "strided output")
    take_flags()


def test_empty_and_scalar_arrays():
    for shape in [(0,), (5, 0, 2), (1,), (1, 1, 1)]:
        a = np.ones(shape, np.float32)
        got = (D.from_host(a) + D.from_host(a)).to_host(False)
        assert got.shape == tuple(shape)
        if a.size:
            assert (got == 2).all()


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.uint16, np.uint32, np.uint64])
def test_small_and_unsigned_ints(dtype):
    """Every Crystal primitive integer: checked / wrapping / floored semantics, flags, compare."""
    info = np.iinfo(dtype)
    rs = np.random.RandomState(13)
    for n in [1, 100, 4099, 70001]:
        hi = min(int(info.max), 2**31 - 1)
        lo = max(int(info.min), -2**31)
        a = rs.randint(lo, hi, size=n, dtype=np.int64).astype(dtype)
        b = rs.randint(lo, hi, size=n, dtype=np.int64).astype(dtype)
        b[b == 0] = 1
        da, db = D.from_host(a), D.from_host(b)
        for op in ["+", "-", "*", "//", "%", "&+", "&-", "&*", "&", "|", "^", "/"]:
            want, wflags = O.ewise(op, a, b)
            got = dev_op(op, da, db).to_host(False)
            assert take_flags() == wflags, (op, np.dtype(dtype))
            assert_bits(got, want, f"{op} {np.dtype(dtype)} n={n}")
        small = (a % 7).astype(dtype)
        e = (rs.randint(0, 5, size=n)).astype(dtype)
        for op in ["**", "&**"]:
            want, wflags = O.ewise(op, small, e)
            got = dev_op(op, D.from_host(small), D.from_host(e)).to_host(False)
            assert take_flags() == wflags
            assert_bits(got, want, f"{op} {np.dtype(dtype)}")
        assert_bits((da > db).to_host(False), O.compare(">", a, b), "cmp")
        assert_bits(da.eq(db).to_host(False), O.compare("==", a, b), "eq")
        s = dtype(3)
        want, wf = O.ewise("*", a, s); got = (da * 3).to_host(False); assert take_flags() == wf; assert_bits(got, want, "scalar *")
        want, wf = O.ewise("-", s, a); got = (3 - da).to_host(False); assert take_flags() == wf; assert_bits(got, want, "scalar - left")
    with pytest.raises(ph.CrDivisionByZeroError):
        _ = D.from_host(np.array([5], dtype)) // D.from_host(np.array([0], dtype))
        D.raise_pending()
    with pytest.raises(ph.CrOverflowError):
        _ = D.from_host(np.array([info.max], dtype)) + D.from_host(np.array([1], dtype))
        D.raise_pending()


def test_reads_are_raise_points():
    """a data-dependent error surfaces at the first synchronising read without any manual
    raise_pending -- `(a + max).to_host()` raises OverflowError, `(a // 0).get(0)` DivisionByZeroError
    (the reference raises at the operator: Int32#+ / Int32#//, src/multi_indexable.cr:942-944)."""
    info = np.iinfo(np.int32)
    a = D.from_host(np.array([1, 2, info.max], np.int32))
    with pytest.raises(ph.CrOverflowError):
        (a + D.from_host(np.full(3, info.max, np.int32))).to_host()
    assert take_flags() == set()                              # the raise consumed the flag
    with pytest.raises(ph.CrDivisionByZeroError):
        (a // 0).get(0)
    with pytest.raises(ph.CrOverflowError):
        _ = a * info.max
        ph.narray.sync()
    assert (a + 1).to_host(False).tolist() == [2, 3, info.min] and take_flags() == {"overflow"}
    ok = (a - 1).to_host()                                    # nothing pending: plain read
    assert ok.tolist() == [0, 1, info.max - 1]


def test_operand_dtypes_are_checked_before_launch():
    """kernels are typed by one element type; a mismatched second operand or an
    unrepresentable scalar is a TypeError on the host, never a wrong-size read on the device."""
    f64 = D.from_host(np.arange(6, dtype=np.float64))
    f32 = D.from_host(np.arange(6, dtype=np.float32))
    i32 = D.from_host(np.arange(6, dtype=np.int32))
    for bad in (lambda: f64 > f32, lambda: f64.eq(f32), lambda: f64.mul_add(f32, f64), lambda: f64.mul_add(f64, f32),
                lambda: f64.broadcast_op("+", f32), lambda: f64.set_mask(f64 > f64, f32), lambda: f64 + f32,
                lambda: i32 < 2.5, lambda: i32 * 2.5, lambda: i32 + (1 << 40), lambda: i32.set_mask(i32 > 1, 0.5)):
        with pytest.raises(TypeError):
            bad()
    assert (i32 < 3.0).to_host().tolist() == [True, True, True, False, False, False]      # 3.0 IS an Int32
    assert (i32 * np.int64(2)).to_host().tolist() == [0, 2, 4, 6, 8, 10]
    assert (f32 * 0.1).to_host().tobytes() == (np.arange(6, dtype=np.float32) * np.float32(0.1)).tobytes()
