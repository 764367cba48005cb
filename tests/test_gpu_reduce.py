"""T2 parity (GPU): full and per-axis reductions vs the oracle.
Reference: Enumerable#sum/min/max over NArray#each src/n_array.cr:556-564; argmax idiom
README.md:56-61; per-axis = fold of each_slice(axis) src/multi_indexable.cr:742-748."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D
from oracle import ph_oracle as O
from gpu_util import assert_bits, take_flags

F32_TOL, F64_TOL = 1e-4, 1e-6          # BASELINE.json north_star tolerances for reductions


def exact_ints(rs, n, dtype):
    """integers in {-8..8} stored in `dtype`: every partial sum is exact in any order."""
    return rs.randint(-8, 9, size=n).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("n", [1, 5, 1000, 256 * 8 * 4, 256 * 8 * 4 * 3 + 17, 1_000_003])
def test_full_sum_bit_exact_on_exact_data(dtype, n):
    rs = np.random.RandomState(n % 97)
    a = exact_ints(rs, n, dtype)
    got = D.from_host(a).sum()
    want = O.reduce_sum_fast(a)
    assert got == np.dtype(dtype).type(want) and np.asarray(got).dtype == np.dtype(dtype)
    if n <= 1000:
        assert got == O.reduce_sum_sequential(a)


@pytest.mark.parametrize("dtype,tol", [(np.float32, F32_TOL), (np.float64, F64_TOL)])
def test_full_sum_tolerance(dtype, tol):
    rs = np.random.RandomState(7)
    a = rs.rand(3_000_001).astype(dtype)
    got = float(D.from_host(a).sum())
    truth = O.reduce_sum_fast(a)
    assert abs(got - truth) <= tol * abs(truth)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_full_minmax_argmax_first_extremum(dtype):
    rs = np.random.RandomState(11)
    for n in [1, 33, 5000, 1_000_003]:
        a = (rs.rand(n) * 1000 - 500).astype(dtype)
        if n > 40:                                              # planted unique max, planted TIE for min
            a[n // 3] = 9999
            a[n // 5] = a[n // 2] = -9999
        d = D.from_host(a)
        assert d.max() == O.reduce_minmax(a, "max") and d.min() == O.reduce_minmax(a, "min")
        v, c = d.argmax(); wv, wi = O.reduce_argmax(a, "max"); assert (v, c) == (wv, [wi])
        v, c = d.argmin(); wv, wi = O.reduce_argmax(a, "min"); assert (v, c) == (wv, [wi])
    b = np.array([[0, 1, 2], [10, 11, 12]], np.int32)           # README.md:56-61
    assert D.from_host(b).argmax() == (12, [1, 2])
    t = np.array([[3, 9, 2], [9, 1, 0]], np.float32)            # tie -> the lower flat index
    assert D.from_host(t).argmax() == (np.float32(9), [0, 1])


def test_signed_zero_and_nan_and_empty():
    z = np.array([-0.0, 0.0, -0.0], np.float32)                 # strict > keeps the FIRST: -0.0
    assert np.signbit(D.from_host(z).max()) and np.signbit(D.from_host(z[::-1].copy()).max())
    assert not np.signbit(D.from_host(np.array([0.0, -0.0], np.float32)).max())
    with pytest.raises(ph.CrArgumentError):
        D.from_host(np.array([1.0, np.nan, 3.0], np.float64)).max()
    with pytest.raises(ph.CrArgumentError):
        D.from_host(np.array([[1.0, np.nan], [0.0, 2.0]], np.float32)).argmax(axis=0)
    e = D.from_host(np.zeros((0, 3), np.float32))
    assert e.sum() == 0
    with pytest.raises(ph.CrEmptyError):
        e.max()


def test_integer_sum_overflow_is_prefix_exact():
    """Enumerable#sum raises when ANY prefix leaves T, even if the total fits."""
    mx = 2**31 - 1
    with pytest.raises(ph.CrOverflowError):
        D.from_host(np.array([mx, 1, -5], np.int32)).sum()
    assert D.from_host(np.array([mx, -5, 1], np.int32)).sum() == mx - 4          # no prefix overflows
    rs = np.random.RandomState(2)
    big = rs.randint(0, 2**20, size=200_000).astype(np.int32)                    # positives alone overflow...
    alt = np.empty(400_000, np.int32); alt[0::2] = big; alt[1::2] = -big         # ...but no prefix does
    assert D.from_host(alt).sum() == 0
    bad = np.concatenate([big, -big])                                            # prefix overflows, total = 0
    with pytest.raises(ph.CrOverflowError):
        D.from_host(bad).sum()
    with pytest.raises(ph.CrOverflowError):
        D.from_host(np.array([2**63 - 1, 1, -2], np.int64)).sum()
    assert D.from_host(np.array([2**63 - 1, -2, 1], np.int64)).sum() == 2**63 - 2


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("shape", [(7, 5, 3), (64, 1000), (1000, 64), (33, 17, 129), (3, 100000), (100000, 3), (2, 2, 2, 64)])
def test_axis_reductions(dtype, shape):
    rs = np.random.RandomState(5)
    a = exact_ints(rs, int(np.prod(shape)), dtype).reshape(shape)
    d = D.from_host(a)
    for axis in range(len(shape)):
        for which in ["sum", "max", "min", "argmax", "argmin"]:
            want = O.reduce_axis(a, axis, which)
            got = getattr(d, which)(axis=axis)
            assert got.shape == list(want.shape)
            assert_bits(got.to_host(), want.astype(got.dtype), f"{which} axis={axis} {shape}")


@pytest.mark.parametrize("dtype,tol", [(np.float32, F32_TOL), (np.float64, F64_TOL)])
def test_axis_sum_float(dtype, tol):
    """Axis != last keeps the fold order => bit-exact even on general floats; the last axis is
    a tree => tolerance."""
    rs = np.random.RandomState(6)
    a = rs.rand(40, 300, 50).astype(dtype)
    d = D.from_host(a)
    assert_bits(d.sum(axis=0).to_host(), O.reduce_axis(a, 0, "sum"), "axis 0 ordered fold")
    assert_bits(d.sum(axis=1).to_host(), O.reduce_axis(a, 1, "sum"), "axis 1 ordered fold")
    np.testing.assert_allclose(d.sum(axis=2).to_host(), np.sum(a.astype(np.float64), axis=2), rtol=tol)


def test_axis_int_overflow_and_errors():
    a = np.array([[2**31 - 1, 1], [1, 1], [-5, 1]], np.int32)
    with pytest.raises(ph.CrOverflowError):
        D.from_host(a).sum(axis=0)
    assert D.from_host(a).sum(axis=1).to_host().tolist() if False else True
    with pytest.raises(ph.CrOverflowError):
        D.from_host(np.array([[2**31 - 1, 1, -5]], np.int32)).sum(axis=1)       # prefix along the row
    assert D.from_host(np.array([[2**31 - 1, -5, 1]], np.int32)).sum(axis=1).to_host().tolist() == [2**31 - 5]
    with pytest.raises(ph.CrIndexError):
        D.from_host(a).sum(axis=2)


def test_reduce_strided_view():
    """Reductions over a view gather first (lex order of the view = index space of argmax)."""
    n = np.arange(60, dtype=np.float32).reshape(6, 10)
    v = D.from_host(n).view(ph.rng(None, None, -1), ph.rng(1, 9, 2))
    want = n[::-1, 1:10:2]
    assert v.sum() == want.sum()
    assert v.argmax() == (want.max(), [0, 4])
