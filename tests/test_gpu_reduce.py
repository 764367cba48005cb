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


@pytest.mark.parametrize("K", [8, 64, 1000, 2048])
def test_last_axis_register_kernel_zero_signs_ties_nan(K):
    """Rows that fit one warp's registers (axis_rowreg_kernel): rows sharing a warp disagree on
    whether the extremum is a zero (first zero's sign must survive), ties keep the first index,
    a NaN anywhere raises."""
    rs = np.random.RandomState(K)
    a = -(rs.rand(6, K).astype(np.float32) + 1)                    # all negative
    a[0, K // 2] = -0.0; a[0, K - 1] = 0.0                          # max = zero, the FIRST one is -0.0
    a[2, 3] = 0.0; a[2, 5] = -0.0                                   # first zero is +0.0
    a[4, 1] = a[4, K - 2] = -0.5; a[4, 0] = -0.75                   # tie for the max on a non-zero value
    a[4, 2:K - 2] = -3.0
    d = D.from_host(a)
    got = d.max(axis=1).to_host()
    assert np.signbit(got[0]) and not np.signbit(got[2]) and got[0] == 0 and got[2] == 0
    assert_bits(got[[1, 3, 5]], a[[1, 3, 5]].max(axis=1), "plain rows")
    arg = d.argmax(axis=1).to_host()
    assert arg[0] == K // 2 and arg[2] == 3 and arg[4] == 1
    assert arg.tolist() == [int(O.reduce_argmax(r, "max")[1]) for r in a]
    mn = d.argmin(axis=1).to_host()
    assert mn.tolist() == [int(O.reduce_argmax(r, "min")[1]) for r in a]
    np.testing.assert_allclose(d.sum(axis=1).to_host(), a.astype(np.float64).sum(axis=1), rtol=F32_TOL)
    a[3, K - 1] = np.nan
    with pytest.raises(ph.CrArgumentError):
        D.from_host(a).max(axis=1)


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64])
def test_remaining_crystal_integer_types(dtype):
    """Every Crystal primitive integer reduces on the device: sum (overflow-checked in T: raises when
    a prefix leaves T), min / max / arg* (first extremum), full and per axis."""
    rs = np.random.RandomState(31)
    info = np.iinfo(dtype)
    lo = -3 if info.min < 0 else 0
    a = rs.randint(lo, 4, size=(6, 40, 24)).astype(dtype)          # |sum| stays far inside Int8 per row
    a[2, 7, 5] = a[4, 1, 2] = 50                                   # tie for the max
    d = D.from_host(a)
    if info.max > a.astype(np.int64).sum():
        assert d.sum() == dtype(int(a.astype(np.int64).sum()))
    else:                                                          # Int8 / UInt8: the total leaves T
        with pytest.raises(ph.CrOverflowError):
            d.sum()
        take_flags()
    assert d.max() == 50 and d.min() == a.min()
    v, c = d.argmax(); assert (v, c) == (dtype(50), [2, 7, 5])
    for axis in range(3):
        for which in ["max", "min", "argmax", "argmin"]:
            assert_bits(getattr(d, which)(axis=axis).to_host(), O.reduce_axis(a, axis, which), f"{which} axis={axis}")
    row_sums = a.astype(np.int64).sum(axis=2)
    if row_sums.max() <= info.max and row_sums.min() >= info.min:
        assert_bits(d.sum(axis=2).to_host(), row_sums.astype(dtype), "sum axis=2")
    big = np.full(300, info.max // 100 + 1, dtype)                 # 300 x (max/100 + 1) overflows T
    with pytest.raises(ph.CrOverflowError):
        D.from_host(big).sum()
    take_flags()


def test_axis_int_overflow_and_errors():
    a = np.array([[2**31 - 1, 1], [1, 1], [-5, 1]], np.int32)
    with pytest.raises(ph.CrOverflowError):
        D.from_host(a).sum(axis=0)
    with pytest.raises(ph.CrOverflowError):
        D.from_host(a).sum(axis=1)                                              # row 0: Int32::MAX + 1
    assert D.from_host(a[1:]).sum(axis=1).to_host().tolist() == [2, -4]
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


def test_sharded_entry_on_one_rank_equals_the_plain_reduction():
    """ph_reduce_full_sharded with a single-rank communicator: the same kernels in record mode (result in the
    pinned host record, flags returned with it); also the empty-shard and integer-overflow decisions."""
    from ph_core_b200 import sharding as S
    S.comm_init(None)
    rs = np.random.RandomState(4)
    a = rs.randint(-8, 9, size=(37, 1000)).astype(np.float32)
    a[5, 7] = a[20, 1] = 99.0
    d = D.from_host(a)
    assert S.reduce_full_sharded(d, "sum") == d.sum() == np.float32(a.sum(dtype=np.float64))
    assert S.reduce_full_sharded(d, "max") == 99.0 and S.reduce_full_sharded(d, "min") == a.min()
    assert S.reduce_full_sharded(d, "argmax") == (np.float32(99.0), 5 * 1000 + 7)
    assert S.reduce_full_sharded(d, "argmin")[1] == int(np.argmin(a.reshape(-1)))
    e = D([0, 4], np.float32)
    assert S.reduce_full_sharded(e, "sum") == 0
    with pytest.raises(ph.CrEmptyError):
        S.reduce_full_sharded(e, "min")
    i = D.from_host(np.array([2**31 - 1, 1, -5], np.int32))
    with pytest.raises(ph.CrOverflowError):                     # a prefix leaves Int32 (exact ordered pass)
        S.reduce_full_sharded(i, "sum")
    assert S.reduce_full_sharded(D.from_host(np.array([2**31 - 1, -5, 1], np.int32)), "sum") == 2**31 - 5
    n = D.from_host(np.array([1.0, np.nan, 3.0], np.float64))
    with pytest.raises(ph.CrArgumentError):
        S.reduce_full_sharded(n, "max")
    assert take_flags() == set()
