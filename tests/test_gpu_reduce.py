"""T2 parity (GPU): full and per-axis reductions vs the oracle.
Reference: Enumerable#sum/min/max over NArray#each src/n_array.cr:556-564; argmax idiom
README.md:56-61; per-axis = fold of each_slice(axis) src/multi_indexable.cr:742-748."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, rng
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


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_reductions_read_strided_views_in_place(dtype):
    """a row-strided slice, a column block and a reversed view are reduced where they lie --
    ONE launch, no gather into a temporary (the launch counter proves it) -- with the same results as the
    gathered form: first extremum on the view's own lexicographic index, ordered folds off the last axis."""
    lib = ph.load()
    rs = np.random.RandomState(12)
    n = (rs.randint(-8, 9, size=(64, 8192)) if np.dtype(dtype).kind == "i" else rs.randint(-8, 9, size=(64, 8192))).astype(dtype)
    n[10, 5000] = n[40, 17] = 50                     # tie for the max: the lower lex index OF THE VIEW wins
    n[33, 1] = -50
    d = D.from_host(n)

    def launches(fn):
        before = lib.ph_launch_count()
        out = fn()
        return out, lib.ph_launch_count() - before

    cases = [("rows 0,2,4..", (rng(0, None, 2), ph.ALL), n[0::2, :]),
             ("column block", (ph.ALL, rng(1024, 5119)), n[:, 1024:5120]),
             ("rows reversed", (rng(None, None, -1), ph.ALL), n[::-1, :])]
    for name, lit, want in cases:
        v = d.view(*lit)
        flat = want.reshape(-1)
        got, k = launches(lambda: v.sum())
        assert got == dtype(flat.astype(np.float64).sum()) and k == 1, (name, got, k)
        got, k = launches(lambda: v.argmax())
        assert got == (dtype(50), list(np.unravel_index(int(np.argmax(flat)), want.shape))) and k == 1, (name, got, k)
        got, k = launches(lambda: v.argmin())
        assert got[0] == flat.min() and got[1] == list(np.unravel_index(int(np.argmin(flat)), want.shape)) and k == 1, (name, got)
    # per axis: strip kernel over a row-strided / reversed-rows view, row kernels over reversed rows
    views = [("rows 0,2,4..", d.view(rng(0, None, 2), ph.ALL), n[0::2, :]),
             ("both reversed", d.view().reverse(), n[::-1, ::-1]),
             ("column block", d.view(ph.ALL, rng(1024, 5119)), n[:, 1024:5120]),
             ("rows reversed", d.view(rng(None, None, -1), ph.ALL), n[::-1, :])]
    for name, v, want in views:
        want = np.ascontiguousarray(want)
        for axis in (0, 1):
            for which in ("sum", "max", "min", "argmax", "argmin"):
                got, k = launches(lambda: getattr(v, which)(axis=axis))
                assert k == 1, (name, axis, which, k)                   # no gather launch in front of the reduction
                assert_bits(got.to_host(), O.reduce_axis(want, axis, which), f"{name} {which} axis={axis}")
    # short reversed rows (the register-resident row kernel) with a zero whose sign depends on the order
    z = np.zeros((7, 64), dtype if np.dtype(dtype).kind == "f" else np.float32)
    z[:, 3] = -0.0
    z[2, 9] = -0.0
    zv = D.from_host(z).view().reverse()
    zw = np.ascontiguousarray(z[::-1, ::-1])
    assert_bits(zv.max(axis=1).to_host(), O.reduce_axis(zw, 1, "max"), "zero sign, reversed rows")
    assert_bits(zv.argmin(axis=1).to_host(), O.reduce_axis(zw, 1, "argmin"), "argmin, reversed rows")
    # a view the kernels cannot address in place (column step 2) still works through the gather
    g = d.view(ph.ALL, rng(0, None, 2))
    assert g.sum() == dtype(n[:, ::2].astype(np.float64).sum())
    assert_bits(g.max(axis=0).to_host(), O.reduce_axis(np.ascontiguousarray(n[:, ::2]), 0, "max"), "gathered view")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_few_column_axis_folds_through_the_staged_kernel(dtype):
    """Per-axis folds with FEW columns (>= 148 strips of 32 columns, < 1.5 MB of columns) run in
    axis_strip_staged_kernel: one warp per strip, cp.async ring, the same k-ordered fold -- float sums stay
    bit-identical to the sequential each_slice fold (src/multi_indexable.cr:742-748), first extremum wins."""
    rs = np.random.RandomState(31)
    # the last three shapes have >= 148 strips of 64 columns: 4-byte elements take the two-columns-per-lane form
    # (256-byte strip rows), incl. a ragged last strip (4900 % 64 = 36) and a strip count that is not a multiple of 2
    for shape, axis in [((200, 4800), 0), ((3, 100, 2000), 1), ((70, 5000), 0), ((2, 77, 3, 800), 1),
                        ((100, 9604), 0), ((3, 70, 3200), 1), ((2, 66, 4900), 1)]:
        if np.dtype(dtype).kind == "f":
            a = (rs.rand(*shape) * 2 - 1).astype(dtype)          # general data: only the exact k order is bit-identical
        else:
            a = rs.randint(-1000, 1000, size=shape).astype(dtype)
        idx = tuple(rs.randint(0, s) for s in shape)
        a[idx] = 5000
        lo = list(idx); lo[axis] = (idx[axis] + 7) % shape[axis]; a[tuple(lo)] = 5000      # a tie along the folded axis
        d = D.from_host(a)
        for which in ("sum", "max", "min", "argmax", "argmin"):
            assert_bits(getattr(d, which)(axis=axis).to_host(), O.reduce_axis(a, axis, which), f"{shape} {which} axis={axis}")
        # a row-reversed view of the same array (negative k stride): read in place
        if axis == 0 and len(shape) == 2:
            v = d.view(rng(None, None, -1), ph.ALL)
            w = np.ascontiguousarray(a[::-1])
            for which in ("sum", "argmax", "min"):
                assert_bits(getattr(v, which)(axis=0).to_host(), O.reduce_axis(w, 0, which), f"reversed {which}")
            # both axes reversed: the strip is copied in memory order and the lanes read it backwards
            v = d.view().reverse()
            w = np.ascontiguousarray(a[::-1, ::-1])
            for which in ("sum", "argmin", "max"):
                assert_bits(getattr(v, which)(axis=0).to_host(), O.reduce_axis(w, 0, which), f"{shape} fully reversed {which}")
    if np.dtype(dtype).kind == "f":
        z = np.zeros((80, 4800), dtype)
        z[3, :] = -0.0
        z[40, 17] = np.nan
        dz = D.from_host(z)
        with pytest.raises(ph.CrArgumentError):
            dz.max(axis=0).to_host()
        z[40, 17] = 0.0
        assert_bits(D.from_host(z).max(axis=0).to_host(), O.reduce_axis(z, 0, "max"), "first zero keeps its sign")
    else:
        big = np.full((80, 4800), np.iinfo(dtype).max // 40, dtype)
        with pytest.raises(ph.CrOverflowError):
            D.from_host(big).sum(axis=0).to_host()
    assert take_flags() == set()
