"""T0: pin oracle/ph_oracle.py against every golden vector the reference's own
spec/ holds for the hot path.  CPU only.

Each test cites the reference spec file:line it replays.  Crystal literals are
written with the oracle's `R` / `rng` / `Step` stand-ins:
    a..b -> rng(a, b)      a...b -> rng(a, b, exclusive=True)
    a..s..b -> rng(a, b, s)     (a..s)..b -> R(R(a, s), b)     a..(s..b) -> R(a, R(s, b))
"""
import numpy as np
import pytest

from oracle import ph_oracle as O
from oracle.ph_oracle import R, Step, rng, IndexRegion

BOUND = 10
MID = BOUND // 2


def fully_defined(bound=BOUND):
    """spec/spec_helper.cr:66-100"""
    mid = bound // 2
    return [
        (rng(mid, mid), dict(first=mid, step=1, last=mid)),
        (rng(None, mid), dict(first=0, step=1, last=mid)),
        (rng(None, mid, exclusive=True), dict(first=0, step=1, last=mid - 1)),
        (0, dict(first=0, step=1, last=0)),
        (bound - 1, dict(first=bound - 1, step=1, last=bound - 1)),
        (rng(mid, 0), dict(first=mid, step=-1, last=0)),
        (rng(mid, None, -1), dict(first=mid, step=-1, last=0)),
        (rng(0, bound, 2, exclusive=True), dict(first=0, step=2, last=bound - 1 - ((bound - 1) % 2))),
        (rng(0, bound - 1, 2, exclusive=True), dict(first=0, step=2, last=bound - 2 - (bound % 2))),
        (R(R(5, -3), 1), dict(first=5, step=-3, last=2)),
        (R(5, R(-3, 1)), dict(first=5, step=-3, last=2)),
        (Step(bound - 1, 0, -1), dict(first=bound - 1, step=-1, last=0)),
        (Step(0, bound, 2, True), dict(first=0, step=2, last=bound - 1 - ((bound - 1) % 2))),
        (Step(0, bound - 1, 2, True), dict(first=0, step=2, last=bound - 2 - (bound % 2))),
    ]


def implicit_bounds(bound=BOUND):
    """spec/spec_helper.cr:103-119"""
    mid = bound // 2
    full = dict(first=0, step=1, last=bound - 1)
    return [
        (rng(None, None), full),
        (rng(None, None, exclusive=True), full),
        (rng(mid, None), dict(first=mid, step=1, last=bound - 1)),
        (rng(None, None, -1), dict(first=bound - 1, step=-1, last=0)),
        (rng(None, 2, -1, exclusive=True), dict(first=bound - 1, step=-1, last=3)),
        (R(R(None, -4), None), dict(first=bound - 1, step=-4, last=(bound - 1) % 4)),
        (R(None, R(-4, None)), dict(first=bound - 1, step=-4, last=(bound - 1) % 4)),
    ]


def negative_indices(bound=BOUND):
    """spec/spec_helper.cr:122-133"""
    mid = bound // 2
    full = dict(first=0, step=1, last=bound - 1)
    return [
        (rng(-bound, None), full),
        (rng(None, -bound), dict(first=0, step=1, last=0)),
        (rng(None, -1), full),
        (rng(-mid, -mid + 2), dict(first=bound - mid, step=1, last=bound - mid + 2)),
        (-mid, dict(first=bound - mid, step=1, last=bound - mid)),
    ]


OUT_OF_BOUNDS = [rng(None, BOUND), rng(BOUND, None), rng(-BOUND - 1, None), rng(None, -BOUND - 1),
                 rng(None, BOUND + 1, exclusive=True)]                       # spec_helper.cr:135-143
EMPTY = [rng(None, 0, exclusive=True), rng(3, 3, exclusive=True)]           # :145-150
STEP_CONFLICT = [rng(4, 2, 1), rng(2, 4, -1), Step(4, 2, 1), Step(2, 4, -1)]  # :152-160
VALID = fully_defined() + negative_indices() + implicit_bounds()


def check_region(r: IndexRegion, v):
    """spec/index_region_spec.cr:12-17"""
    assert r.first[0] == v["first"]
    assert r.step[0] == v["step"]
    assert r.last[0] == v["last"]
    assert r.shape[0] == (v["last"] - v["first"]) // v["step"] + 1


# ---- IndexRegion.new(region_literal, bound_shape): index_region_spec.cr:30-80 ----
@pytest.mark.parametrize("lit,v", VALID)
def test_region_valid_literals(lit, v):
    check_region(IndexRegion.new([lit], [BOUND]), v)


@pytest.mark.parametrize("lit", OUT_OF_BOUNDS)
def test_region_out_of_bounds(lit):
    with pytest.raises(O.CrIndexError):
        IndexRegion.new([lit], [BOUND])


@pytest.mark.parametrize("lit", STEP_CONFLICT)
def test_region_step_conflict(lit):
    with pytest.raises(O.CrIndexError):
        IndexRegion.new([lit], [BOUND])


@pytest.mark.parametrize("lit", EMPTY)
def test_region_empty(lit):
    r = IndexRegion.new([lit], [BOUND])
    assert O.shape_to_size(r.shape) == 0


def test_region_multidim():
    """index_region_spec.cr:64-77"""
    r = IndexRegion.new([rng(0, 8, 2), rng(None, None, -1)], [10, 4])
    assert len(r.shape) == 2 and r.shape == [5, 4]
    assert r.first == [0, 3] and r.last == [8, 0] and r.step == [2, -1]
    with pytest.raises(O.CrIndexError):
        IndexRegion.new([rng(0, 3), rng(0, 8, 2)], [10, 4])


# ---- IndexRegion.new(region_literal) absolute: index_region_spec.cr:88-138 ----
@pytest.mark.parametrize("lit,v", fully_defined())
def test_absolute_fully_defined(lit, v):
    check_region(IndexRegion.absolute([lit]), v)


@pytest.mark.parametrize("lit,v", implicit_bounds())
def test_absolute_implicit_bounds_raise(lit, v):
    with pytest.raises(Exception):
        IndexRegion.absolute([lit])


@pytest.mark.parametrize("lit,v", negative_indices())
def test_absolute_negative_raise(lit, v):
    with pytest.raises(O.CrIndexError):
        IndexRegion.absolute([lit])


@pytest.mark.parametrize("lit", STEP_CONFLICT)
def test_absolute_step_conflict(lit):
    with pytest.raises(O.CrIndexError):
        IndexRegion.absolute([lit])


@pytest.mark.parametrize("lit", EMPTY)
def test_absolute_empty(lit):
    assert O.shape_to_size(IndexRegion.absolute([lit]).shape) == 0


def test_absolute_multidim():
    """index_region_spec.cr:131-138"""
    r = IndexRegion.absolute([rng(0, 8, 2), rng(3, None, -1)])
    assert r.shape == [5, 4] and r.first == [0, 3] and r.last == [8, 0] and r.step == [2, -1]


# ---- copy ctor / fits_in? / trim! / translate! / reverse! / includes? ----
@pytest.mark.parametrize("lit,v", VALID)
def test_fits_in(lit, v):
    """index_region_spec.cr:140-166, 246-266"""
    r = IndexRegion.new([lit], [BOUND])
    assert r.fits_in([BOUND]) and r.fits_in([BOUND + 5])
    assert not r.fits_in([max(v["first"], v["last"])])
    mx = max(v["first"], v["last"])
    if mx > 0:
        assert not r.fits_in([mx - 1])


def test_fits_in_multidim():
    r = IndexRegion.absolute([rng(0, 8, 2), rng(3, None, -1)])
    assert r.fits_in([10, 4]) and not r.fits_in([4, 10])


@pytest.mark.parametrize("lit,v", VALID)
def test_trim(lit, v):
    """index_region_spec.cr:268-305"""
    r = IndexRegion.new([lit], [BOUND])
    check_region(r.trim([BOUND]), v)
    if v["last"] > v["first"]:
        t = r.trim([v["last"]])
        check_region(t, dict(first=v["first"], step=v["step"], last=v["last"] - v["step"]))
    elif v["first"] > v["last"]:
        t = r.trim([v["first"]])
        check_region(t, dict(first=v["first"] + v["step"], step=v["step"], last=v["last"]))
    else:
        assert O.shape_to_size(r.trim([v["last"]]).shape) == 0


def test_trim_multidim():
    r = IndexRegion.cover([20, 20]).trim([10, 4])
    assert r.fits_in([10, 4]) and not r.fits_in([4, 10])


def test_translate():
    """index_region_spec.cr:307-340"""
    offset = [-3, 5]
    r = IndexRegion.absolute([rng(3, 20, 4), rng(7, None, -2)])
    t = r.translate(offset)
    assert t.shape == r.shape
    assert t.first == [a + b for a, b in zip(r.first, offset)]
    assert t.last == [a + b for a, b in zip(r.last, offset)]
    assert t.local_to_absolute_unsafe([3, 1]) == [a + b for a, b in zip(r.local_to_absolute_unsafe([3, 1]), offset)]
    with pytest.raises(O.CrIndexError):
        r.translate([0, -2])
    with pytest.raises(O.CrIndexError):
        r.translate([-4, 0])


def test_reverse():
    """index_region_spec.cr:342-371"""
    r = IndexRegion.absolute([rng(2, 6), rng(8, 1, -2)])
    rev = r.reverse()
    assert rev.first == [6, 2]
    assert rev.first == r.last and rev.last == r.first
    assert rev.step == [-s for s in r.step]
    assert rev.shape == r.shape and rev.degeneracy == r.degeneracy


def test_includes():
    """index_region_spec.cr:218-244"""
    r = IndexRegion.absolute([rng(3, 5), rng(10, 2, -2)])
    for c in ([6, 4], [2, 4], [4, 11], [4, 0], [0, 0], [4, 5]):
        assert not r.includes(c)
    for a in range(3, 6):
        for b in range(10, 1, -2):
            assert r.includes([a, b])


def test_ndim_and_dropping():
    """spec/spec_helper.cr:163-193 (ndim / ndim_dropped shapes)."""
    b = [10, 10]
    assert IndexRegion.new([rng(0, 8, 2), rng(0, 3)], b).shape == [5, 4]
    assert IndexRegion.new([rng(3, 3, exclusive=True), rng(0, 3)], b).shape == [0, 4]
    assert IndexRegion.new([rng(0, 0), rng(0, 0)], b).shape == [1, 1]
    assert IndexRegion.new([0, rng(0, 8, 2)], b).shape == [5]
    assert IndexRegion.new([rng(3, 3, exclusive=True), 1], b).shape == [0]
    assert IndexRegion.new([rng(0, 0), 1], b).shape == [1]
    assert IndexRegion.new([1, 1], b).shape == [1]
    assert IndexRegion.new([rng(0, 1), 4, 3, rng(1, 1)], [10] * 4).shape == [2, 1]
    assert IndexRegion.new([rng(0, 1), 4, 3, rng(1, 1)], [10] * 4, drop=False).shape == [2, 1, 1, 1]
    assert IndexRegion.new([1], [2, 3]).shape == [3]                # index_region.cr:181-183 doc
    with pytest.raises(O.DimensionError):
        IndexRegion.new([rng(None, None), 3], [3])                   # :186-187


# ---- coord_util_spec.cr / shape_util_spec.cr ----
def test_coord_util():
    """spec/coord_util_spec.cr"""
    i32max, i32min = 2**31 - 1, -2**31
    assert O.has_index(0, 5) and O.has_index(-5, 5) and not O.has_index(5, 5) and not O.has_index(-6, 5)
    assert O.has_index(i32max - 1, i32max) and not O.has_index(i32max, i32max)
    assert O.has_index(-i32max, i32max) and not O.has_index(i32min, i32max)
    assert O.canonicalize_index(-1, 5) == 4 and O.canonicalize_index(3, 5) == 3
    with pytest.raises(O.CrIndexError):
        O.canonicalize_index(5, 5)
    assert O.canonicalize_coord([-1, 0, -3], [2, 3, 4]) == [1, 0, 1]
    with pytest.raises(O.DimensionError):
        O.canonicalize_coord([0], [2, 3])
    assert O.has_coord([1, -3], [2, 3]) and not O.has_coord([2, 0], [2, 3]) and not O.has_coord([0], [2, 3])


def test_shape_util():
    """spec/shape_util_spec.cr:7-52"""
    assert O.compatible_shapes([], []) and not O.compatible_shapes([], [1]) and not O.compatible_shapes([1], [])
    assert O.compatible_shapes([2, 3], [2, 3]) and O.compatible_shapes([2, 3], [2, 3, 1, 1])
    assert O.compatible_shapes([2, 3, 1], [2, 3]) and not O.compatible_shapes([2, 3], [3, 2])
    assert not O.compatible_shapes([2, 3], [2, 3, 2]) and not O.compatible_shapes([1, 2, 3], [2, 3])
    assert O.shape_to_size([]) == 0 and O.shape_to_size([1]) == 1 and O.shape_to_size([2, 3]) == 6
    assert O.axis_strides([2, 3, 4]) == [12, 4, 1]
    assert O.index_to_coord(17, [2, 3, 4]) == [1, 1, 1]


# ---- n_array_spec.cr gather / scatter / mask / elementwise ----
def stock():
    return np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32)


def test_fetch_chunk_goldens():
    """spec/n_array_spec.cr:211-229"""
    s = stock()
    assert O.fetch_chunk(s, IndexRegion.new([1, rng(0, 2, 2)], [2, 3])).tolist() == [3, 5]
    assert O.fetch_chunk(s, IndexRegion.new([-2, rng(-1, 0)], [2, 3])).tolist() == [2, 1, 0]
    e = O.fetch_chunk(s, IndexRegion.new([rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)], [2, 3]))
    assert e.shape == (0, 0)
    assert s[1, 1] == 4                                              # :231-236


def test_set_chunk_goldens():
    """spec/n_array_spec.cr:238-295"""
    n = stock(); O.set_chunk_array(n, IndexRegion.new([1, rng(0, 2, 2)], [2, 3]), np.array([6, 7], np.int32))
    assert n.tolist() == [[0, 1, 2], [6, 4, 7]]
    n = stock(); O.set_chunk_array(n, IndexRegion.new([-2, rng(-1, 0)], [2, 3]), np.array([6, 7, 8], np.int32))
    assert n.tolist() == [[8, 7, 6], [3, 4, 5]]
    empty = IndexRegion.new([rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)], [2, 3])
    n = stock(); O.set_chunk_array(n, empty, np.zeros(0, np.int32)); assert n.tolist() == stock().tolist()
    n = stock(); O.set_chunk_scalar(n, IndexRegion.new([1, rng(0, 2, 2)], [2, 3]), 6)
    assert n.tolist() == [[0, 1, 2], [6, 4, 6]]
    n = stock(); O.set_chunk_scalar(n, IndexRegion.new([-2, rng(-1, 0)], [2, 3]), 6)
    assert n.tolist() == [[6, 6, 6], [3, 4, 5]]
    n = stock(); O.set_chunk_scalar(n, empty, 6); assert n.tolist() == stock().tolist()


def test_multi_writable_goldens():
    """spec/multi_writable_spec.cr:14-41, 44-65 on the 3x4 buffer (numeric stand-in 0..11)."""
    base = np.arange(12, dtype=np.int64).reshape(3, 4)
    n = base.copy(); O.set_chunk(n, [rng(1, None), rng(1, None)], 10)
    exp = base.reshape(-1).copy(); exp[5:8] = 10; exp[9:12] = 10
    assert n.reshape(-1).tolist() == exp.tolist()
    n = base.copy(); O.set_chunk(n, [rng(1, None), rng(1, None)], np.arange(10, 16).reshape(2, 3))
    exp = base.reshape(-1).copy(); exp[5:8] = [10, 11, 12]; exp[9:12] = [13, 14, 15]
    assert n.reshape(-1).tolist() == exp.tolist()
    with pytest.raises(O.ShapeError):
        O.set_chunk(base.copy(), [rng(2, None), rng(2, None)], np.arange(10, 16).reshape(2, 3))
    # relative set_element :75-81
    idx = O.coord_to_index_fast(O.canonicalize_coord([-1, -2], [3, 4]), O.axis_strides([3, 4]))
    assert idx == (3 - 1) * 4 + (4 - 2)
    for bad in ([10, 10], [-10, -10]):
        with pytest.raises(O.CrIndexError):
            O.canonicalize_coord(bad, [3, 4])


def test_mask_goldens():
    """spec/n_array_spec.cr:297-333"""
    mask = np.array([[True, False, True], [False, True, False]])
    n = stock(); O.mask_set(n, mask, 6); assert n.tolist() == [[6, 1, 6], [3, 6, 5]]
    src, fl = O.ewise("+", stock(), np.int32(10)); assert not fl
    n = stock(); O.mask_set(n, mask, src); assert n.tolist() == [[10, 1, 12], [3, 14, 5]]
    with pytest.raises(O.DimensionError):
        O.mask_set(stock(), mask.reshape(3, 2), 6)


def test_elementwise_goldens():
    """spec/n_array_spec.cr:317-319, 446-447, 462-466; README.md:22-41."""
    assert O.ewise("**", stock(), np.int32(2))[0].tolist() == [[0, 1, 4], [9, 16, 25]]
    assert O.ewise("*", stock(), np.int32(2))[0].tolist() == [[0, 2, 4], [6, 8, 10]]
    a = np.array([[1, 0, 0], [0, 1, 0]], np.int32)
    b = np.array([[0, 1, 2], [10, 11, 12]], np.int32)
    assert O.ewise("+", a, b)[0].tolist() == [[1, 1, 2], [10, 12, 12]]
    assert O.ewise("*", a, b)[0].tolist() == [[0, 0, 0], [0, 11, 0]]
    assert O.get_chunk(a, [rng(None, None), 1]).tolist() == [0, 1]            # README.md:46
    val, idx = O.reduce_argmax(b)                                              # README.md:56-61
    assert (int(val), O.index_to_coord(idx, [2, 3])) == (12, [1, 2])
    assert [s.tolist() for s in O.each_slice(b, 1)] == [[0, 10], [1, 11], [2, 12]]   # README.md:63-64
    with pytest.raises(O.ShapeError):
        O.check_same_shape([2, 3], [1, 2])                                     # multi_indexable.cr:935-940
    with pytest.raises(O.DimensionError):
        O.check_same_shape([2, 3], [1, 2], what="eq")                          # :900-902


# ---- transforms / views ----
def test_transform_goldens():
    """spec/view_util/{permute,reshape,reverse}_transform_spec.cr"""
    assert O.PermuteTransform([2, 6, 5, 0, 1, 3, 4]).apply([3, 5, 2, 1, 0, 9, 4]) == [1, 0, 3, 9, 4, 2, 5]
    assert O.ReshapeTransform([3, 4], [6, 2]).apply([2, 1]) == [1, 1]
    assert O.ReverseTransform([5, 3]).apply([1, 2]) == [3, 0]


def test_iteration_orders():
    """spec/multi_indexable/elem_iterator_spec.cr:10-68"""
    n = np.arange(1, 10).reshape(3, 3)
    cov = IndexRegion.cover([3, 3])
    lex = [int(n[tuple(c)]) for c in O.lex_coords(cov.first, cov.step, cov.last)]
    colex = [int(n[tuple(c)]) for c in O.colex_coords(cov.first, cov.step, cov.last)]
    assert lex == [1, 2, 3, 4, 5, 6, 7, 8, 9]
    assert colex == [1, 4, 7, 2, 5, 8, 3, 6, 9]
    rev = cov.reverse()
    assert [int(n[tuple(c)]) for c in O.lex_coords(rev.first, rev.step, rev.last)] == lex[::-1]
    assert [int(n[tuple(c)]) for c in O.colex_coords(rev.first, rev.step, rev.last)] == colex[::-1]
    assert [list(c) for c in O.lex_coords(cov.first, cov.step, cov.last)][:4] == [[0, 0], [0, 1], [0, 2], [1, 0]]


def valid_regions(shape):
    """spec/multi_indexable/multi_indexable_tester.cr:129-163"""
    return [
        [rng(0, a, exclusive=True) for a in shape],
        [rng(1, a, exclusive=True) if a > 1 else rng(0, a, exclusive=True) for a in shape],
        [rng(0, a - 1, exclusive=True) if a > 2 else rng(0, a, exclusive=True) for a in shape],
        [rng(0, a, 2, exclusive=True) for a in shape],
        [rng(a - 1, 0, -2) for a in shape],
        [rng(None, None)] * len(shape),
        [],
    ]


def invalid_regions(shape):
    """spec/multi_indexable/multi_indexable_tester.cr:165-182"""
    return [
        ([rng(a, 2 * a) for a in shape], O.CrIndexError),
        ([rng(0, a, exclusive=True) for a in shape] + [rng(0, 0)], O.DimensionError),
        ([rng(-a - 1, None, exclusive=True) for a in shape], O.CrIndexError),
        ([rng(a // 2, a + 3) for a in shape], O.CrIndexError),
        ([rng(a, 0, 1) for a in shape], O.CrIndexError),
    ]


@pytest.mark.parametrize("shape", [[2, 3, 4], [3, 5], [3, 4], [1], [1, 1, 1]])
@pytest.mark.parametrize("drop", [True, False])
def test_conformance_regions_against_numpy(shape, drop):
    """Tester regions (multi_indexable_tester.cr:416-456): oracle gather == numpy basic
    slicing wherever the two semantics coincide (positive-size regions)."""
    n = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)
    for lit in valid_regions(shape):
        region = IndexRegion.new(lit, shape, drop)
        got = O.fetch_chunk(n, region)
        sl = tuple(slice(f, (l + s) if (l + s) >= 0 else None, s) for f, s, l in zip(region.first, region.step, region.last))
        assert got.reshape(-1).tolist() == n[sl].reshape(-1).tolist()
        assert got.tolist() == O.fetch_chunk_fast(n, region).tolist()
    for lit, exc in invalid_regions(shape):
        with pytest.raises(exc):
            IndexRegion.new(lit, shape, drop)


@pytest.mark.parametrize("shape", [[0], [5, 0, 2], [3, 0, 0, 1]])
def test_empty_shapes(shape):
    """spec/n_array_spec.cr:335-442 family: empty arrays iterate zero elements."""
    cov = IndexRegion.cover(shape)
    assert list(O.lex_coords(cov.first, cov.step, cov.last)) == []
    assert O.shape_to_size(shape) == 0


def test_view_chain_matches_numpy():
    """View / transforms (view.cr:43-126): region -> permute -> reverse chains agree with
    the equivalent numpy slicing / transpose / flip."""
    n = np.arange(2 * 3 * 4, dtype=np.int64).reshape(2, 3, 4)
    v = O.View(n).view([rng(None, None), rng(2, 0), rng(0, 3, 2)])
    assert v.to_narr().tolist() == n[:, ::-1, 0:4:2].tolist()
    p = v.permute([2, 0, 1])
    assert p.shape == [2, 2, 3]
    assert p.to_narr().tolist() == np.transpose(n[:, ::-1, 0:4:2], (2, 0, 1)).tolist()
    assert O.View(n).permute().to_narr().tolist() == np.transpose(n).tolist()          # default = reversed axes
    r = p.reverse()
    assert r.to_narr().tolist() == np.transpose(n[:, ::-1, 0:4:2], (2, 0, 1))[::-1, ::-1, ::-1].tolist()
    d = O.View(n).view([1, rng(None, None), 2])                                          # dropped axes
    assert d.shape == [3] and d.to_narr().tolist() == n[1, :, 2].tolist()
    rs = O.View(n).reshape([6, 4])
    assert rs.to_narr().tolist() == n.reshape(6, 4).tolist()
    with pytest.raises(O.ShapeError):
        O.View(n).reshape([5, 5])
    with pytest.raises(O.CrIndexError):
        O.View(n).permute([0, 1, 3])


def test_mutable_view_scatter():
    """mutable_view.cr:16-18 through multi_writable.cr:29-44."""
    n = np.zeros((3, 4), dtype=np.int64)
    v = O.View(n).permute([1, 0])                      # view shape [4, 3]
    v.set_chunk([rng(None, None), rng(None, None)], np.arange(12).reshape(4, 3))
    assert n.tolist() == np.arange(12).reshape(4, 3).T.tolist()
    v.set_chunk([rng(1, 2), 0], 99)
    assert n[0, 1] == 99 and n[0, 2] == 99


def test_tile_golden():
    """multi_indexable.cr:806-816 doc example."""
    unit = np.array([[1, 2], [3, 4]])
    assert O.tile(unit, [2, 3]).tolist() == [[1, 2, 1, 2, 1, 2], [3, 4, 3, 4, 3, 4]] * 2


# ---- number semantics (restated; unpinned by the reference, pinned here by construction) ----
def test_integer_semantics():
    i32 = np.int32
    mx, mn = np.iinfo(i32).max, np.iinfo(i32).min
    r, f = O.ewise("+", np.array([mx, 1], i32), np.array([1, 1], i32)); assert "overflow" in f and r.tolist() == [mn, 2]
    r, f = O.ewise("&+", np.array([mx], i32), np.array([1], i32)); assert not f and r.tolist() == [mn]
    r, f = O.ewise("//", np.array([7, -7, 7, -7], i32), np.array([2, 2, -2, -2], i32)); assert r.tolist() == [3, -4, -4, 3]
    r, f = O.ewise("%", np.array([7, -7, 7, -7], i32), np.array([2, 2, -2, -2], i32)); assert r.tolist() == [1, 1, -1, -1]
    r, f = O.ewise("//", np.array([1], i32), np.array([0], i32)); assert "div0" in f
    r, f = O.ewise("//", np.array([mn], i32), np.array([-1], i32)); assert "argument" in f
    r, f = O.ewise("/", np.array([1, 3], i32), np.array([2, 2], i32)); assert r.dtype == np.float64 and r.tolist() == [0.5, 1.5]
    r, f = O.ewise("**", np.array([2, 3], i32), np.array([31, 3], i32)); assert "overflow" in f
    r, f = O.ewise("**", np.array([2], i32), np.array([-1], i32)); assert "argument" in f
    r, f = O.ewise("&**", np.array([2], i32), np.array([31], i32)); assert not f and r.tolist() == [mn]
    r, f = O.unary("-", np.array([mn, 5], i32)); assert "overflow" in f
    r, f = O.ewise("-", i32(10), np.array([1, 2], i32)); assert r.tolist() == [9, 8]        # scalar on the left
    assert O.ewise("^", np.array([6], i32), np.array([3], i32))[0].tolist() == [5]
    assert O.unary("~", np.array([0], i32))[0].tolist() == [-1]


def test_float_semantics():
    f32 = np.float32
    a, b, c = f32(1.0000001), f32(3.0000002), f32(-3.0000005)
    two_step = f32(f32(a * b) + c)                      # two roundings, never an FMA
    r1, _ = O.ewise("*", np.array([a]), np.array([b])); r2, _ = O.ewise("+", r1, np.array([c]))
    assert r2[0] == two_step
    r, f = O.ewise("%", np.array([5.5, -5.5], f32), np.array([2.0, 2.0], f32)); assert r.tolist() == [1.5, 0.5]
    r, f = O.ewise("%", np.array([1.0], f32), np.array([0.0], f32)); assert "div0" in f
    r, f = O.ewise("//", np.array([5.5, -5.5], f32), np.array([2.0, 2.0], f32)); assert r.tolist() == [2.0, -3.0]
    r, _ = O.ewise("**", np.array([3.0, 0.5], np.float64), np.int32(3)); assert r.tolist() == [27.0, 0.125]
    r, _ = O.ewise("**", np.array([2.0], np.float64), np.int32(-2)); assert r.tolist() == [0.25]
    assert O.compare(">", np.array([np.nan, 1.0]), np.array([0.0, 0.0])).tolist() == [False, True]


def test_reductions_semantics():
    a = np.array([[3, 9, 2], [9, 1, 0]], np.float32)
    assert O.reduce_argmax(a) == (np.float32(9), 1)                  # first maximum wins
    assert O.reduce_minmax(a, "max") == 9 and O.reduce_minmax(a, "min") == 0
    assert O.reduce_sum_sequential(a) == np.float32(24)
    assert O.reduce_axis(a, 0, "sum").tolist() == [12, 10, 2]
    assert O.reduce_axis(a, 1, "max").tolist() == [9, 9]
    assert O.reduce_axis(a, 1, "argmax").tolist() == [1, 0]
    assert O.reduce_axis(a, 0, "argmax").tolist() == [1, 0, 0]
    with pytest.raises(O.CrArgumentError):
        O.reduce_minmax(np.array([1.0, np.nan]), "max")
    with pytest.raises(O.CrEmptyError):
        O.reduce_minmax(np.zeros(0), "max")
    with pytest.raises(O.CrOverflowError):
        O.reduce_sum_sequential(np.array([2**31 - 1, 1, -5], np.int32))
    # stagnation of the sequential f32 fold
    big = np.full(1 << 12, 0.5, np.float32)
    assert O.reduce_sum_sequential(np.concatenate([[np.float32(2**24)], big])) == np.float32(2**24)


def test_heat_example_anchor():
    """examples/heat_equation.cr: constants and the survey's self-consistency anchor."""
    c = O.heat_example_coeff()
    assert c == 0.0003901234567901234 and c.hex() == "0x1.9912f7d0247d5p-12"
    s = O.heat_simulate_1d_example()
    assert s.shape == (21,)
    assert abs(float(s.sum()) - 480.0) < 1e-9                        # zero-flux ends conserve heat
    np.testing.assert_allclose(s[[0, 1, -2, -1]], [14.381532, 15.076701, 39.693196, 42.473872], atol=5e-7)


def test_heat_nd_definition():
    rs = np.random.RandomState(3)
    s = rs.rand(6, 7, 5).astype(np.float32)
    n = O.heat_step_nd(s, 0.1)
    assert np.array_equal(n[0], s[0]) and np.array_equal(n[:, :, -1], s[:, :, -1])
    f = np.float32
    i, j, k = 2, 3, 2
    c = s[i, j, k]
    d0 = f(f(s[i - 1, j, k] - f(f(2) * c)) + s[i + 1, j, k])
    d1 = f(f(s[i, j - 1, k] - f(f(2) * c)) + s[i, j + 1, k])
    d2 = f(f(s[i, j, k - 1] - f(f(2) * c)) + s[i, j, k + 1])
    assert n[i, j, k] == f(c + f(f(f(d0 + d1) + d2) * f(0.1)))


def test_c_port_heat_and_sum_match_the_numpy_oracle():
    """The C ports timed as CPU baselines (reference-structured and flat OpenMP) restate the
    same arithmetic as the numpy oracle: 3-D heat step bit for bit, sequential f32 sum exactly."""
    from oracle import c_oracle as CO
    rs = np.random.RandomState(9)
    for shape in [(3, 3, 3), (5, 4, 7), (20, 17, 33), (2, 9, 9)]:
        s = (rs.rand(*shape) * 100).astype(np.float32)
        want = O.heat_step_nd(s, np.float32(0.1))
        assert CO.ref_heat_step_3d_f32(s, 0.1).tobytes() == want.tobytes(), shape
        assert CO.flat_heat_step_3d_f32(s, 0.1).tobytes() == want.tobytes(), shape
    x = rs.rand(100_003).astype(np.float32)
    acc = np.float32(0)
    for v in x[:2000]:
        acc = np.float32(acc + v)                                    # Enumerable#sum: left fold in T
    assert CO.ref_sum_f32(x[:2000].copy()) == acc
    assert abs(CO.flat_sum_f32(x) - float(x.astype(np.float64).sum())) < 1e-6 * x.size


def test_powi_is_pinned_to_an_independent_runtime_implementation(tmp_path):
    """`Float ** Int32` lowers to llvm.powi, i.e. compiler-rt's __powisf2 / __powidf2.  Neither
    Crystal nor compiler-rt is in this image; libgcc's __powisf2 / __powidf2 (reached through gcc's
    __builtin_powi) are an INDEPENDENT implementation of the same published square-and-multiply order
    (odd exponent bits multiply the running square into the result, 1 / r at the end for n < 0).  The
    oracle's restatement must agree with it bit for bit."""
    import ctypes
    import subprocess
    src = tmp_path / "powi_ref.c"
    src.write_text("double powi_f64(double x, int n) { return __builtin_powi(x, n); }\n"
                   "float powi_f32(float x, int n) { return __builtin_powif(x, n); }\n")
    lib_path = tmp_path / "libpowi_ref.so"
    subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(lib_path), str(src)], check=True)
    nm = subprocess.run(["nm", str(lib_path)], capture_output=True, text=True).stdout         # libgcc.a links them statically
    assert "__powidf2" in nm and "__powisf2" in nm                 # really the runtime routine, not an inlined pow()
    lib = ctypes.CDLL(str(lib_path))
    lib.powi_f64.restype, lib.powi_f64.argtypes = ctypes.c_double, [ctypes.c_double, ctypes.c_int]
    lib.powi_f32.restype, lib.powi_f32.argtypes = ctypes.c_float, [ctypes.c_float, ctypes.c_int]
    rs = np.random.RandomState(31)
    xs64 = np.concatenate([rs.randn(200) * 3, [0.0, -0.0, 1.0, -1.0, 0.05, 1e-160, -1e160, 5e-324, np.inf, -np.inf, np.nan]])
    exps = [0, 1, 2, 3, 4, 5, 7, 10, 13, 31, 64, 255, 1000, -1, -2, -3, -7, -64, 2 ** 31 - 1, -(2 ** 31) + 1]
    for n in exps:
        got64 = O._powi(xs64.astype(np.float64), n)
        want64 = np.array([lib.powi_f64(float(x), n) for x in xs64], dtype=np.float64)
        assert np.array_equal(np.isnan(got64), np.isnan(want64))
        assert got64[~np.isnan(got64)].tobytes() == want64[~np.isnan(want64)].tobytes(), f"f64 ** {n}"
        xs32 = xs64.astype(np.float32)
        got32 = O._powi(xs32, n)
        want32 = np.array([lib.powi_f32(float(x), n) for x in xs32], dtype=np.float32)
        assert np.array_equal(np.isnan(got32), np.isnan(want32))
        assert got32[~np.isnan(got32)].tobytes() == want32[~np.isnan(want32)].tobytes(), f"f32 ** {n}"
    # the example's constant: SPACING ** 2 with Float64 ** Int32 (examples/heat_equation.cr:20)
    assert lib.powi_f64(0.05, 2) == 0.05 * 0.05 == float(O._powi(np.array(0.05), 2))


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64])
def test_integer_semantics_against_exact_arithmetic(dtype):
    """Crystal's integer operators restated INDEPENDENTLY with Python's unbounded integers: `+ - *` are exact
    or raise OverflowError; `&+ &- &*` wrap modulo 2^bits; `//` and `%` are floored (the sign of the divisor),
    x // 0 and x % 0 raise DivisionByZeroError, MIN // -1 raises ArgumentError; `**` is exact or overflows,
    `&**` wraps, a negative exponent raises ArgumentError.  The oracle's numpy restatement must agree
    element by element: value (as wrapped) and the set of errors."""
    info = np.iinfo(dtype)
    bits = info.bits
    lo, hi = int(info.min), int(info.max)

    def wrap(v):
        v &= (1 << bits) - 1
        return v - (1 << bits) if (lo < 0 and v > hi) else v

    rs = np.random.RandomState(bits + (1 if lo < 0 else 0))
    edge = [lo, lo + 1, hi, hi - 1, 0, 1, 2, 3, 7, hi // 2, hi // 2 + 1] + ([-1, -2, -3, lo // 2] if lo < 0 else [])
    pool = edge + [int(v) for v in rs.randint(max(lo, -1000), min(hi, 1000) + 1, size=40)]
    pairs = [(x, y) for x in pool for y in edge + pool[-8:]]

    def check(op, model):
        # one element at a time so the error set is per element
        for (x, y) in pairs[:: max(1, len(pairs) // 400)]:
            r, flags = O.ewise(op, np.array([x], dtype), np.array([y], dtype))
            want_v, want_f = model(x, y)
            assert flags == want_f, f"{np.dtype(dtype)} {x} {op} {y}: flags {flags} != {want_f}"
            if want_v is not None and not want_f:
                assert int(r[0]) == want_v, f"{np.dtype(dtype)} {x} {op} {y}: {int(r[0])} != {want_v}"

    def exact(fn):
        def model(x, y):
            v = fn(x, y)
            return (v, set()) if lo <= v <= hi else (None, {"overflow"})
        return model

    check("+", exact(lambda x, y: x + y))
    check("-", exact(lambda x, y: x - y))
    check("*", exact(lambda x, y: x * y))
    check("&+", lambda x, y: (wrap(x + y), set()))
    check("&-", lambda x, y: (wrap(x - y), set()))
    check("&*", lambda x, y: (wrap(x * y), set()))

    def floordiv(x, y):
        if y == 0:
            return None, {"div0"}
        if lo < 0 and x == lo and y == -1:
            return None, {"argument"}
        return x // y, set()

    def mod(x, y):
        if y == 0:
            return None, {"div0"}
        return x % y, set()                                            # Python's % is floored, like Crystal's

    check("//", floordiv)
    check("%", mod)
    for base in [0, 1, 2, 3, 7, 10] + ([-1, -2, -3] if lo < 0 else []):
        for e in [0, 1, 2, 3, 5, 7, 8, 15, 16, 31, 32, 63, 64] + ([-1] if lo < 0 else []):
            if not (lo <= e <= hi):
                continue
            r, flags = O.ewise("**", np.array([base], dtype), np.array([e], dtype))
            if e < 0:
                assert flags == {"argument"}
                continue
            v = base ** e
            assert flags == (set() if lo <= v <= hi else {"overflow"}), f"{base} ** {e}"
            if not flags:
                assert int(r[0]) == v
            r, flags = O.ewise("&**", np.array([base], dtype), np.array([e], dtype))
            assert not flags and int(r[0]) == wrap(v), f"{base} &** {e}"
