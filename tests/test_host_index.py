"""T0/T1 (CPU): the C++ host layer (include/ph_host.h) against the reference's golden vectors
and against the oracle's coordinate enumeration.  No GPU: these functions are pure host code
living in libphgpu.so.  Reference: src/range_syntax/range_syntax.cr, src/index_region.cr,
src/view_util/transforms.cr, src/shape_util.cr, src/coord_util.cr."""
import ctypes as C

import numpy as np
import pytest

import ph_core_b200 as ph
from ph_core_b200 import _lib, make_region
from oracle.ph_oracle import rng, R, Step   # the marshaller is duck-typed: one literal feeds both sides
from ph_core_b200.narray import host_check, _i64
from oracle import ph_oracle as O
import test_oracle_goldens as G


def desc_offsets(d: ph.PhDesc) -> np.ndarray:
    """Enumerate the buffer offsets a descriptor addresses, in lex order of its coordinates."""
    total = np.zeros((), dtype=np.int64) + d.offset
    for i in range(d.rank):
        shp = [1] * d.rank
        shp[i] = d.extent[i]
        total = total + (np.arange(d.extent[i], dtype=np.int64) * d.stride[i]).reshape(shp)
    if d.rank == 0:
        return np.array([d.offset])
    return np.asarray(total).reshape(-1)


def region_desc(literal, shape, drop=True):
    lib = _lib.load()
    reg = make_region(literal, shape, drop)
    src, out = ph.PhDesc.contiguous(shape), ph.PhDesc()
    host_check(lib.ph_desc_region(C.byref(src), C.byref(reg), C.byref(out)))
    return reg, out


def check_axis(reg, v):
    assert reg.first[0] == v["first"] and reg.step[0] == v["step"] and reg.last[0] == v["last"]
    assert reg.shape[0] == (v["last"] - v["first"]) // v["step"] + 1


@pytest.mark.parametrize("lit,v", G.VALID)
def test_valid_literals(lit, v):
    """spec/index_region_spec.cr:30-40 through ph_region_new"""
    check_axis(make_region([lit], [G.BOUND]), v)


@pytest.mark.parametrize("lit", G.OUT_OF_BOUNDS + G.STEP_CONFLICT)
def test_raising_literals(lit):
    with pytest.raises(ph.CrIndexError):
        make_region([lit], [G.BOUND])


@pytest.mark.parametrize("lit", G.EMPTY)
def test_empty_literals(lit):
    reg = make_region([lit], [G.BOUND])
    assert reg.shape == [0]


def test_multidim_and_drop():
    reg = make_region([rng(0, 8, 2), rng(None, None, -1)], [10, 4])
    assert reg.shape == [5, 4] and list(reg.first)[:2] == [0, 3] and list(reg.last)[:2] == [8, 0]
    assert list(reg.step)[:2] == [2, -1]
    with pytest.raises(ph.CrIndexError):
        make_region([rng(0, 3), rng(0, 8, 2)], [10, 4])
    with pytest.raises(ph.DimensionError):
        make_region([rng(None, None), 3], [3])
    b = [10, 10]
    assert make_region([0, rng(0, 8, 2)], b).shape == [5]
    assert make_region([rng(3, 3, exclusive=True), 1], b).shape == [0]
    assert make_region([1, 1], b).shape == [1]
    assert make_region([rng(0, 1), 4, 3, rng(1, 1)], [10] * 4).shape == [2, 1]
    assert make_region([rng(0, 1), 4, 3, rng(1, 1)], [10] * 4, drop=False).shape == [2, 1, 1, 1]
    assert make_region([1], [2, 3]).shape == [3]


@pytest.mark.parametrize("lit,v", G.VALID)
def test_trim_fits(lit, v):
    """spec/index_region_spec.cr:246-305"""
    lib = _lib.load()
    reg = make_region([lit], [G.BOUND])
    fits = C.c_int32()
    host_check(lib.ph_region_fits_in(C.byref(reg), _i64([G.BOUND]), 1, C.byref(fits))); assert fits.value == 1
    host_check(lib.ph_region_fits_in(C.byref(reg), _i64([max(v["first"], v["last"])]), 1, C.byref(fits))); assert fits.value == 0
    oreg = O.IndexRegion.new([lit], [G.BOUND])
    for nb in range(0, G.BOUND + 2):
        r2 = make_region([lit], [G.BOUND])
        host_check(lib.ph_region_trim(C.byref(r2), _i64([nb]), 1))
        t = oreg.trim([nb])
        assert (r2.first[0], r2.step[0], r2.last[0], r2.proper_shape[0]) == (t.first[0], t.step[0], t.last[0], t.proper_shape[0])
    with pytest.raises(ph.DimensionError):
        host_check(lib.ph_region_trim(C.byref(reg), _i64([4, 3]), 2))


def test_reverse_translate():
    lib = _lib.load()
    reg = make_region([rng(2, 6), rng(8, 1, -2)], [20, 20])
    host_check(lib.ph_region_reverse(C.byref(reg)))
    assert list(reg.first)[:2] == [6, 2] and list(reg.step)[:2] == [-1, 2]
    reg = make_region([rng(3, 20, 4), rng(7, None, -2)], [30, 30])
    f0, l0 = list(reg.first)[:2], list(reg.last)[:2]
    host_check(lib.ph_region_translate(C.byref(reg), _i64([-3, 5]), 2))
    assert list(reg.first)[:2] == [f0[0] - 3, f0[1] + 5] and list(reg.last)[:2] == [l0[0] - 3, l0[1] + 5]
    with pytest.raises(ph.CrIndexError):
        host_check(lib.ph_region_translate(C.byref(reg), _i64([0, -20]), 2))


def test_shape_and_coord_util():
    lib = _lib.load()
    ok = C.c_int32()
    for a, b, want in [([], [], 1), ([], [1], 0), ([2, 3], [2, 3, 1, 1], 1), ([2, 3, 1], [2, 3], 1),
                       ([2, 3], [3, 2], 0), ([2, 3], [2, 3, 2], 0), ([1, 2, 3], [2, 3], 0)]:
        host_check(lib.ph_shapes_compatible(_i64(a), len(a), _i64(b), len(b), C.byref(ok)))
        assert ok.value == want, (a, b)
    out = (C.c_int64 * 3)()
    host_check(lib.ph_canonicalize_coord(_i64([-1, 0, -3]), 3, _i64([2, 3, 4]), 3, out)); assert list(out) == [1, 0, 1]
    with pytest.raises(ph.DimensionError):
        host_check(lib.ph_canonicalize_coord(_i64([0]), 1, _i64([2, 3]), 2, out))
    with pytest.raises(ph.CrIndexError):
        host_check(lib.ph_canonicalize_coord(_i64([2, 0]), 2, _i64([2, 3]), 2, out))
    bs = (C.c_int64 * 2)()
    host_check(lib.ph_broadcast_shapes(_i64([4, 1]), _i64([1, 5]), 2, bs)); assert list(bs) == [4, 5]
    with pytest.raises(ph.ShapeError):
        host_check(lib.ph_broadcast_shapes(_i64([4, 2]), _i64([1, 5]), 2, bs))


def test_gather_goldens_as_descriptors():
    """spec/n_array_spec.cr:211-229: the descriptor enumerates exactly the golden elements."""
    stock = np.arange(6)
    _, d = region_desc([1, rng(0, 2, 2)], [2, 3]); assert stock[desc_offsets(d)].tolist() == [3, 5]
    _, d = region_desc([-2, rng(-1, 0)], [2, 3]); assert stock[desc_offsets(d)].tolist() == [2, 1, 0]
    reg, d = region_desc([rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)], [2, 3])
    assert reg.shape == [0, 0] and desc_offsets(d).size == 0


@pytest.mark.parametrize("shape", [[2, 3, 4], [3, 5], [3, 4], [1], [1, 1, 1], [7, 1, 6]])
@pytest.mark.parametrize("drop", [True, False])
def test_conformance_regions(shape, drop):
    """multi_indexable_tester.cr:129-182 regions: descriptor offsets == oracle lex iteration."""
    for lit in G.valid_regions(shape):
        reg, d = region_desc(lit, shape, drop)
        oreg = O.IndexRegion.new(lit, shape, drop)
        assert reg.shape == oreg.shape
        assert desc_offsets(d).tolist() == O.lex_buffer_indices(oreg, shape)
    for lit, exc in G.invalid_regions(shape):
        want = {O.CrIndexError: ph.CrIndexError, O.DimensionError: ph.DimensionError}[exc]
        with pytest.raises(want):
            make_region(lit, shape, drop)


def random_literal(rs, bound):
    kind = rs.randint(0, 6)
    if kind == 0:
        return int(rs.randint(-bound, bound))
    a, b = int(rs.randint(-bound, bound)), int(rs.randint(-bound, bound))
    excl = bool(rs.randint(0, 2))
    if kind == 1:
        return rng(a, b, exclusive=excl)
    if kind == 2:
        return rng(a, None) if rs.randint(0, 2) else rng(None, b, exclusive=excl)
    if kind == 3:
        return rng(None, None)
    step = int(rs.choice([-3, -2, -1, 1, 2, 3]))
    if kind == 4:
        return rng(a, b, step, exclusive=excl)
    return R(R(None, step), None) if rs.randint(0, 2) else R(a, R(step, None))


def test_random_regions_match_oracle():
    """SURVEY.md 7.2 verification, re-created: ~2500 random literals x drop, including
    negative steps, empties and raising literals -- same outcome class and same offsets."""
    rs = np.random.RandomState(1234)
    n_ok = 0
    for _ in range(2500):
        rank = int(rs.randint(1, 5))
        shape = [int(rs.randint(1, 7)) for _ in range(rank)]
        lit = [random_literal(rs, shape[i]) for i in range(int(rs.randint(0, rank + 1)))]
        drop = bool(rs.randint(0, 2))
        try:
            oreg = O.IndexRegion.new(lit, shape, drop)
        except O.CrIndexError:
            with pytest.raises(ph.CrIndexError):
                make_region(lit, shape, drop)
            continue
        reg, d = region_desc(lit, shape, drop)
        assert reg.shape == oreg.shape, (lit, shape)
        n = O.shape_to_size(oreg.shape)
        want = O.lex_buffer_indices(oreg, shape)[:n] if n else []
        assert desc_offsets(d).tolist() == want, (lit, shape, drop)
        n_ok += 1
    assert n_ok > 800


def test_random_view_chains_match_oracle():
    """Region / Permute / Reverse / Reshape chains fold into one descriptor that addresses
    the same source elements as the reference's newest-first transform chain."""
    lib = _lib.load()
    rs = np.random.RandomState(99)
    checked = 0
    for _ in range(1500):
        rank = int(rs.randint(1, 5))
        shape = [int(rs.randint(1, 6)) for _ in range(rank)]
        n = int(np.prod(shape))
        src = np.arange(n, dtype=np.int64).reshape(shape)
        ov = O.View(src)
        d, cur_shape = ph.PhDesc.contiguous(shape), list(shape)
        ok = True
        for _ in range(int(rs.randint(1, 5))):
            kind = rs.randint(0, 4)
            nd = ph.PhDesc()
            if kind == 0:
                lit = [random_literal(rs, cur_shape[i]) for i in range(len(cur_shape))]
                try:
                    oreg = O.IndexRegion.new(lit, cur_shape)
                except O.CrIndexError:
                    continue
                if O.shape_to_size(oreg.shape) == 0:
                    continue
                ov = ov.view(oreg)
                reg = make_region(lit, cur_shape)
                host_check(lib.ph_desc_region(C.byref(d), C.byref(reg), C.byref(nd)))
            elif kind == 1:
                order = list(rs.permutation(len(cur_shape))) if rs.randint(0, 2) else None
                ov = ov.permute([int(o) for o in order] if order is not None else None)
                if order is None:
                    host_check(lib.ph_desc_permute(C.byref(d), None, 0, C.byref(nd)))
                else:
                    pat = (C.c_int32 * len(order))(*[int(o) for o in order])
                    host_check(lib.ph_desc_permute(C.byref(d), pat, len(order), C.byref(nd)))
            elif kind == 2:
                ov = ov.reverse()
                host_check(lib.ph_desc_reverse(C.byref(d), C.byref(nd)))
            else:
                total = O.shape_to_size(cur_shape)
                divs = [k for k in range(1, total + 1) if total % k == 0]
                a = int(rs.choice(divs))
                new_shape = [a, total // a]
                st = lib.ph_desc_reshape(C.byref(d), _i64(new_shape), 2, C.byref(nd))
                if st == ph.K["PH_HOST_NEEDS_COPY"]:
                    continue                                     # host would materialise first
                host_check(st)
                ov = ov.reshape(new_shape)
            d = nd
            cur_shape = [int(d.extent[i]) for i in range(d.rank)]
            assert cur_shape == ov.shape
        got = src.reshape(-1)[desc_offsets(d)]
        assert got.tolist() == ov.to_narr().reshape(-1).tolist()
        checked += 1
    assert checked == 1500


def test_transform_goldens():
    """spec/view_util/*_transform_spec.cr as descriptor offsets."""
    lib = _lib.load()
    off = C.c_int64()
    # permute: coord [3,5,2,1,0,9,4] in the view with pattern [2,6,5,0,1,3,4] -> src coord [1,0,3,9,4,2,5]
    shape = [10] * 7
    d, nd = ph.PhDesc.contiguous(shape), ph.PhDesc()
    pat = (C.c_int32 * 7)(2, 6, 5, 0, 1, 3, 4)
    host_check(lib.ph_desc_permute(C.byref(d), pat, 7, C.byref(nd)))
    host_check(lib.ph_desc_offset_of(C.byref(nd), _i64([3, 5, 2, 1, 0, 9, 4]), 7, C.byref(off)))
    assert off.value == O.coord_to_index_fast([1, 0, 3, 9, 4, 2, 5], O.axis_strides(shape))
    # reshape [3,4] -> [6,2]: view coord [2,1] -> src coord [1,1]
    d = ph.PhDesc.contiguous([3, 4])
    host_check(lib.ph_desc_reshape(C.byref(d), _i64([6, 2]), 2, C.byref(nd)))
    host_check(lib.ph_desc_offset_of(C.byref(nd), _i64([2, 1]), 2, C.byref(off)))
    assert off.value == O.coord_to_index_fast([1, 1], [4, 1])
    # reverse [5,3]: [1,2] -> [3,0]
    d = ph.PhDesc.contiguous([5, 3])
    host_check(lib.ph_desc_reverse(C.byref(d), C.byref(nd)))
    host_check(lib.ph_desc_offset_of(C.byref(nd), _i64([1, 2]), 2, C.byref(off)))
    assert off.value == O.coord_to_index_fast([3, 0], [3, 1])
    with pytest.raises(ph.ShapeError):
        host_check(lib.ph_desc_reshape(C.byref(d), _i64([4, 4]), 2, C.byref(nd)))
    with pytest.raises(ph.CrIndexError):
        bad = (C.c_int32 * 2)(0, 2)
        host_check(lib.ph_desc_permute(C.byref(d), bad, 2, C.byref(nd)))


def test_trimmed_regions_match_oracle():
    """IndexRegion.new(literal, bound_shape, trim_to:) (index_region.cr:133-168) -- the region
    behind MultiIndexable#get_available -- against the oracle on random literals."""
    lib = _lib.load()
    rs = np.random.RandomState(77)
    ok = 0
    for _ in range(1500):
        rank = int(rs.randint(1, 4))
        shape = [int(rs.randint(1, 8)) for _ in range(rank)]
        lit = []
        for i in range(int(rs.randint(0, rank + 1))):
            a, b = int(rs.randint(0, 12)), int(rs.randint(0, 12))
            kind = rs.randint(0, 4)
            lit.append(a if kind == 0 else rng(a, b) if kind == 1 else rng(a, None) if kind == 2
                       else rng(a, b, int(rs.choice([-2, -1, 1, 2, 3]))))
        drop = bool(rs.randint(0, 2))
        lits = [ph.region.marshal(l) for l in lit]
        arr = (_lib.PhRangeLit * max(1, len(lits)))(*lits)
        reg = _lib.PhRegion()
        st = lib.ph_region_new_trimmed(arr, len(lits), _i64(shape), _i64(shape), rank, int(drop), C.byref(reg))
        try:
            oreg = O.IndexRegion.new_trimmed(lit, shape, bound_shape=shape, drop=drop)
        except O.CrIndexError:
            assert st == ph.K["PH_HOST_INDEX_ERROR"], (lit, shape)
            continue
        except O.CrDivisionByZeroError:
            continue
        assert st == 0, (lit, shape, lib.ph_host_last_error())
        assert reg.shape == oreg.shape, (lit, shape)
        for i in range(rank):
            if oreg.step[i] != 0:
                assert (reg.first[i], reg.step[i], reg.last[i]) == (oreg.first[i], oreg.step[i], oreg.last[i]), (lit, shape)
        ok += 1
    assert ok > 500


def test_float_text_follows_crystal_float_to_s():
    """f-4 / ADVICE r1: io.format_float restates Crystal 1.0's Float#to_s (shortest round-trip digits in the
    element's own width; positional for decimal points in [-3, 15], else d.de+X with an unpadded exponent);
    NaN / Infinity are refused like JSON::Builder#number does; from_json does not coerce element types."""
    import numpy as np
    import pytest
    from ph_core_b200 import io
    cases = {0.1: "0.1", 2.0: "2.0", -1.5e-7: "-1.5e-7", 1e22: "1.0e+22", 5e-324: "5.0e-324", 1e14: "100000000000000.0",
             1e15: "1.0e+15", 1e-4: "0.0001", 1e-5: "1.0e-5", -0.0: "-0.0", 0.30000000000000004: "0.30000000000000004",
             123.456: "123.456"}
    for v, text in cases.items():
        assert io.format_float(np.float64(v)) == text
        assert float(io.format_float(np.float64(v))) == v
    assert io.format_float(np.float32(0.1)) == "0.1" and io.format_float(np.float32(16777216.0)) == "16777216.0"
    assert io.host_to_json(np.array([[0.5, 1e16]], np.float32)) == '{"shape":[1,2],"elements":[0.5,1.0e+16]}'
    assert io.host_to_yaml(np.array([True, False])) == "---\nshape: [2]\nelements: [true, false]\n"
    for bad in (np.nan, np.inf, -np.inf):
        with pytest.raises(ValueError):
            io.format_float(np.float64(bad))
    with pytest.raises(ValueError):
        io._typed_elements([1, 2.5], np.int32, "JSON")
    with pytest.raises(ValueError):
        io._typed_elements([1, 300], np.uint8, "JSON")
    with pytest.raises(ValueError):
        io._typed_elements([1, True], np.int32, "JSON")
    assert io._typed_elements([1, 2], np.float32, "JSON").tolist() == [1.0, 2.0]


def test_concatenate_push_wrap_oracle_and_the_host_shape_rule():
    """NArray.concatenate / push / wrap (src/n_array.cr:321-344, 666-750).  The reference holds no spec for them
    (`.wrap` is `pending`, spec/n_array_spec.cr:109), so the oracle's loop-for-loop restatement of
    `concatenate_to_slice` is checked against the independent numpy definition, and ph_concat_shape (the host half
    of the device form) against the oracle's `compatible?`, quirks included."""
    from ph_core_b200.narray import _concat_shape, DimensionError, CrIndexError
    rs = np.random.RandomState(5)
    cases = [([(2, 3), (4, 3)], 0), ([(2, 3), (2, 5), (2, 1)], 1), ([(2, 3, 4), (2, 1, 4)], 1), ([(2, 3, 4), (2, 3, 2)], 2),
             ([(5,), (3,)], 0), ([(0, 3), (2, 3)], 0), ([(2, 0), (2, 3)], 1), ([(3, 2, 2)] * 4, 0), ([(2, 3), (2, 3)], -1),
             ([(2, 3, 4), (2, 3, 4)], -2)]
    for shapes, axis in cases:
        arrs = [rs.randint(-50, 50, size=s).astype(np.int16) for s in shapes]
        want = np.concatenate(arrs, axis)
        got = O.concatenate(arrs, axis)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (shapes, axis)
        shape, ax = _concat_shape([list(s) for s in shapes], axis)
        assert shape == list(want.shape) and ax == axis % len(shapes[0])
    # `compatible?` compares `idx != axis` on the raw argument: a negative axis excludes nothing
    for shapes, axis, exc_o, exc_h in [([(2, 3), (2, 4)], -1, O.DimensionError, DimensionError),
                                       ([(2, 3), (3, 3)], 1, O.DimensionError, DimensionError),
                                       ([(2, 3), (2,)], 0, O.CrIndexError, CrIndexError),
                                       ([(2, 3), (2, 3)], 2, O.CrIndexError, CrIndexError)]:
        with pytest.raises(exc_o):
            O.concatenate([np.zeros(s) for s in shapes], axis)
        with pytest.raises(exc_h):
            _concat_shape([list(s) for s in shapes], axis)
    a, b = np.arange(6).reshape(2, 3), np.arange(10, 13).reshape(1, 3)
    assert O.push(a, [b]).tolist() == [[0, 1, 2], [3, 4, 5], [10, 11, 12]]
    assert O.push(a, [b, b], axis=0).shape == (4, 3)
    with pytest.raises(O.DimensionError):
        O.push(a, [np.zeros((1, 4))])
    assert O.wrap([a, a + 1, a + 2]).shape == (3, 2, 3) and O.wrap([a, a + 1])[1].tolist() == (a + 1).tolist()
    with pytest.raises(O.DimensionError):
        O.wrap([a, b])


def test_concat_shape_agrees_with_the_oracle_on_random_shapes():
    """ph_concat_shape vs oracle.shapes_compatible_except / concatenate on seeded random shape lists: same
    accept / reject decision (exception class included) and the same result shape."""
    from ph_core_b200.narray import _concat_shape, DimensionError, CrIndexError
    rs = np.random.RandomState(77)
    outcomes = {"ok": 0, "dim": 0, "idx": 0}
    for _ in range(600):
        rank = int(rs.randint(1, 5))
        base = [int(v) for v in rs.randint(0, 4, size=rank)]
        n = int(rs.randint(1, 5))
        axis = int(rs.randint(-rank - 1, rank + 1))
        shapes = []
        for k in range(n):
            sh = list(base)
            if rs.rand() < 0.5 and -rank <= axis < rank:
                sh[axis % rank] = int(rs.randint(0, 5))            # differ along the (canonical) axis only
            if rs.rand() < 0.15:
                sh[int(rs.randint(0, rank))] += 1                  # differ somewhere else
            if rs.rand() < 0.07 and k > 0:
                sh = sh[:-1]                                       # a shorter shape
            shapes.append(sh)
        try:
            want = ("ok", list(O.concatenate([np.zeros(s, np.int8) for s in shapes], axis).shape))
        except O.DimensionError:
            want = ("dim", None)
        except O.CrIndexError:
            want = ("idx", None)
        try:
            got = ("ok", _concat_shape(shapes, axis)[0])
        except DimensionError:
            got = ("dim", None)
        except CrIndexError:
            got = ("idx", None)
        assert got == want, (shapes, axis, got, want)
        outcomes[want[0]] += 1
    assert min(outcomes.values()) > 20, outcomes                   # every branch is exercised
