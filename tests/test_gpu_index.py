"""T2 parity (GPU): strided gather / scatter / fill / mask store / views vs the oracle, bit-exact.
Reference: NArray#unsafe_fetch_chunk / unsafe_set_chunk / []=(mask) src/n_array.cr:450-551,
View / MutableView src/view.cr, src/mutable_view.cr, transforms src/view_util/transforms.cr."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D
from oracle import ph_oracle as O
from oracle.ph_oracle import rng, R
from gpu_util import assert_bits
import test_oracle_goldens as G


def stock():
    return np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32)


def test_fetch_chunk_goldens():
    """spec/n_array_spec.cr:211-229"""
    d = D.from_host(stock())
    assert d[1, rng(0, 2, 2)].to_host().tolist() == [3, 5]
    assert d[-2, rng(-1, 0)].to_host().tolist() == [2, 1, 0]
    e = d[rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)]
    assert e.shape == [0, 0] and e.to_host().shape == (0, 0)
    assert d.get(1, 1) == 4 and d.get([-1, -1]) == 5
    assert d[rng(None, None), 1].to_host().tolist() == [1, 4]          # README.md:46


def test_set_chunk_goldens():
    """spec/n_array_spec.cr:238-295"""
    d = D.from_host(stock()); d[1, rng(0, 2, 2)] = D.from_host(np.array([6, 7], np.int32))
    assert d.to_host().tolist() == [[0, 1, 2], [6, 4, 7]]
    d = D.from_host(stock()); d[-2, rng(-1, 0)] = D.from_host(np.array([6, 7, 8], np.int32))
    assert d.to_host().tolist() == [[8, 7, 6], [3, 4, 5]]
    empty = ph.make_region([rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)], [2, 3])
    d = D.from_host(stock()); d.unsafe_set_chunk(empty, D.from_host(np.zeros(0, np.int32)))     # the spec calls the unsafe form
    assert d.to_host().tolist() == stock().tolist()
    with pytest.raises(ph.ShapeError):                                  # set_chunk checks compatible_shapes? ([0] vs [0, 0])
        d[rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)] = D.from_host(np.zeros(0, np.int32))
    d = D.from_host(stock()); d[1, rng(0, 2, 2)] = 6
    assert d.to_host().tolist() == [[0, 1, 2], [6, 4, 6]]
    d = D.from_host(stock()); d[-2, rng(-1, 0)] = 6
    assert d.to_host().tolist() == [[6, 6, 6], [3, 4, 5]]
    d = D.from_host(stock()); d[rng(0, 0, exclusive=True), rng(0, 0, exclusive=True)] = 6
    assert d.to_host().tolist() == stock().tolist()


def test_multi_writable_goldens():
    """spec/multi_writable_spec.cr:14-41, 67-93"""
    base = np.arange(12, dtype=np.int64).reshape(3, 4)
    d = D.from_host(base); d[rng(1, None), rng(1, None)] = 10
    exp = base.reshape(-1).copy(); exp[5:8] = 10; exp[9:12] = 10
    assert d.to_host().reshape(-1).tolist() == exp.tolist()
    d = D.from_host(base); d[rng(1, None), rng(1, None)] = D.from_host(np.arange(10, 16).reshape(2, 3))
    exp = base.reshape(-1).copy(); exp[5:8] = [10, 11, 12]; exp[9:12] = [13, 14, 15]
    assert d.to_host().reshape(-1).tolist() == exp.tolist()
    with pytest.raises(ph.ShapeError):
        d[rng(2, None), rng(2, None)] = D.from_host(np.arange(10, 16).reshape(2, 3))
    d = D.from_host(base); d.set_element([-1, -2], 77)
    assert d.to_host()[2, 2] == 77
    for bad in ([10, 10], [-10, -10]):
        with pytest.raises(ph.CrIndexError):
            d.set_element(bad, 1)
    # trailing-ones compatibility (shape_util.cr:6-32): a [2,3,1] source into a [2,3] region
    d = D.from_host(base); d[rng(1, None), rng(1, None)] = D.from_host(np.arange(20, 26).reshape(2, 3, 1))
    assert d.to_host()[1:, 1:].reshape(-1).tolist() == list(range(20, 26))


def test_mask_goldens():
    """spec/n_array_spec.cr:297-333"""
    mask = D.from_host(np.array([[True, False, True], [False, True, False]]))
    d = D.from_host(stock()); d[mask] = 6
    assert d.to_host().tolist() == [[6, 1, 6], [3, 6, 5]]
    d = D.from_host(stock()); d[mask] = D.from_host(stock()) + 10
    assert d.to_host().tolist() == [[10, 1, 12], [3, 14, 5]]
    with pytest.raises(ph.DimensionError):
        d[D.from_host(np.zeros((3, 2), np.bool_))] = 6
    assert d[mask] is d                                                # multi_indexable.cr:479-481
    # `narr[mask] += 1` expands to narr[mask] = narr[mask] + 1
    d = D.from_host(stock()); d[mask] = d[mask] + 1
    assert d.to_host().tolist() == [[1, 1, 3], [3, 5, 5]]


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
def test_mask_store_large(dtype):
    rs = np.random.RandomState(4)
    for n in [1, 31, 4096, 100003]:
        a = (rs.rand(n) * 100).astype(dtype)
        v = (rs.rand(n) * 100).astype(dtype)
        m = rs.rand(n) < 0.3
        m[: n // 3] = False                                           # long untouched runs
        want = a.copy(); O.mask_set(want, m, v)
        d = D.from_host(a); d[D.from_host(m)] = D.from_host(v)
        assert_bits(d.to_host(), want, f"mask array {n}")
        want = a.copy(); want[m] = dtype(7)
        d = D.from_host(a); d[D.from_host(m)] = 7
        assert_bits(d.to_host(), want, f"mask scalar {n}")


@pytest.mark.parametrize("shape", [[2, 3, 4], [3, 5], [3, 4], [1], [1, 1, 1]])
@pytest.mark.parametrize("drop", [True, False])
def test_conformance_regions(shape, drop):
    """multi_indexable_tester.cr:416-456 replayed on the device array."""
    n = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)
    d = D.from_host(n)
    for lit in G.valid_regions(shape):
        want = O.fetch_chunk(n, O.IndexRegion.new(lit, shape, drop))
        got = d.get_chunk(lit, drop)
        assert got.shape == list(want.shape)
        assert_bits(got.to_host(), want, f"{lit}")
    for lit, exc in G.invalid_regions(shape):
        want = {O.CrIndexError: ph.CrIndexError, O.DimensionError: ph.DimensionError}[exc]
        with pytest.raises(want):
            d.get_chunk(lit, drop)
    for c in np.ndindex(*shape):                                     # get over all coords (:374-387)
        assert d.get(*c) == n[c]


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
def test_gather_scatter_shapes(dtype):
    """Gather + scatter over regions that exercise every kernel: contiguous, row-strided,
    column-strided, reversed, tiny inner extent, odd offsets."""
    rs = np.random.RandomState(8)
    shape = [37, 50, 24]
    n = (rs.rand(*shape) * 200).astype(dtype)
    d = D.from_host(n)
    lits = [
        [rng(None, None)],
        [rng(0, None, 2), rng(None, -1)],
        [rng(None, None), rng(0, None, 2)],
        [rng(None, None), rng(None, None), rng(0, None, 2)],
        [rng(None, None, -1), rng(None, None, -1), rng(None, None, -1)],
        [rng(5, 30, 3), 7, rng(20, 2, -3)],
        [3, rng(1, 48), rng(1, 22)],
        [rng(1, 36), rng(1, 49), 5],
        [rng(2, 2), rng(None, None), rng(None, None)],
        [rng(36, 0, -5), rng(49, 1, -7), rng(23, 0, -1)],
    ]
    for lit in lits:
        reg = O.IndexRegion.new(lit, shape)
        want = O.fetch_chunk_fast(n, reg)
        assert_bits(d.get_chunk(lit).to_host(), want, f"gather {lit}")
        src = (rs.rand(*want.shape) * 200).astype(dtype)
        exp = n.copy(); O.set_chunk_fast(exp, reg, src)
        t = D.from_host(n); t.set_chunk(lit, D.from_host(src))
        assert_bits(t.to_host(), exp, f"scatter {lit}")
        exp = n.copy(); O.set_chunk_fast(exp, reg, dtype(9))
        t = D.from_host(n); t.set_chunk(lit, 9)
        assert_bits(t.to_host(), exp, f"fill {lit}")


@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.float64])
def test_views_and_transposes(dtype):
    """View chains -> one descriptor -> one gather (view.cr:43-126); permute/reverse/reshape
    copies (multi_indexable.cr:795-803); MutableView scatter (mutable_view.cr:16-18)."""
    rs = np.random.RandomState(3)
    n = (rs.rand(45, 70, 33) * 250).astype(dtype)
    d = D.from_host(n)
    assert_bits(d.permute().to_host(), np.ascontiguousarray(np.transpose(n)), "permute default")
    for order in [(0, 2, 1), (1, 0, 2), (2, 0, 1), (1, 2, 0)]:
        assert_bits(d.permute(*order).to_host(), np.ascontiguousarray(np.transpose(n, order)), f"permute {order}")
    assert_bits(d.reverse().to_host(), np.ascontiguousarray(n[::-1, ::-1, ::-1]), "reverse")
    assert_bits(d.view().reshape(45 * 70, 33).to_narr().to_host(), n.reshape(45 * 70, 33), "reshape view")
    r = d.reshape(70, 45, 33)                                        # aliases the buffer (n_array.cr:429-433)
    r[0, 0, 0] = 123
    assert d.get(0, 0, 0) == dtype(123)
    n = d.to_host()
    # chain: region -> permute -> reverse -> region, against the oracle's transform chain
    ov = O.View(n).view([rng(2, 40, 2), rng(None, None), rng(30, 3, -3)]).permute([2, 0, 1]).reverse().view([rng(1, 7), rng(None, None), rng(0, None, 5)])
    dv = d.view(rng(2, 40, 2), rng(None, None), rng(30, 3, -3)).permute(2, 0, 1).reverse().view(rng(1, 7), rng(None, None), rng(0, None, 5))
    assert dv.shape == ov.shape
    assert_bits(dv.to_narr().to_host(), ov.to_narr(), "view chain")
    # non-contiguous reshape materialises first
    ov2 = O.View(n).view([rng(0, None, 2), rng(None, None), rng(None, None)]).permute([1, 0, 2])
    dv2 = d.view(rng(0, None, 2), rng(None, None), rng(None, None)).permute(1, 0, 2)
    assert_bits(dv2.reshape(70 * 23, 33).to_narr().to_host(), ov2.to_narr().reshape(70 * 23, 33), "reshape after permute")
    # MutableView: transposed scatter
    z = D.from_host(np.zeros((64, 48), dtype))
    src = (rs.rand(48, 64) * 250).astype(dtype)
    z.mutable_view().permute(1, 0)[rng(None, None), rng(None, None)] = D.from_host(src)
    assert_bits(z.to_host(), np.ascontiguousarray(src.T), "transposed scatter")
    z.mutable_view().permute(1, 0)[rng(1, 2), 0] = 99
    h = z.to_host(); assert h[0, 1] == dtype(99) and h[0, 2] == dtype(99)
    with pytest.raises(ph.CrIndexError):
        d.view().permute(0, 1, 3)
    with pytest.raises(ph.ShapeError):
        d.view().reshape(5, 5)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(1000, 1030), (257, 33), (64, 64), (8, 4099), (3, 5, 1024, 40)])
def test_transpose_tile_kernel(dtype, shape):
    """The shared-memory tile transpose on ragged 2-D tiles and batched (rank-4) permutes."""
    n = np.arange(int(np.prod(shape)), dtype=dtype).reshape(shape)
    d = D.from_host(n)
    order = tuple(range(len(shape) - 2)) + (len(shape) - 1, len(shape) - 2)
    assert_bits(d.permute(*order).to_host(), np.ascontiguousarray(np.transpose(n, order)), "tile transpose")
    # transposed + reversed + strided source
    v = d.view().permute(*order).reverse()
    assert_bits(v.to_narr().to_host(), np.ascontiguousarray(np.transpose(n, order)[tuple(slice(None, None, -1) for _ in shape)]), "rev transpose")


def test_slices_tile_equals():
    """README.md:63-64, multi_indexable.cr:806-827, n_array.cr:440-447"""
    b = np.array([[0, 1, 2], [10, 11, 12]], np.int32)
    d = D.from_host(b)
    assert [s.to_host().tolist() for s in d.slices(axis=1)] == [[0, 10], [1, 11], [2, 12]]
    assert [s.to_host().tolist() for s in d.slices()] == [[0, 1, 2], [10, 11, 12]]
    unit = D.from_host(np.array([[1, 2], [3, 4]], np.int32))
    assert unit.tile([2, 3]).to_host().tolist() == [[1, 2, 1, 2, 1, 2], [3, 4, 3, 4, 3, 4]] * 2
    big = np.arange(35, dtype=np.float32).reshape(5, 7)
    assert_bits(D.from_host(big).tile([3, 2]).to_host(), O.tile(big, [3, 2]), "tile")
    assert d.equals(D.from_host(b)) and not d.equals(D.from_host(b + 1)) and not d.equals(unit)
    c = d.clone(); c[0, 0] = 5
    assert d.get(0, 0) == 0 and c.get(0, 0) == 5                       # clone deep-copies
    assert D.fill([3, 2], 7, np.int64).to_host().tolist() == [[7, 7]] * 3


def test_large_strided_configs_small_scale():
    """BASELINE config 2 shapes at 1/16 scale: narr[0..2.., ..-1], narr[.., 0..2..], reversed,
    transposed copy, transposed scatter, and the literal narr[..2, ..-1] (rows 0..2 inclusive)."""
    nrow = ncol = 1024
    n = np.arange(nrow * ncol, dtype=np.float64).reshape(nrow, ncol)     # value = flat index
    d = D.from_host(n)
    assert_bits(d[rng(0, None, 2), rng(None, -1)].to_host(), n[0::2, :], "rows strided")
    assert_bits(d[rng(None, None), rng(0, None, 2)].to_host(), np.ascontiguousarray(n[:, 0::2]), "cols strided")
    assert_bits(d[rng(None, None, -1), rng(None, None, -1)].to_host(), np.ascontiguousarray(n[::-1, ::-1]), "reversed")
    assert_bits(d.permute().to_host(), np.ascontiguousarray(n.T), "transposed copy")
    lit = d[rng(None, 2), rng(None, -1)]
    assert lit.shape == [3, ncol]
    assert_bits(lit.to_host(), n[0:3, :], "literal ..2, ..-1")
    z = D.fill([nrow, ncol], 0.0, np.float64)
    z.mutable_view().permute()[rng(None, None), rng(None, None)] = d
    assert_bits(z.to_host(), np.ascontiguousarray(n.T), "transposed scatter")


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.uint8, np.int16])
def test_long_single_row_strided_copies(dtype):
    """One long strided row (axes of a reversed contiguous array coalesce into a single axis): the
    copy folds it into rows of L elements plus a remainder; gathers and scatters, steps -1, 2, -3."""
    n = 5 * 16384 + 77                                                   # not a multiple of any row length
    a = (np.arange(n) % 251).astype(dtype)
    d = D.from_host(a)
    assert_bits(d[rng(None, None, -1)].to_host(), np.ascontiguousarray(a[::-1]), "1-D reversed")
    assert_bits(d[rng(0, None, 2)].to_host(), np.ascontiguousarray(a[0::2]), "1-D step 2")
    assert_bits(d[rng(n - 1, 0, -3)].to_host(), np.ascontiguousarray(a[n - 1::-3]), "1-D step -3")
    m = a[:n - 77].reshape(5, 16384)                                     # 2-D, both axes reversed -> one axis of stride -1
    assert_bits(D.from_host(m)[rng(None, None, -1), rng(None, None, -1)].to_host(), np.ascontiguousarray(m[::-1, ::-1]), "2-D reversed")
    z = D.fill([n], 0, dtype)                                            # scatter through a reversed destination
    z[rng(None, None, -1)] = d
    assert_bits(z.to_host(), np.ascontiguousarray(a[::-1]), "reversed scatter")
    z2 = D.fill([2 * n], 0, dtype)
    z2[rng(1, None, 2)] = d
    want = np.zeros(2 * n, dtype); want[1::2] = a
    assert_bits(z2.to_host(), want, "step-2 scatter")


def test_get_available_and_optional_chunk():
    """multi_indexable.cr:397-413 (get_available), :313-318 (has_region?), :540-546 ([]?)."""
    n = np.arange(6, dtype=np.int32).reshape(2, 3) + 1                  # [[1,2,3],[4,5,6]]
    d = D.from_host(n)
    assert d.get_available([rng(1, 5), 1]).to_host().tolist() == [5]    # the doc example
    with pytest.raises(ph.CrIndexError):
        d.get_chunk([rng(1, 5), 1])
    assert d.has_region([rng(0, None), rng(1, 2)]) and not d.has_region([rng(1, None), rng(10, 12)])
    assert d.get_chunk_or_none([rng(1, None), rng(10, 12)]) is None
    assert d.get_chunk_or_none([rng(0, None), rng(1, 2)]).to_host().tolist() == [[2, 3], [5, 6]]
    big = np.arange(7 * 9, dtype=np.float32).reshape(7, 9)
    db = D.from_host(big)
    for lit in ([rng(2, 30), rng(0, 100, 2)], [rng(20, 3, -3), rng(4, None)], [rng(5, 5), rng(8, 40)]):
        want = O.fetch_chunk(big, O.IndexRegion.new_trimmed(lit, list(big.shape), bound_shape=list(big.shape)))
        assert_bits(db.get_available(lit).to_host(), want, f"get_available {lit}")
    assert_bits(db.match(5.0).to_host(), big == 5.0, "=~")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8])
def test_short_last_outer_axis_is_reordered(dtype):
    """Plans whose last outer axis is short (tile counts of 2, a [N, 3, C] slice) walk their LONGEST
    outer axis inside a block instead (map_kernels.cuh); results cannot depend on that choice."""
    rs = np.random.RandomState(12)
    src = (rs.rand(300, 264) * 200).astype(dtype)
    for counts in ([2, 2], [3, 1], [1, 5], [2, 3]):
        assert_bits(D.from_host(src).tile(counts).to_host(), np.tile(src, counts), f"tile {counts}")
    cube = (rs.rand(70, 3, 520) * 200).astype(dtype)
    d = D.from_host(cube)
    # a strided rank-3 operand (axes cannot merge) with a last outer axis of extent 3, elementwise and gathered
    sub = d[rng(None, None), rng(None, None), rng(0, 511)]
    assert_bits(sub.to_host(), cube[:, :, :512], "gather [70,3,512]")
    v = d.view(rng(None, None), rng(None, None), rng(0, 511))
    if np.dtype(dtype).kind == "f":
        want, _ = O.ewise("*", cube[:, :, :512].copy(), cube[:, :, 8:].copy())
        got = v * d.view(rng(None, None), rng(None, None), rng(8, None))
        assert_bits(got.to_host(), want, "strided * strided, outer [70,3]")
    # scatter into a region whose last outer axis has extent 2
    dst = D.fill([40, 2, 300], 0, dtype)
    dst[rng(None, None), rng(None, None), rng(4, 259)] = D.from_host(np.ascontiguousarray(cube[:40, :2, :256]))
    want = np.zeros((40, 2, 300), dtype); want[:, :, 4:260] = cube[:40, :2, :256]
    assert_bits(dst.to_host(), want, "scatter [40,2,256]")
    # row-vector operand with a short last outer axis: the ROWVEC detection is order-independent
    if np.dtype(dtype).kind == "f":
        rowv = D.from_host(cube[:1, :1, :512].copy())
        want, _ = O.ewise_broadcast("+", cube[:, :, :512].copy(), cube[:1, :1, :512].copy())
        assert_bits(v.broadcast_op("+", rowv).to_host(), want, "rowvec + strided")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8])
def test_slices_along_every_axis_are_one_batched_copy(dtype):
    """MultiIndexable#slices (src/multi_indexable.cr:776-786): per index the chunk self[.., i, ..]; the device
    produces all of them with ONE permuting copy and hands out ranges of its buffer.  They must equal the
    oracle's each_slice, be independent of each other and of the source, and cost one launch."""
    rs = np.random.RandomState(21)
    host = (rs.rand(5, 36, 130) * 200).astype(dtype)
    d = D.from_host(host)
    for axis in range(3):
        before = ph.load().ph_launch_count()
        got = d.slices(axis)
        assert ph.load().ph_launch_count() - before == 1
        want = list(O.each_slice(host, axis))
        assert len(got) == len(want) == host.shape[axis]
        for g, w in zip(got, want):
            assert g.shape == list(w.shape)
            assert_bits(g.to_host(), w, f"slice along {axis}")
    sl = d.slices(1)
    sl[3][rng(None, None), rng(None, None)] = 7                      # writing one slice ...
    assert_bits(sl[2].to_host(), host[:, 2, :], "neighbour slice untouched")
    assert_bits(sl[4].to_host(), host[:, 4, :], "neighbour slice untouched")
    assert_bits(d.to_host(), host, "source untouched")               # ... touches neither its neighbours nor the source
    assert (sl[3].to_host() == 7).all()
    keep = sl[5]
    del sl, got                                                      # the shared buffer lives as long as any slice does
    assert_bits(keep.clone().to_host(), host[:, 5, :], "slice outlives its siblings")
    one_d = D.from_host(np.arange(4, dtype=dtype))
    assert [s.to_host().tolist() for s in one_d.slices()] == [[0], [1], [2], [3]]       # all axes dropped -> [1]
    assert [s.shape for s in D.fill([3, 0, 2], 1, dtype).slices(0)] == [[0, 2]] * 3 and D.fill([3, 0, 2], 1, dtype).slices(1) == []
    with pytest.raises(ph.CrIndexError):
        d.slices(3)
    strided = d.view(rng(None, None), rng(0, None, 5), rng(None, None, -1))         # slices of a VIEW
    for g, w in zip(strided.slices(2), O.each_slice(np.ascontiguousarray(host[:, ::5, ::-1]), 2)):
        assert_bits(g.to_host(), w, "slices of a strided view")


@pytest.mark.parametrize("shape", [[0], [5, 0, 2], [3, 0, 0, 1]])
def test_empty_arrays_through_every_entry_point(shape):
    """The reference's empties (spec/multi_indexable_spec.cr: shapes [0], [5,0,2], [3,0,0,1]): every entry point
    accepts a zero-element array, launches nothing that could fault, and keeps the shape rules."""
    host = np.zeros(shape, np.float32)
    failures = []

    def check(name, fn):
        try:
            assert fn(), "returned False"
        except Exception as e:                                           # collect, report all at once
            failures.append(f"{name}: {type(e).__name__}: {e}")

    d = D.from_host(host)
    D.take_flags()                                                       # clear whatever earlier tests left behind
    check("size/empty", lambda: d.size == 0 and d.empty() and not d.scalar())
    check("to_host", lambda: d.to_host().shape == tuple(shape))
    check("clone", lambda: d.clone().shape == shape)
    check("binary", lambda: (d * d + d).shape == shape)
    check("scalar both sides", lambda: (2 * d - 1).shape == shape)
    check("unary", lambda: (-d).shape == shape)
    check("compare", lambda: (d > d).shape == shape and (d > d).dtype == np.dtype(np.bool_) and d.eq(0).shape == shape)
    check("equals", lambda: d.equals(D.from_host(host)))
    check("mask store", lambda: _runs(lambda: d.set_mask(d > d, 1.0), lambda: d.set_mask(d > d, d)))
    # a literal `..` on a zero-length axis raises IndexError in the reference too (range_syntax.cr:120-122:
    # last = bound - 1 = -1); the whole of an empty array is IndexRegion.cover (index_region.cr:232-238)
    check("`..` on an empty axis raises IndexError", lambda: _raises(ph.CrIndexError, lambda: d.get_chunk([ph.ALL] * len(shape))))
    cover = ph.cover_region(shape)
    check("cover gather", lambda: d.unsafe_fetch_chunk(cover).shape == shape)
    check("fill region", lambda: _runs(lambda: d.unsafe_set_chunk(cover, 3.0)))
    check("scatter", lambda: _runs(lambda: d.unsafe_set_chunk(cover, D.from_host(host))))
    check("view chain", lambda: d.view().permute().reverse().to_narr().shape == shape[::-1])
    check("reshape", lambda: d.reshape([0]).shape == [0] and d.flatten().shape == [0])
    check("sum is zero", lambda: d.sum() == 0)
    for name in ("min", "max", "argmax"):
        def raises(name=name):
            try:
                getattr(d, name)()
            except ph.CrEmptyError:
                return True
            return False
        check(f"{name} raises EmptyError", raises)
    check("first raises ShapeError", lambda: _raises(ph.ShapeError, d.first))
    check("tile", lambda: d.tile([2] * len(shape)).shape == [2 * s for s in shape])
    check("slices", lambda: [s.shape for s in d.slices(0)] == [shape[1:] or [1]] * shape[0])
    check("fused mul_add", lambda: d.mul_add(d, d).shape == shape)
    if len(shape) > 1:
        ax = shape.index(0)
        check("axis sum over the empty axis", lambda: (d.sum(axis=ax).to_host() == 0).all()
              and d.sum(axis=ax).shape == ([s for i, s in enumerate(shape) if i != ax] or [1]))
        check("axis max over the empty axis raises", lambda: _raises(ph.CrEmptyError, lambda: d.max(axis=ax)))
        other = 0 if ax != 0 else len(shape) - 1
        check("axis sum over another axis", lambda: d.sum(axis=other).size == 0)
    check("no arithmetic flags", lambda: D.take_flags() == 0)
    assert failures == [], "\\n".join(failures)


def _raises(exc, fn):
    try:
        fn()
    except exc:
        return True
    return False


def _runs(*fns):
    """True once every call has returned (an exception propagates to the caller's collector)."""
    for fn in fns:
        fn()
    return True


def test_concatenate_push_wrap_on_the_device():
    """NArray.concatenate / #concatenate / #push / << / NArray.wrap (src/n_array.cr:321-344, 666-750) on device
    arrays and views: one strided copy per input, bit-exact vs the oracle's restatement of concatenate_to_slice."""
    rs = np.random.RandomState(9)
    lib = ph.load()
    for dtype in (np.float32, np.int64, np.uint8):
        for shapes, axis in [([(2, 3), (4, 3)], 0), ([(2, 3), (2, 5), (2, 1)], 1), ([(3, 4, 5), (3, 1, 5), (3, 7, 5)], 1),
                             ([(3, 4, 5), (3, 4, 2)], 2), ([(5,), (3,)], 0), ([(0, 3), (2, 3)], 0), ([(2, 0), (2, 3)], 1),
                             ([(64, 96), (64, 96)], -1), ([(300, 129), (41, 129), (1, 129)], 0)]:
            arrs = [rs.randint(0, 200, size=s).astype(dtype) for s in shapes]
            devs = [D.from_host(a) for a in arrs]
            before = lib.ph_launch_count()
            got = D.concatenate(*devs, axis=axis)                       # class form
            launches = lib.ph_launch_count() - before
            assert launches <= sum(1 for a in arrs if a.size), (shapes, launches)     # one copy per non-empty input
            assert_bits(got.to_host(), O.concatenate(arrs, axis), f"concatenate {shapes} axis={axis}")
            assert_bits(devs[0].concatenate(*devs[1:], axis=axis).to_host(), O.concatenate(arrs, axis), "instance form")
    a = rs.rand(6, 8).astype(np.float64)
    d = D.from_host(a)
    # views join like arrays (the source is read through its descriptor): transposed and strided pieces
    got = D.concatenate(d.view().permute(), d.view(rng(None, None, 2), rng(None, None)).permute(), axis=1)
    assert_bits(got.to_host(), O.concatenate([a.T, a[::2].T], 1), "views")
    with pytest.raises(ph.DimensionError):
        D.concatenate(d, D.from_host(a[:, :7]), axis=0)
    with pytest.raises(ph.DimensionError):
        D.concatenate(d, D.from_host(a[:5]), axis=-1)                    # a negative axis excludes nothing (compatible?)
    with pytest.raises(ph.CrIndexError):
        D.concatenate(d, d, axis=2)
    with pytest.raises(TypeError):
        D.concatenate(d, D.from_host(a.astype(np.float32)))
    # push / << grow shape[0] in place; reshape aliases made before keep the OLD buffer (like the reference's new Slice)
    p = D.from_host(a)
    alias = p.reshape(8, 6)
    assert (p << D.from_host(a[:2] + 1)) is p and p.shape == [8, 8]
    assert_bits(p.to_host(), O.push(a, [a[:2] + 1]), "push")
    assert_bits(alias.to_host(), a.reshape(8, 6), "an alias made before push keeps the old contents")
    p.push(D.from_host(a[:1]), D.from_host(a[3:]), axis=0)
    assert_bits(p.to_host(), O.push(O.push(a, [a[:2] + 1]), [a[:1], a[3:]]), "push of two")
    with pytest.raises(ph.DimensionError):
        p.push(D.from_host(a[:, :3]))
    w = D.wrap(d, d + 1.0, d.view().reverse())
    assert_bits(w.to_host(), O.wrap([a, a + 1.0, a[::-1, ::-1]]), "wrap")
    with pytest.raises(ph.DimensionError):
        D.wrap(d, D.from_host(a[:5]))


def test_get_chunk_by_corner_and_shape():
    """MultiIndexable#get_chunk(coord, region_shape) (src/multi_indexable.cr:369-395), the source's own examples first."""
    n = np.arange(1, 10, dtype=np.int32).reshape(3, 3)
    d = D.from_host(n)
    assert d.get_chunk([1, 0], [1, 3]).to_host().tolist() == [[4, 5, 6]]
    with pytest.raises(ph.ShapeError):
        d.get_chunk([1, 0], [10, 10])
    with pytest.raises(ph.DimensionError):
        d.get_chunk([0], [1])
    with pytest.raises(ph.DimensionError):
        d.get_chunk([0, 0], [1])
    with pytest.raises(ph.CrArgumentError):
        d.get_chunk([-1, 0], [1, 1])
    with pytest.raises(ph.CrArgumentError):
        d.get_chunk([0, 0], [1, -1])
    rs = np.random.RandomState(4)
    a = rs.rand(7, 9, 11).astype(np.float32)
    da = D.from_host(a)
    for coord, shape in [([0, 0, 0], [7, 9, 11]), ([2, 3, 4], [3, 1, 7]), ([6, 8, 10], [1, 1, 1]), ([1, 2, 3], [0, 4, 2]), ([7, 9, 11], [0, 0, 0])]:
        got = da.get_chunk(coord, shape)
        assert got.shape == shape
        assert_bits(got.to_host(), O.get_chunk_at(a, coord, shape), f"get_chunk({coord}, {shape})")
    v = da.view().permute()                                           # the same through a view
    assert_bits(v.get_chunk([3, 1, 2], [5, 4, 3]).to_host(), O.get_chunk_at(a.transpose(2, 1, 0), [3, 1, 2], [5, 4, 3]), "view")
