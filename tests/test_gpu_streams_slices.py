"""GPU: the round-2 additions of the array API -- descriptor-only each_slice, integer <=>, streams /
pinned buffers / asynchronous transfers and the row pipeline built from them, the position-weighted
checksum."""
import numpy as np
import pytest

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, DeviceView, rng
from oracle import ph_oracle as O
from gpu_util import assert_bits

pytestmark = pytest.mark.gpu


def test_each_slice_is_views_and_a_fold_over_them_is_the_axis_sum():
    """the reference's per-axis idiom (each_slice + `+`, src/multi_indexable.cr:742-786) moves
    no bytes for the slices themselves and reproduces sum(axis:) bit for bit (same fold order)."""
    lib = ph.load()
    a = (np.random.RandomState(3).rand(7, 33, 40) * 10 - 5).astype(np.float32)
    d = D.from_host(a)
    for axis in range(3):
        before = lib.ph_launch_count()
        sl = list(d.each_slice(axis))
        assert lib.ph_launch_count() == before                           # descriptors only: no launch, no copy
        assert len(sl) == a.shape[axis] and all(isinstance(s, DeviceView) and s._buf is d._buf for s in sl)
        for i in (0, a.shape[axis] - 1):
            assert_bits(sl[i].to_host(), np.ascontiguousarray(np.take(a, i, axis=axis)), f"slice {i} of axis {axis}")
        acc = sl[0]
        for s in sl[1:]:
            acc = acc + s
        assert_bits(acc.to_host(), O.reduce_axis(a, axis, "sum"), f"fold of each_slice({axis}) vs oracle")
        if axis < 2:                                                     # sum(axis:) keeps the fold order off the last axis
            assert_bits(acc.to_host(), d.sum(axis=axis).to_host(), f"fold of each_slice({axis})")
        else:                                                            # last axis: lanes share a row (tolerance class)
            np.testing.assert_allclose(acc.to_host(), d.sum(axis=axis).to_host(), rtol=1e-4, atol=1e-4)
    # slices of a vector are 1-element arrays of shape [1]; writes through a slice reach the source
    v = D.from_host(np.arange(5, dtype=np.int32))
    assert [s.to_host().tolist() for s in v.each_slice(0)] == [[0], [1], [2], [3], [4]]
    m = D.from_host(np.zeros((3, 4), np.int32))
    for i, s in enumerate(m.each_slice(1)):
        s.set_chunk([ph.ALL], i + 1)
    assert m.to_host().tolist() == [[1, 2, 3, 4]] * 3
    with pytest.raises(ph.CrIndexError):
        list(d.each_slice(3))


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.uint8, np.int16])
def test_spaceship_operator_on_integers(dtype):
    """`<=>` of the operator list (src/multi_indexable.cr:981): Int#<=> -> -1 / 0 / 1 as Int32."""
    rs = np.random.RandomState(9)
    info = np.iinfo(dtype)
    a = rs.randint(max(info.min, -50), min(info.max, 50), size=(37, 21)).astype(dtype)
    b = rs.randint(max(info.min, -50), min(info.max, 50), size=(37, 21)).astype(dtype)
    a[0, :3] = [info.min, info.max, 0]
    b[0, :3] = [info.max, info.min, 0]
    want = (a.astype(object) > b.astype(object)).astype(np.int32) - (a.astype(object) < b.astype(object)).astype(np.int32)
    got = D.from_host(a).cmp(D.from_host(b))
    assert got.dtype == np.int32
    assert_bits(got.to_host(), want.astype(np.int32), "<=>")
    assert_bits(D.from_host(a).cmp(3).to_host(), ((a > 3).astype(np.int32) - (a < 3).astype(np.int32)), "<=> scalar")
    assert_bits(D.from_host(a).view(rng(None, None, -1), rng(0, None, 2)).cmp(D.from_host(b).view(rng(None, None, -1), rng(0, None, 2))).to_host(),
                np.ascontiguousarray(want[::-1, ::2]).astype(np.int32), "<=> on views")
    with pytest.raises(TypeError):
        D.from_host(a.astype(np.float32)).cmp(D.from_host(b.astype(np.float32)))       # Float#<=> is nilable
    with pytest.raises(ph.ShapeError):
        D.from_host(a).cmp(D.from_host(b[:5]))


def test_streams_async_transfers_and_the_row_pipeline():
    """the chunked host -> device -> host pipeline is a public entry written with the array
    API (pinned buffers, from_host_async / to_host_async, library streams)."""
    rs = np.random.RandomState(1)
    rows, cols = 1000, 768                                              # 1000 rows / 8 chunks: a ragged last chunk
    a, c = (rs.rand(rows, cols) * 2 - 1).astype(np.float32), (rs.rand(rows, cols) * 2 - 1).astype(np.float32)
    b = (rs.rand(1, cols) * 2 - 1).astype(np.float32)
    t, _ = O.ewise_broadcast("*", a, b)
    want, _ = O.ewise("+", t, c)
    a_pin, b_pin, c_pin = ph.pinned_from(a), ph.pinned_from(b), ph.pinned_from(c)
    out = ph.pinned_empty(a.shape, np.float32)
    pipe = ph.pipeline.RowPipeline(chunks=8, streams=3)
    for _ in range(3):                                                   # streams and pool blocks are reused
        out[...] = 0
        pipe.map_rows(lambda x, z, y: x.broadcast_op("*", y) + z, rows=[a_pin, c_pin], out=out, shared=[b_pin])
        assert_bits(np.array(out), want, "pipelined a*b+c")
    # the tapered schedule of the bench's e2e leg (few large chunks, the last one cut into halves): same bits
    taper = ph.pipeline.RowPipeline(chunks=4, taper=7, ramp=5)
    out[...] = 0
    taper.map_rows(lambda x, z, y: x.broadcast_op("*", y) + z, rows=[a_pin, c_pin], out=out, shared=[b_pin])
    assert_bits(np.array(out), want, "pipelined a*b+c, tapered chunks")
    taper.close()
    # data-dependent errors of a pipelined step surface at its synchronising end
    ia = ph.pinned_from(np.full((64, 8), 2**31 - 1, np.int32))
    io = ph.pinned_empty((64, 8), np.int32)
    with pytest.raises(ph.CrOverflowError):
        pipe.map_rows(lambda x: x + 1, rows=[ia], out=io)
    # an expression that raises half way leaves the pipeline usable (the queued chunks drain before anything is released)
    calls = {"n": 0}
    def flaky(x, z, y):
        calls["n"] += 1
        if calls["n"] == 3:
            raise ValueError("third chunk")
        return x.broadcast_op("*", y) + z
    with pytest.raises(ValueError):
        pipe.map_rows(flaky, rows=[a_pin, c_pin], out=out, shared=[b_pin])
    ph.sync()
    out[...] = 0
    pipe.map_rows(lambda x, z, y: x.broadcast_op("*", y) + z, rows=[a_pin, c_pin], out=out, shared=[b_pin])
    assert_bits(np.array(out), want, "pipeline reused after a failed call")
    pipe.close()
    # explicit streams: two independent chains, ordered by wait(), joined by sync()
    s1, s2 = ph.Stream(), ph.Stream()
    with s1:
        x = D.from_host_async(a_pin) * 2.0
    s2.wait(s1)
    with s2:
        (x + 1.0).to_host_async(out)
    ph.narray.main_stream_wait(s2)
    ph.sync()
    assert_bits(np.array(out), (a * np.float32(2.0)) + np.float32(1.0), "two streams")
    s1.close(); s2.close()
    with pytest.raises(ph.ShapeError):
        D.from_host(a).to_host_async(ph.pinned_empty((3, 3), np.float32))


def test_checksum64_is_position_sensitive_and_adds_over_shards():
    rs = np.random.RandomState(2)
    a = rs.randint(0, 2**31, size=(64, 256)).astype(np.int64)
    d = D.from_host(a)
    words = a.reshape(-1).astype(object)
    want = sum(int(w) * (2 * i + 1) for i, w in enumerate(words)) % (1 << 64)
    assert d.checksum64() == want
    parts = [D.from_host(a[r0:r1]).checksum64(r0 * 256) for r0, r1 in ((0, 10), (10, 41), (41, 64))]
    assert sum(parts) % (1 << 64) == want                              # shards add up with their global word offsets
    swapped = a.copy()
    swapped[[3, 40]] = swapped[[40, 3]]
    assert D.from_host(swapped).checksum64() != want                   # same words elsewhere: a different value
    f = D.from_host(np.arange(1 << 20, dtype=np.float32))
    assert f.checksum64() == f.clone().checksum64() != (f + 0.0).view(rng(None, None, -1)).to_narr().checksum64()
