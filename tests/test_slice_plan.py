"""CPU: the host plan of slicing across shards (sharding.slice_plan): every rank's send /
land descriptors, replayed with numpy for worlds of 1..13 ranks, reassemble exactly `global[region]` -- range,
stepped, reversed and Int-indexed leading axes, ranks that own nothing, every axis indexed (shape [1])."""
import itertools

import numpy as np
import pytest

import ph_core_b200  # loads the host library; no GPU needed
from ph_core_b200 import sharding as S, rng, ALL

G = np.arange(11 * 6 * 5, dtype=np.int64).reshape(11, 6, 5)
CASES = [((rng(2, 8),), np.s_[2:9]), ((rng(None, None, -1),), np.s_[::-1]), ((rng(1, None, 3), rng(0, None, 2)), np.s_[1::3, ::2]),
         ((4,), np.s_[4]), ((10, rng(1, 4), 2), np.s_[10, 1:5, 2]), ((rng(9, 2, -2), ALL, rng(None, None, -1)), np.s_[9:1:-2, :, ::-1]),
         ((3, 2, 1), np.s_[3, 2, 1]), ((rng(5, 5),), np.s_[5:6]), ((0, 0), np.s_[0, 0]), ((-1, ALL, rng(3, 0)), np.s_[-1, :, 3::-1])]


def _replay(g, key, npkey, world):
    want = np.asarray(g[npkey])
    want = want.reshape(want.shape if want.ndim else (1,))
    plans = [S.slice_plan(g.shape, list(key), world, r) for r in range(world)]
    assert all(p["new_shape"] == list(want.shape) for p in plans)
    assert not plans[0]["local"]
    outs = []
    for q in range(world):
        j0, j1 = S.shard_range(want.shape[0], world, q)
        outs.append(np.full([j1 - j0] + list(want.shape[1:]), -1, g.dtype))
    for r in range(world):
        r0, r1 = S.shard_range(g.shape[0], world, r)
        loc = np.ascontiguousarray(g[r0:r1]).reshape(-1)
        for q in range(world):
            sd, dd = plans[r]["send"][q], plans[r]["land"][q]
            assert (sd is None) == (dd is None)
            lo, hi = plans[q]["recv"][r]
            if not sd:
                assert hi <= lo
                continue
            (e, strd, off), (de, dstr, doff) = sd, dd
            assert list(e) == list(de)
            flat = outs[q].reshape(-1)
            for idx in itertools.product(*[range(x) for x in e]):
                flat[doff + sum(c * s for c, s in zip(idx, dstr))] = loc[off + sum(c * s for c, s in zip(idx, strd))]
            j0, _ = S.shard_range(want.shape[0], world, q)
            assert hi - lo == de[0] and (lo - j0) * dstr[0] == doff        # what q expects from r is what r sends
    got = np.concatenate(outs, axis=0)
    assert got.tobytes() == np.ascontiguousarray(want).tobytes()


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 13])
def test_every_ranks_blocks_reassemble_the_global_slice(world):
    for key, npkey in CASES:
        _replay(G, key, npkey, world)


def test_whole_leading_axis_is_local_and_bad_literals_raise_like_the_reference():
    p = S.slice_plan(G.shape, [ALL, rng(1, 3)], 4, 2)
    assert p["local"] and p["new_shape"] == [11, 3, 5]
    with pytest.raises(ph_core_b200.CrIndexError):
        S.slice_plan(G.shape, [rng(0, 11)], 4, 0)                          # ..11 on a bound of 11 (spec_helper.cr:133-139)
    with pytest.raises(Exception):
        S.slice_plan(G.shape, [0, 0, 0, 0], 4, 0)                          # more entries than axes -> DimensionError


def test_row_chunks_cover_the_rows_once_and_taper_at_the_end():
    """pipeline.row_chunks: equal chunks, the last one optionally cut into halves (the un-overlapped tail of a
    host -> device -> host pipeline is its last chunk); every schedule is a partition of [0, n) in order."""
    from ph_core_b200.pipeline import row_chunks
    for n in (0, 1, 7, 10, 1000, 8192):
        for chunks in (1, 2, 3, 4, 8, 16, 50):
            for taper in (0, 1, 3, 7, 12):
                b = row_chunks(n, chunks, taper)
                assert [r for lo, hi in b for r in range(lo, hi)] == list(range(n)), (n, chunks, taper)
                assert all(lo < hi for lo, hi in b)
    for n in (0, 1, 7, 10, 1000, 8192):
        for ramp in (1, 2, 5, 9):
            b = row_chunks(n, 3, 2, ramp)
            assert [r for lo, hi in b for r in range(lo, hi)] == list(range(n)) and all(lo < hi for lo, hi in b)
    assert [hi - lo for lo, hi in row_chunks(8192, 4, 7, 5)] == [64, 64, 128, 256, 512, 1024, 2048, 2048, 1024, 512, 256, 128, 64, 32, 16, 16]
    b = row_chunks(8192, 4, 7)
    assert [hi - lo for lo, hi in b] == [2048, 2048, 2048, 1024, 512, 256, 128, 64, 32, 16, 16]
    assert row_chunks(8192, 16, 0) == [(k * 512, (k + 1) * 512) for k in range(16)]
