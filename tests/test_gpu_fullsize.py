"""T3 parity at BASELINE.json's FULL sizes (GPU).  Where the oracle finishes in seconds the
result is compared element for element; otherwise through size-independent properties
(round trips, involutions, exact checksums, planted extrema, locality of the stencil)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, heat, rng, _lib
from oracle import ph_oracle as O, c_oracle as CO
from gpu_util import assert_bits


def philox(shape, stream, dtype=np.float32):
    g = np.random.Generator(np.random.Philox(key=20261017, counter=[0, 0, stream, 0]))
    return (g.random(shape, dtype=dtype) * 2 - 1).astype(dtype)


def test_config1_elementwise_8192_full():
    """BASELINE configs[1]: a*b+c, b = [1,8192] row vector, 8192x8192 f32 -- every element,
    against the oracle's C port (flat, all cores) and its reference-structured twin on a slab."""
    a, b, c = philox((8192, 8192), 1), philox((1, 8192), 2), philox((8192, 8192), 3)
    a[0, :6] = [0.0, -0.0, np.inf, -np.inf, np.nan, np.float32(1e-45)]      # specials ride along
    CO.use_all_cores()
    want = CO.flat_mul_rowvec_add_f32(a, b, c)
    da, db, dc = D.from_host(a), D.from_host(b), D.from_host(c)
    got = (da.broadcast_op("*", db) + dc).to_host()
    assert_bits(got, want, "two-kernel a*b+c")
    assert_bits(da.mul_add(db, dc).to_host(), want, "fused a*b+c")
    ref = CO.ref_mul_rowvec_add_f32(a[:256].copy(), b, c[:256].copy())       # reference structure, 1 core
    assert_bits(got[:256], ref, "reference-structured slab")
    assert ph.DeviceNArray.take_flags() == 0


def test_config2_views_16384_f64():
    """BASELINE configs[2] at 16384x16384 f64 (2 GiB): value = flat index (exact in f64) makes any
    misplacement visible; checked through involutions and closed-form checksums."""
    n = 16384
    lib = _lib.load()
    src = D([n, n], np.float64)
    # build value = flat index on the device: (row * n) + col via broadcast add of two vectors
    rows = D.from_host((np.arange(n, dtype=np.float64) * n).reshape(n, 1))
    cols = D.from_host(np.arange(n, dtype=np.float64).reshape(1, n))
    src = rows.broadcast_op("+", cols)
    total = float(n * n) * (n * n - 1) / 2
    assert src.get(n - 1, n - 1) == n * n - 1 and src.get(5, 7) == 5 * n + 7
    # (i) rows 0,2,4..: element [i, j] must be (2i)*n + j
    g = src[rng(0, None, 2), rng(None, -1)]
    assert g.shape == [n // 2, n] and g.get(3, 9) == 6 * n + 9 and g.get(n // 2 - 1, n - 1) == (n - 2) * n + n - 1
    # row i of g holds (2i)*n + j: its sum is 2i*n^2 + n(n-1)/2, an integer < 2^42, and every partial sum of
    # the row sums is a multiple of 2^13 below 2^55 -- all exact in f64 whatever the fold order
    half = n // 2
    assert float(g.sum(axis=1).sum()) == float(n * n * half * (half - 1) + half * (n * (n - 1) // 2))
    assert g.sum(axis=1).get(half - 1) == float(2 * (half - 1) * n * n + n * (n - 1) // 2)
    # (ii) columns 0,2,4..
    g2 = src[rng(None, None), rng(0, None, 2)]
    assert g2.shape == [n, n // 2] and g2.get(11, 5) == 11 * n + 10
    # (iii) full reversal is an involution and maps [i, j] -> [n-1-i, n-1-j]
    r = src[rng(None, None, -1), rng(None, None, -1)]
    assert r.get(0, 0) == n * n - 1 and r.get(2, 3) == (n - 3) * n + (n - 4)
    assert r[rng(None, None, -1), rng(None, None, -1)].equals(src)
    del r, g, g2
    # (iv) transposed copy: [i, j] -> j*n + i ; transposing twice is the identity
    t = src.permute()
    assert t.get(3, 9) == 9 * n + 3 and t.get(n - 1, 0) == n - 1
    assert t.permute().equals(src)
    # (v) transposed scatter through a mutable view == transposed copy
    z = D([n, n], np.float64)
    z.mutable_view().permute().set_chunk([], src)
    assert z.equals(t)
    # (vi) the literal narr[..2, ..-1]: rows 0..2 inclusive
    lit = src[rng(None, 2), rng(None, -1)]
    assert lit.shape == [3, n]
    assert_bits(lit.to_host(), (np.arange(3, dtype=np.float64) * n).reshape(3, 1) + np.arange(n, dtype=np.float64), "literal")
    # checksum of checksums: row sums of the transpose = column sums of the source (exact in f64)
    cs = t.sum(axis=1).to_host()
    np.testing.assert_array_equal(cs, np.arange(n, dtype=np.float64) * n + n * (n - 1) / 2 * n)


def test_config3_reductions_1e9_f32():
    """BASELINE configs[3]: [1000,1000,1000] f32.  Parity data = integers in {-8..8} (every
    partial sum exact => bit-exact in any order) + planted unique max + planted tie."""
    shape = (1000, 1000, 1000)
    tile = np.random.RandomState(3).randint(-8, 9, size=1_000_000).astype(np.float32)
    x = D(shape, np.float32)
    t = D.from_host(tile)
    lib = _lib.load()
    for k in range(1000):                                         # 1000 copies of the same plane
        ph.check(lib.ph_d2d(x.ptr + k * 4_000_000, t.ptr, 4_000_000))
    plane_sum = int(tile.astype(np.int64).sum())
    assert float(x.sum()) == float(np.float32(plane_sum * 1000))
    s0 = x.sum(axis=0).to_host()                                  # ordered fold: exact
    np.testing.assert_array_equal(s0.reshape(-1), tile * np.float32(1000))
    s2 = x.sum(axis=2).to_host()
    np.testing.assert_array_equal(s2[0], tile.reshape(1000, 1000).sum(axis=1))
    np.testing.assert_array_equal(s2[999], s2[0])
    # planted extrema
    x.set_element([123, 456, 789], 1000.0)                         # unique max
    x.set_element([700, 1, 2], -1000.0)                            # tie for the min ...
    x.set_element([20, 3, 4], -1000.0)                             # ... lower flat index must win
    assert x.argmax() == (np.float32(1000.0), [123, 456, 789])
    assert x.argmin() == (np.float32(-1000.0), [20, 3, 4])
    assert x.max() == 1000.0 and x.min() == -1000.0
    am0 = x.argmax(axis=0)
    assert am0.get(456, 789) == 123
    am2 = x.argmax(axis=2)
    assert am2.get(123, 456) == 789
    assert x.max(axis=1).get(123, 789) == 1000.0
    # U[0,1) timing data: tolerance vs an f64 truth
    u = np.random.RandomState(4).rand(1_000_000).astype(np.float32)
    tu = D.from_host(u)
    for k in range(1000):
        ph.check(lib.ph_d2d(x.ptr + k * 4_000_000, tu.ptr, 4_000_000))
    truth = float(u.astype(np.float64).sum()) * 1000
    assert abs(float(x.sum()) - truth) <= 1e-4 * truth


def test_config5_heat_large_locality():
    """3-D heat at 1024x1024x1024 f32 (the 2048^3 grid has its own test below): after k steps a cell depends only on cells within distance k, so a sub-cube
    compared with the oracle run on a cropped domain must match bit for bit."""
    n, k = 1024, 4
    rs = np.random.RandomState(11)
    tile = (rs.rand(64, 64, 1024) * 100).astype(np.float32)
    g = D([n, n, n], np.float32)
    th = D.from_host(tile)
    lib = _lib.load()
    # fill the grid with a non-periodic pattern: blocks of the tile scaled by position
    gv = g.view()
    for bz in range(0, n, 64):
        for by in range(0, n, 64):
            g.set_chunk([rng(bz, bz + 63), rng(by, by + 63)], th + float((bz // 64) * 3 + (by // 64)))
    # crops: a corner (touches the fixed boundary) and an interior block
    crops = [(0, 40, 0, 48, 0, 160), (500, 540, 300, 348, 700, 860)]
    before = [g[rng(z0, z1 - 1), rng(y0, y1 - 1), rng(x0, x1 - 1)].to_host() for z0, z1, y0, y1, x0, x1 in crops]
    fin = heat.simulate(g, 0.1, k)
    for (z0, z1, y0, y1, x0, x1), b in zip(crops, before):
        want = b.copy()
        for _ in range(k):
            want = O.heat_step_nd(want, np.float32(0.1))
        got = fin[rng(z0, z1 - 1), rng(y0, y1 - 1), rng(x0, x1 - 1)].to_host()
        # faces of the crop that are NOT grid boundaries received wrong (held) data: trim k cells there
        sl = tuple(slice(0 if lo == 0 else k, None if hi == n else -k) for lo, hi in ((z0, z1), (y0, y1), (x0, x1)))
        assert_bits(got[sl], want[sl], f"crop {(z0, y0, x0)}")


def test_config5_heat_2048_cubed_crops_vs_oracle():
    """BASELINE configs[4] at its FULL size on one GPU (2 x 34.4 GB): 3-D heat on 2048^3 f32 with a
    non-constant field (a constant one is a fixed point of the stencil and could not tell a wrong kernel from
    a right one).  After k steps a cell depends only on cells within distance k, so crops replayed by the
    oracle must match bit for bit: a corner touching three fixed faces, an interior block that straddles the
    128-plane march boundary at plane 129 and tile boundaries in y / x, and the far corner.  k = 5 exercises
    two two-step passes plus the odd single step."""
    n, k = 2048, 5
    lib = _lib.load()
    tile = (np.random.RandomState(21).rand(1 << 22) * 100).astype(np.float32)      # 16 MiB, not a divisor of a plane row count
    t = D.from_host(tile)
    g, other = D([n, n, n], np.float32), D([n, n, n], np.float32)
    total = n ** 3
    step = tile.size - 4099 * 4                                                       # shifted copies: no two planes alike
    pos, k0 = 0, 0
    while pos < total:
        m = min(step, total - pos)
        ph.check(lib.ph_d2d(g.ptr + pos * 4, t.ptr + ((k0 * 52) % 4099) * 16, m * 4))
        pos += m
        k0 += 1
    crops = [(0, 40, 0, 48, 0, 160), (110, 150, 1000, 1048, 1900, 2048), (2008, 2048, 2000, 2048, 0, 136)]
    before = [g[rng(z0, z1 - 1), rng(y0, y1 - 1), rng(x0, x1 - 1)].to_host() for z0, z1, y0, y1, x0, x1 in crops]
    fin = heat.simulate_into(g, other, 0.1, k)
    for (z0, z1, y0, y1, x0, x1), b in zip(crops, before):
        want = b.copy()
        for _ in range(k):
            want = O.heat_step_nd(want, np.float32(0.1))
        got = fin[rng(z0, z1 - 1), rng(y0, y1 - 1), rng(x0, x1 - 1)].to_host()
        sl = tuple(slice(0 if lo == 0 else k, None if hi == n else -k) for lo, hi in ((z0, z1), (y0, y1), (x0, x1)))
        assert not np.array_equal(got[sl], b[sl])                                     # the field did move
        assert_bits(got[sl], want[sl], f"2048^3 crop {(z0, y0, x0)}")


def test_two_steps_per_pass_equal_single_steps_at_scale():
    """The temporally blocked kernels (two time steps per pass over HBM) against the SAME library's
    single-step entry point applied step by step, at sizes the oracle cannot reach: 2-D 16384^2
    f32 (BASELINE configs[0] generalised) and a 3-D slab of 2048^2 planes.  Bit-identical."""
    lib = _lib.load()
    for shape, steps in (((16384, 16384), 4), ((96, 2048, 2048), 5)):
        total = int(np.prod(shape))
        tile = (np.random.RandomState(len(shape)).rand(1 << 22) * 100).astype(np.float32)
        g = D(list(shape), np.float32)
        t = D.from_host(tile)
        for pos in range(0, total, tile.size):
            ph.check(lib.ph_d2d(g.ptr + pos * 4, t.ptr, min(tile.size, total - pos) * 4))
        cur = g
        for _ in range(steps):
            cur = heat.update_temp(cur, 0.1)                      # one launch per step (ph_heat_step)
        fused = heat.simulate(g.clone(), 0.1, steps)              # ph_heat_run: two steps per pass (+ an odd one)
        assert fused.equals(cur), f"{shape}: fused run differs from {steps} single steps"
        del cur, fused, g
