"""T2 parity (GPU): the heat stencil vs the oracle.
Reference: examples/heat_equation.cr:5-51; N-D rule: SURVEY.md 8(a) a-9.
One step is bit-exact (same operator order, no FMA); multi-step runs are held to the
north_star tolerances (1e-6 f64, 1e-4 f32) -- and in fact stay bit-exact too."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, heat
from oracle import ph_oracle as O
from gpu_util import assert_bits


def test_example_1d_replay():
    """The example verbatim: N = 21 f64, 10 001 steps (examples/heat_equation.cr:22-36)."""
    state = np.full(21, 20.0); state[0], state[-1] = 0.0, 100.0
    c = O.heat_example_coeff()
    one = heat.update_temp(D.from_host(state), c, heat.EXAMPLE1D).to_host()
    assert_bits(one, O.heat_step_1d_example(state, c), "one step")
    got = heat.simulate(D.from_host(state), c, 10001, heat.EXAMPLE1D).to_host()
    want = O.heat_simulate_1d_example()
    np.testing.assert_allclose(got, want, rtol=1e-6)
    assert_bits(got, want, "10001 steps")
    assert abs(got.sum() - 480.0) < 1e-9


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [3, 21, 1024, 1025, 5000])
def test_1d_both_modes(dtype, n):
    rs = np.random.RandomState(n)
    s = (rs.rand(n) * 100).astype(dtype)
    c = dtype(0.25)
    assert_bits(heat.update_temp(D.from_host(s), c, heat.EXAMPLE1D).to_host(), O.heat_step_1d_example(s, c), "example")
    assert_bits(heat.update_temp(D.from_host(s), c, heat.FIXED).to_host(), O.heat_step_nd(s, c), "fixed")
    want = s.copy()
    for _ in range(7):
        want = O.heat_step_nd(want, c)
    assert_bits(heat.simulate(D.from_host(s), c, 7, heat.FIXED).to_host(), want, "7 fixed steps")
    want = s.copy()
    for _ in range(6):
        want = O.heat_step_1d_example(want, c)
    assert_bits(heat.simulate(D.from_host(s), c, 6, heat.EXAMPLE1D).to_host(), want, "6 example steps")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(3, 3), (5, 4), (64, 128), (37, 1000), (300, 33), (130, 1028), (4, 4100)])
def test_2d_one_step_and_run(dtype, shape):
    rs = np.random.RandomState(shape[0])
    s = (rs.rand(*shape) * 100).astype(dtype)
    c = dtype(0.1)
    assert_bits(heat.update_temp(D.from_host(s), c).to_host(), O.heat_step_nd(s, c), f"2d step {shape}")
    want = s.copy()
    for _ in range(5):
        want = O.heat_step_nd(want, c)
    assert_bits(heat.simulate(D.from_host(s), c, 5).to_host(), want, f"2d 5 steps {shape}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_2d_two_step_pass_tiles_and_chunks(dtype):
    """The 2-D two-steps-per-pass kernel (warp-shuffle temporal blocking): shapes with one partial
    tile, several tiles (30 groups of output each), a last tile holding only the final columns,
    row counts that split into several chunks, and odd / even step counts."""
    rs = np.random.RandomState(17)
    for shape, steps in [((9, 8), 4), ((40, 240), 6), ((33, 248), 7), ((300, 488), 6), ((700, 1024), 8),
                         ((131, 2056), 5), ((520, 16), 10)]:
        s = (rs.rand(*shape) * 100).astype(dtype)
        c = dtype(0.1)
        want = s.copy()
        for _ in range(steps):
            want = O.heat_step_nd(want, c)
        assert_bits(heat.simulate(D.from_host(s), c, steps).to_host(), want, f"2d {steps} steps {shape}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_2d_two_step_slabs_match_whole_grid(dtype):
    """Rank-2 slabs with two ghost rows per side advanced two steps per pass (ph_heat_pass_slab)."""
    import ctypes as C
    from ph_core_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(23)
    shape = (44, 520)
    s = (rs.rand(*shape) * 100).astype(dtype)
    c = np.array(0.1, dtype)
    steps = 6
    want = s.copy()
    for _ in range(steps):
        want = O.heat_step_nd(want, dtype(0.1))
    g, half = 2, shape[0] // 2
    slabs = []
    for r in range(2):
        loc = np.zeros((half + 2 * g, shape[1]), dtype)
        loc[g:-g] = s[r * half:(r + 1) * half]
        if r == 0:
            loc[-g:] = s[half:half + g]
        else:
            loc[:g] = s[half - g:half]
        slabs.append([D.from_host(loc), D.from_host(loc)])
    ext = (C.c_int64 * 2)(half + 2 * g, shape[1])
    row = shape[1] * np.dtype(dtype).itemsize
    code = ph.K["PH_F32"] if dtype == np.float32 else ph.K["PH_F64"]
    for t in range(steps // 2):
        for r in range(2):
            src, dst = slabs[r][t & 1], slabs[r][(t & 1) ^ 1]
            for b, e in ((g, 2 * g), (half, half + g), (2 * g, half)):
                ph.check(lib.ph_heat_pass_slab(code, 2, ext, c.ctypes.data, g, 1, int(r > 0), int(r < 1), b, e,
                                               src.ptr, dst.ptr, None))
        a, b = slabs[0][(t & 1) ^ 1], slabs[1][(t & 1) ^ 1]
        ph.check(lib.ph_d2d(a.ptr + (half + g) * row, b.ptr + g * row, g * row))
        ph.check(lib.ph_d2d(b.ptr, a.ptr + half * row, g * row))
    fin = (steps // 2) & 1
    got = np.concatenate([slabs[0][fin].to_host()[g:-g], slabs[1][fin].to_host()[g:-g]])
    assert_bits(got, want, "two 2-ghost 2-D slabs, two steps per pass == whole grid")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(3, 3, 3), (5, 6, 4), (9, 16, 128), (40, 37, 132), (7, 70, 31), (66, 9, 260), (34, 8, 1024)])
def test_3d_one_step_and_run(dtype, shape):
    rs = np.random.RandomState(shape[1])
    s = (rs.rand(*shape) * 100).astype(dtype)
    c = dtype(0.1)
    assert_bits(heat.update_temp(D.from_host(s), c).to_host(), O.heat_step_nd(s, c), f"3d step {shape}")
    want = s.copy()
    for _ in range(4):
        want = O.heat_step_nd(want, c)
    assert_bits(heat.simulate(D.from_host(s), c, 4).to_host(), want, f"3d 4 steps {shape}")
    for _ in range(3):                                             # 7 = 3 two-step passes + 1 single step
        want = O.heat_step_nd(want, c)
    assert_bits(heat.simulate(D.from_host(s), c, 7).to_host(), want, f"3d 7 steps {shape}")


def test_3d_100_steps_tolerance_and_conservation():
    """256-class grid, 100 steps, f32: north_star tolerance 1e-4 vs the oracle."""
    rs = np.random.RandomState(20261017 % 2**31)
    shape = (48, 64, 128)
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    s = (np.exp(-4 * (x * x + y * y + z * z)) * 100 + rs.rand(*shape)).astype(np.float32)
    want = s.copy()
    for _ in range(100):
        want = O.heat_step_nd(want, np.float32(0.1))
    got = heat.simulate(D.from_host(s), np.float32(0.1), 100).to_host()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    assert_bits(got, want, "100 steps stay bit-exact")
    assert np.array_equal(got[0], s[0]) and np.array_equal(got[:, :, -1], s[:, :, -1])   # boundary held


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_3d_two_step_pass_tiles_and_chunks(dtype):
    """Shapes that span several TMA tiles in x / y and several z-chunks, so tile halos, the
    recomputed t+1 planes at chunk seams and partial tiles of the two-step kernel are all hit."""
    rs = np.random.RandomState(5)
    for shape in [(150, 40, 272), (70, 35, 516), (200, 19, 128)]:
        s = (rs.rand(*shape) * 100).astype(dtype)
        c = dtype(0.1)
        want = s.copy()
        for _ in range(6):
            want = O.heat_step_nd(want, c)
        assert_bits(heat.simulate(D.from_host(s), c, 6).to_host(), want, f"6 steps {shape}")


def test_heat_slab_matches_whole_grid():
    """Slab decomposition emulated on one GPU: two slabs with ghost planes,
    halos exchanged by copies, must be bit-identical to the undivided grid."""
    import ctypes as C
    from ph_core_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(1)
    shape = (20, 24, 64)
    s = (rs.rand(*shape) * 100).astype(np.float32)
    c = np.array(0.1, np.float32)
    want = s.copy()
    steps = 6
    for _ in range(steps):
        want = O.heat_step_nd(want, np.float32(0.1))
    half = shape[0] // 2
    slabs = []
    for r in range(2):
        loc = np.zeros((half + 2,) + shape[1:], np.float32)
        loc[1:-1] = s[r * half:(r + 1) * half]
        if r == 0:
            loc[-1] = s[half]
        else:
            loc[0] = s[half - 1]
        slabs.append([D.from_host(loc), D.from_host(loc)])
    ext = (C.c_int64 * 3)(half + 2, shape[1], shape[2])
    plane = shape[1] * shape[2] * 4
    for t in range(steps):
        for r in range(2):
            src, dst = slabs[r][t & 1], slabs[r][(t & 1) ^ 1]
            ph.check(lib.ph_heat_step_slab(ph.K["PH_F32"], 3, ext, c.ctypes.data, int(r > 0), int(r < 1), 1, half + 1,
                                           src.ptr, dst.ptr, None))
        a, b = slabs[0][(t & 1) ^ 1], slabs[1][(t & 1) ^ 1]
        ph.check(lib.ph_d2d(a.ptr + (half + 1) * plane, b.ptr + 1 * plane, plane))       # rank0 hi ghost <- rank1 first
        ph.check(lib.ph_d2d(b.ptr, a.ptr + half * plane, plane))                          # rank1 lo ghost <- rank0 last
    got = np.concatenate([slabs[0][steps & 1].to_host()[1:-1], slabs[1][steps & 1].to_host()[1:-1]])
    assert_bits(got, want, "two slabs == whole grid")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_two_step_slabs_match_whole_grid(dtype):
    """Two slabs with TWO ghost planes per side, advanced two time steps per pass
    (ph_heat_pass_slab, two_steps = 1) with the 2-plane halos exchanged by copies every pass:
    bit-identical to the undivided grid stepped by the oracle."""
    import ctypes as C
    from ph_core_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(11)
    shape = (22, 37, 144)
    s = (rs.rand(*shape) * 100).astype(dtype)
    c = np.array(0.1, dtype)
    steps = 6
    want = s.copy()
    for _ in range(steps):
        want = O.heat_step_nd(want, dtype(0.1))
    g, half = 2, shape[0] // 2
    slabs = []
    for r in range(2):
        loc = np.zeros((half + 2 * g,) + shape[1:], dtype)
        loc[g:-g] = s[r * half:(r + 1) * half]
        if r == 0:
            loc[-g:] = s[half:half + g]
        else:
            loc[:g] = s[half - g:half]
        slabs.append([D.from_host(loc), D.from_host(loc)])
    ext = (C.c_int64 * 3)(half + 2 * g, shape[1], shape[2])
    plane = shape[1] * shape[2] * np.dtype(dtype).itemsize
    code = ph.K["PH_F32"] if dtype == np.float32 else ph.K["PH_F64"]
    for t in range(steps // 2):
        for r in range(2):
            src, dst = slabs[r][t & 1], slabs[r][(t & 1) ^ 1]
            # edges first, then the interior: the split a sharded run uses
            for b, e in ((g, 2 * g), (half, half + g), (2 * g, half)):
                ph.check(lib.ph_heat_pass_slab(code, 3, ext, c.ctypes.data, g, 1, int(r > 0), int(r < 1), b, e,
                                               src.ptr, dst.ptr, None))
        a, b = slabs[0][(t & 1) ^ 1], slabs[1][(t & 1) ^ 1]
        ph.check(lib.ph_d2d(a.ptr + (half + g) * plane, b.ptr + g * plane, g * plane))    # rank0 hi ghosts <- rank1 first two
        ph.check(lib.ph_d2d(b.ptr, a.ptr + half * plane, g * plane))                      # rank1 lo ghosts <- rank0 last two
    fin = (steps // 2) & 1
    got = np.concatenate([slabs[0][fin].to_host()[g:-g], slabs[1][fin].to_host()[g:-g]])
    assert_bits(got, want, "two 2-ghost slabs, two steps per pass == whole grid")


def test_sharded_run_single_rank_two_ghost_layout():
    """ph_heat_run_sharded on a 1-rank communicator with the 2-ghost layout (no neighbours: the
    ghost planes are unused and planes 2 / n-3 are the fixed boundary) == the plain run."""
    from ph_core_b200 import sharding as S
    S.comm_init()
    rs = np.random.RandomState(12)
    field = (rs.rand(19, 21, 40) * 100).astype(np.float32)
    for steps in (5, 8):
        want = field.copy()
        for _ in range(steps):
            want = O.heat_step_nd(want, np.float32(0.1))
        loc = S.slab_from_global(field, 1, 0, ghost=2)
        a, b = D.from_host(loc), D.from_host(loc)
        got = S.heat_run_sharded(a, b, 0.1, steps, ghost=2).to_host()[2:-2]
        assert_bits(got, want, f"1-rank sharded run, 2 ghost planes, {steps} steps")


def test_example_update_temp_written_with_the_reference_operators():
    """examples/heat_equation.cr:38-51 replayed LITERALLY on device arrays: one-element chunks
    (`state[1] - state[0]`), array-valued `[]=`, `2 * center_temp` with the scalar on the left,
    `to_scalar` on fully indexed chunks.  Every operator is its own launch; the result must be
    bit-identical to the oracle's restatement and to the fused PH_HEAT_EXAMPLE1D kernel."""
    from ph_core_b200.region import rng
    coeff = O.heat_example_coeff()
    host = np.full(21, 20.0)
    host[0], host[-1] = 0.0, 100.0
    state = D.from_host(host)
    want = host.copy()
    for _ in range(3):
        temp_diff = D.fill(state.shape, 0.0, np.float64)                       # :39
        temp_diff[0] = (state[1] - state[0]) * coeff                           # :43
        temp_diff[-1] = (state[-2] - state[-1]) * coeff                        # :44
        centre = state[rng(1, -1, exclusive=True)]                              # :46
        assert centre.shape == [19]
        for idx in range(centre.shape[0]):
            center_temp = centre[idx].to_scalar()                               # each_with_index yields elements
            temp_diff[idx + 1] = (state[idx] - 2 * center_temp + state[idx + 2]) * coeff   # :47
        state = state + temp_diff                                               # :50 (map_with_coord { el + temp_diff.get(idx) })
        want = O.heat_step_1d_example(want, coeff)
        assert_bits(state.to_host(), want, "literal update_temp")
    assert_bits(heat.simulate(D.from_host(host), coeff, 3, heat.EXAMPLE1D).to_host(), want, "fused example kernel")
    assert state[0].scalar() and not state.scalar() and not state.empty()
    assert state.first() == want[0] and state.last() == want[-1] and state[3].to_f() == float(want[3])
    with pytest.raises(ph.ShapeError):
        state.to_scalar()
    assert state.to_scalar_or_none() is None and state.sample() in want
