#!/usr/bin/env python
"""Per-kernel bandwidth table for every BASELINE.json config.  One JSON row per
kernel: algorithmic bytes, best/median ms over `reps` launches (CUDA events on the launching
stream), GB/s, fraction of the measured copy peak and of the nominal 8 TB/s.
    python benchmarks/bench_kernels.py [--quick] [--only NAME] > gpurun_out/kernels.jsonl"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time as _time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# Rows here repeat ONE kernel on the SAME operands back to back; with the library's alternating
# traversal direction every repetition would find the far end of its inputs in L2 (+10 % on the
# elementwise rows).  That reuse is real for chained operators (bench.py measures it on a*b+c) but
# an artefact here, so the per-kernel table streams every launch the same way: HBM numbers.
os.environ.setdefault("PH_FLAT_NO_ALTERNATE", "1")
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, heat, rng, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--warm", type=int, default=3, help="untimed launches per row (0 under ncu)")
ap.add_argument("--inner", type=int, default=4, help="back-to-back launches per timing (1 under ncu)")
ap.add_argument("--big-heat", action="store_true", help="include the 2048^3 f32 stencil (69 GB)")
ap.add_argument("--heat-shape", default="", help="only run the 3-D stencil on this z,y,x grid")
args = ap.parse_args()

ph.init(0)
lib = ph.load()
if args.heat_shape:
    args.only = "custom-heat"
peak = 6546.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=None, warm=None, inner=None):
    """`inner` back-to-back launches between two CUDA events (a lone launch would add ~2.5 us of
    launch latency and ramp to kernels that run for 80 us), repeated `reps` times."""
    reps = min(reps or args.reps, args.reps)       # --reps 1 (ncu runs) caps the per-row defaults too
    warm = args.warm if warm is None else warm
    inner = args.inner if inner is None else inner
    for _ in range(warm):
        fn()
    ts = []
    ms = C.c_float()
    for _ in range(reps):
        ph.check(lib.ph_timer_start())
        for _ in range(inner):
            fn()
        ph.check(lib.ph_timer_stop(C.byref(ms)))
        ts.append(ms.value / inner)
    return min(ts), statistics.median(ts)


def row(name, nbytes, fn, unit_count=None, unit=None, reps=None, note="", inner=None):
    if args.only and args.only not in name:
        return
    best, med = timeit(fn, reps, inner=inner)
    r = {"kernel": name, "bytes": int(nbytes), "ms_best": round(best, 5), "ms_median": round(med, 5),
         "gbs": round(nbytes / (med * 1e-3) / 1e9, 1), "frac_measured": round(nbytes / (med * 1e-3) / 1e9 / peak, 4),
         "frac_nominal_8000": round(nbytes / (med * 1e-3) / 1e9 / 8000, 4)}
    if unit_count:
        r[unit] = round(unit_count / (med * 1e-3) / 1e9, 2)
    if note:
        r["note"] = note
    print(json.dumps(r), flush=True)


def rand(shape, dtype, seed):
    g = np.random.Generator(np.random.Philox(key=20261017, counter=[0, 0, seed, 0]))
    if np.dtype(dtype).kind == "f":
        return g.random(shape, dtype=dtype)
    return g.integers(-8, 9, size=shape).astype(dtype)


def dev_rand(shape, dtype, seed):
    """Fill large arrays on the device from a small random tile (keeps host RAM/time small)."""
    total = int(np.prod(shape))
    tile = rand(min(total, 1 << 22), dtype, seed)
    d = D(shape, dtype)
    flat = D([total], dtype, d._buf)
    t = D.from_host(tile)
    pos = 0
    while pos < total:
        n = min(tile.size, total - pos)
        ph.check(lib.ph_d2d(flat.ptr + pos * flat.dtype.itemsize, t.ptr, n * flat.dtype.itemsize))
        pos += n
    ph.check(lib.ph_sync())
    return d


if args.heat_shape:
    H3 = tuple(int(v) for v in args.heat_shape.split(","))
    g = D(H3, np.float32)
    ph.check(lib.ph_fill_region(4, g.ptr, C.byref(g.desc()), np.array(1.0, np.float32).ctypes.data))
    cells = int(np.prod(H3))
    row(f"custom-heat 3-D {H3} f32 x4 steps", 8 * cells * 4, lambda: heat.simulate(g, 0.1, 4), cells * 4, "gcell_per_s", reps=3)
    sys.exit(0)

# ------------------------------------------------------------------ config 1: elementwise 8192^2 f32
E = (8192, 8192) if not args.quick else (2048, 8192)
NE = E[0] * E[1]
a, c = dev_rand(E, np.float32, 1), dev_rand(E, np.float32, 3)
b = D.from_host(rand((1, E[1]), np.float32, 2))
out = D(E, np.float32)
F32 = ph.K["PH_F32"]
da, dc, do, db = a.desc(), c.desc(), out.desc(), b.bcast_desc(E)
row("ewise a+c same shape f32 (flat)", 3 * NE * 4,
    lambda: ph.check(lib.ph_ewise_binary(ph.K["PH_ADD"], F32, a.ptr, C.byref(da), c.ptr, C.byref(dc), out.ptr, C.byref(do))))
row("ewise a*b rowvec broadcast f32 (rows)", 2 * NE * 4 + E[1] * 4,
    lambda: ph.check(lib.ph_ewise_binary(ph.K["PH_MUL"], F32, a.ptr, C.byref(da), b.ptr, C.byref(db), out.ptr, C.byref(do))))
row("ewise fused (a*b)+c f32", 3 * NE * 4 + E[1] * 4,
    lambda: ph.check(lib.ph_ewise_mul_add(F32, a.ptr, C.byref(da), b.ptr, C.byref(db), c.ptr, C.byref(dc), out.ptr, C.byref(do))))
s2 = np.array(2.0, np.float32)
row("ewise a*2 scalar f32", 2 * NE * 4,
    lambda: ph.check(lib.ph_ewise_scalar(ph.K["PH_MUL"], F32, a.ptr, C.byref(da), s2.ctypes.data, 0, out.ptr, C.byref(do))))
mask = D(E, np.bool_)
dm = mask.desc()
row("compare a>c -> bool f32", 2 * NE * 4 + NE,
    lambda: ph.check(lib.ph_compare(ph.K["PH_GT"], F32, a.ptr, C.byref(da), c.ptr, C.byref(dc), mask.ptr, C.byref(dm))))
row("mask store scalar f32", NE + 2 * NE * 4,
    lambda: ph.check(lib.ph_mask_set_scalar(4, out.ptr, C.byref(do), mask.ptr, C.byref(dm), s2.ctypes.data)),
    note="bytes = mask + read-modify-write of dst (upper bound; untouched groups skip dst)")
row("mask store array f32", NE + 3 * NE * 4,
    lambda: ph.check(lib.ph_mask_set_array(4, out.ptr, C.byref(do), mask.ptr, C.byref(dm), a.ptr, C.byref(da))))
ai = dev_rand(E, np.int32, 4); ci = dev_rand(E, np.int32, 5); oi = D(E, np.int32)
row("ewise a+c i32 overflow-checked", 3 * NE * 4,
    lambda: ph.check(lib.ph_ewise_binary(ph.K["PH_ADD"], ph.K["PH_I32"], ai.ptr, C.byref(da), ci.ptr, C.byref(dc), oi.ptr, C.byref(do))))
# ---- SURVEY.md 8(f) "next" rows on the same arrays: tile (f-2) and fill
row("fill 8192^2 f32 (NArray.fill)", NE * 4,
    lambda: ph.check(lib.ph_fill_region(4, out.ptr, C.byref(do), s2.ctypes.data)), note="write-only stream")
def tile_descs(srcarr, counts):
    """The rank-2N stride-0 source descriptor of MultiIndexable#tile and the contiguous destination."""
    d = srcarr.desc()
    ext, strd = [], []
    for i, cnt in enumerate(counts):
        ext += [int(cnt), int(d.extent[i])]
        strd += [0, int(d.stride[i])]
    return ph.PhDesc.make(ext, strd, d.offset), ph.PhDesc.contiguous(ext)
ts, td = tile_descs(b, [E[0], 1])
row("tile [1,8192] x [8192,1] f32 (broadcast oracle)", NE * 4 + E[1] * 4,
    lambda: ph.check(lib.ph_copy_strided(4, b.ptr, C.byref(ts), out.ptr, C.byref(td))),
    note="f-2: rank-4 stride-0 descriptor, no modulo on the device")
row("tile [1,8192] x [8192,1] f32 through the Python mirror", NE * 4 + E[1] * 4, lambda: b.tile([E[0], 1]),
    note="same kernel + result allocation + ~10 ctypes calls per launch: host-bound")
quarter = a[rng(0, E[0] // 2, exclusive=True), rng(0, E[1] // 2, exclusive=True)]
qs, qd = tile_descs(quarter, [2, 2])
row("tile [4096,4096] x [2,2] f32", NE * 4 + NE, lambda: ph.check(lib.ph_copy_strided(4, quarter.ptr, C.byref(qs), out.ptr, C.byref(qd))),
    note="f-2: every source element read 4 times (3 of them from L2)")
del quarter
del a, c, out, mask, ai, ci, oi

# ------------------------------------------------------------------ config 2: strided views 16384^2 f64
S = 16384 if not args.quick else 4096
src = dev_rand((S, S), np.float64, 6)
NS = S * S
row("gather narr[0..2.., ..-1] f64 (row-strided)", 2 * (NS // 2) * 8, lambda: src[rng(0, None, 2), rng(None, -1)])
row("gather narr[.., 0..2..] f64 (col-strided)", 2 * (NS // 2) * 8, lambda: src[rng(None, None), rng(0, None, 2)],
    note="physical sectors = 1.5x algorithmic; ceiling 66.7% on algorithmic bytes")
row("gather reversed narr[..-1.., ..-1..] f64", 2 * NS * 8, lambda: src[rng(None, None, -1), rng(None, None, -1)])
row("transposed copy narr.permute f64", 2 * NS * 8, lambda: src.permute())
dst = D([S, S], np.float64)
row("transposed scatter mutable_view.permute[..]=src f64", 2 * NS * 8,
    lambda: dst.mutable_view().permute().set_chunk([], src))
row("clone (contiguous copy) f64", 2 * NS * 8, lambda: src.clone())
del dst
# ---- reductions over VIEWS of the same array, read in place (no gather into a temporary; PH_REDUCE_GATHER=1
#      restores the gather-first form for A/B): bytes = the bytes of the view (+ the per-axis output)
v_rows = src.view(rng(0, None, 2), ph.ALL)
row("view reduce narr[0..2.., ..].sum f64 (row-strided, in place)", (NS // 2) * 8, lambda: v_rows.sum(), reps=10)
row("view reduce narr[0..2.., ..].max f64 (row-strided, in place)", (NS // 2) * 8, lambda: v_rows.max(), reps=10)
v_rev = src.view().reverse()
row("view reduce view.reverse.max(axis: 1) f64 (reversed rows, in place)", NS * 8 + S * 8, lambda: v_rev.max(axis=1), reps=10)
row("view reduce view.reverse.sum(axis: 0) f64 (reversed rows, in place)", NS * 8 + S * 8, lambda: v_rev.sum(axis=0), reps=10)
row("view reduce view.reverse.argmax(axis: 1) f64 (reversed rows, in place)", NS * 8 + S * 8, lambda: v_rev.argmax(axis=1), reps=10)
del v_rows, v_rev
# ---- a 2-D matrix folded down its columns has FEW columns (16384): the cp.async-staged strip kernel
#      (PH_AXIS_STAGED=0 restores one thread per column); long rows: the one-pass row kernel
for name in ["sum", "max", "argmax"]:
    row(f"reduce axis=0 {name} f64 [{S},{S}] (few columns: staged strips)", NS * 8 + S * 8, lambda name=name: getattr(src, name)(axis=0), reps=10)
src32 = dev_rand((S, S), np.float32, 11)
for name in ["sum", "max"]:
    row(f"reduce axis=0 {name} f32 [{S},{S}] (few columns: staged strips)", NS * 4 + S * 4, lambda name=name: getattr(src32, name)(axis=0), reps=10)
del src32
for name in ["sum", "max", "argmax"]:
    row(f"reduce axis=1 {name} f64 [{S},{S}] (128 KiB rows)", NS * 8 + S * 8, lambda name=name: getattr(src, name)(axis=1), reps=10)
# ---- f-2: each_slice on a rank-3 view of the same buffer (one gather launch per slice)
cube = D([64, S // 64 * S // 1024, 1024], np.float64, src._buf) if S % 64 == 0 else None
if cube is not None:
    ncube = cube.size
    sl_out = D([cube.shape[1], 1024], np.float64)
    cd = cube.desc()
    def slice_descs(axis, idx):
        ext = [int(cd.extent[i]) for i in range(3) if i != axis]
        strd = [int(cd.stride[i]) for i in range(3) if i != axis]
        return ph.PhDesc.make(ext, strd, idx * int(cd.stride[axis])), ph.PhDesc.contiguous(ext)
    s0 = [slice_descs(0, i) for i in range(64)]
    def run_slices(descs):
        for sd, dd in descs:
            ph.check(lib.ph_copy_strided(8, cube.ptr, C.byref(sd), sl_out.ptr, C.byref(dd)))
    row("each_slice(axis=0) of [64,%d,1024] f64 (64 contiguous gathers)" % cube.shape[1], 2 * ncube * 8,
        lambda: run_slices(s0), reps=5, inner=1, note="f-2: 64 launches of 32 MiB each, descriptors prebuilt")
    step1 = max(1, cube.shape[1] // 64)
    s1 = [slice_descs(1, j) for j in range(0, cube.shape[1], step1)][:64]
    row("each_slice(axis=1)[::%d] of the same f64 (strided gathers)" % step1, 2 * 64 * 64 * 1024 * 8,
        lambda: run_slices(s1), reps=5, inner=1, note="f-2: 64 slices of [64,1024] (1 KiB rows, 32 MiB apart), 512 KiB per launch: launch-bound")
    # what each_slice IS now: views over the source (no copy, no launch); the reference's per-axis idiom -- a fold over
    # each_slice -- then costs only the consumer's kernels, which read the strided slices in place
    t0 = _time.perf_counter(); views = list(cube.each_slice(1)); t1 = _time.perf_counter()
    if not args.only or "each_slice" in args.only:
        print(json.dumps({"kernel": "each_slice(axis=1) of [64,%d,1024] f64: %d views" % (cube.shape[1], len(views)), "bytes": 0,
                          "kernel_launches": 0, "host_ms": round((t1 - t0) * 1e3, 2),
                          "note": "f-2: descriptors over the source buffer -- 0 bytes moved, 0 launches (round 1: one gather per index, the row above)"}), flush=True)
    nfold = 64
    def fold_views():
        acc = views[0] + views[1]
        for v in views[2:nfold]:
            acc = acc + v
        return acc
    row("fold of the first %d each_slice(axis=1) views with `+` (the reference's per-axis idiom)" % nfold,
        (nfold - 1) * 3 * 64 * 1024 * 8, fold_views, reps=5, inner=1,
        note="f-2: %d launches of 1.5 MiB each, strided views read in place: launch-bound by construction; sum(axis: 1) is the one-launch form" % (nfold - 1))
    del views
    perm_src = cube.view().permute(1, 0, 2)
    perm_out = D(perm_src.shape, np.float64)
    psd, pod = perm_src.desc(), perm_out.desc()
    row("slices(axis=1), all %d at once: ONE permuting copy [64,%d,1024] -> [%d,64,1024] f64" % (cube.shape[1], cube.shape[1], cube.shape[1]),
        2 * ncube * 8, lambda: ph.check(lib.ph_copy_strided(8, cube.ptr, C.byref(psd), perm_out.ptr, C.byref(pod))), reps=5,
        note="f-2: what DeviceNArray#slices does now; the per-index gathers above are the reference's structure")
    del perm_out
    row("slices(axis=0) through the Python mirror", 2 * ncube * 8, lambda: cube.slices(0), reps=5, inner=1,
        note="one copy + 64 Python objects")
    del sl_out
# ---- f-4: binary dump / load of a device array (host file system either side of the path)
from ph_core_b200 import io as phio
import tempfile
if not args.only or "dump" in args.only:
    part = D([(1 << 23) if args.quick else (1 << 26)], np.float64, src._buf)      # 64 MiB quick / 512 MiB
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "a.phbin")
        t0 = _time.perf_counter(); phio.dump(part, path); t1 = _time.perf_counter()
        back = phio.load(path); ph.check(lib.ph_sync()); t2 = _time.perf_counter()
        ok = bool(back.equals(part))
    nb = part.size * 8
    print(json.dumps({"kernel": "io dump / load binary f64 (%d MiB, tmpfs or local disk)" % (nb >> 20), "bytes": nb,
                      "dump_gbs": round(nb / (t1 - t0) / 1e9, 2), "load_gbs": round(nb / (t2 - t1) / 1e9, 2),
                      "round_trip_bit_exact": ok, "note": "f-4: wall clock incl. D2H / H2D and file I/O; host-bound"}), flush=True)
    del part, back
del src

# ------------------------------------------------------------------ config 3: reductions 1e9 f32
R = (1000, 1000, 1000) if not args.quick else (250, 1000, 1000)
NR = int(np.prod(R))
x = dev_rand(R, np.float32, 7)
for name in ["sum", "max", "argmax"]:
    row(f"reduce full {name} f32 {R}", NR * 4, lambda name=name: getattr(x, name)(), reps=10)
for axis in range(3):
    for name in ["sum", "max", "argmax"]:
        osz = NR // R[axis] * (8 if name == "argmax" else 4)
        row(f"reduce axis={axis} {name} f32 {R}", NR * 4 + osz, lambda name=name, axis=axis: getattr(x, name)(axis=axis), reps=10)
del x

# ------------------------------------------------------------------ configs 4/5: heat
H2 = (16384, 16384) if not args.quick else (4096, 4096)
g = dev_rand(H2, np.float32, 8)
cells = int(np.prod(H2))
steps = 10
row(f"heat 2-D {H2} f32 x{steps} steps", 8 * cells * steps, lambda: heat.simulate(g, 0.1, steps), cells * steps, "gcell_per_s", reps=5)
del g
H3 = (1024, 1024, 1024) if not args.quick else (256, 512, 512)
g = dev_rand(H3, np.float32, 9)
cells = int(np.prod(H3))
row(f"heat 3-D {H3} f32 x{steps} steps", 8 * cells * steps, lambda: heat.simulate(g, 0.1, steps), cells * steps, "gcell_per_s", reps=5)
del g
if args.big_heat:
    H3 = (2048, 2048, 2048)
    g = D(H3, np.float32)
    ph.check(lib.ph_fill_region(4, g.ptr, C.byref(g.desc()), np.array(1.0, np.float32).ctypes.data))
    cells = int(np.prod(H3))
    row(f"heat 3-D {H3} f32 x4 steps", 8 * cells * 4, lambda: heat.simulate(g, 0.1, 4), cells * 4, "gcell_per_s", reps=3)
