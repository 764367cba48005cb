"""Run under torchrun (one rank per GPU): N-GPU results vs the 1-GPU / oracle results.
  * heat: slab-decomposed run with NCCL halo exchange must be BIT-IDENTICAL to the undivided
    grid (slabbing changes no cell's arithmetic) -- checked against the oracle on a small grid
    and against a single-GPU run of the same library on a larger one;
  * reductions: allreduce of per-GPU partials (sum on exactly-summable data: exact; max: exact;
    argmax: first extremum across shards).
Prints MGPU_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, sharding as S, heat
from oracle import ph_oracle as O


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ph.init(local)
    world, rank = S.comm_init(dist)
    run_checks(world, rank)
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} p2p={S.p2p_ready()}")
    dist.destroy_process_group()


_SOFT = False
_FAILED = []
_CHECKS = 0


def _chk(cond, msg) -> None:
    """A check.  Hard mode (the test): AssertionError at once.  Soft mode (bench.py): the disagreement is recorded
    and the rank carries on, so that every rank still walks the same sequence of collectives -- a rank that
    stopped at its own failed comparison would leave the others waiting in the next all-gather."""
    global _CHECKS
    _CHECKS += 1
    if cond:
        return
    if not _SOFT:
        raise AssertionError(msg)
    _FAILED.append(str(msg)[:200])


def run_checks(world: int, rank: int, soft: bool = False) -> dict:
    """Every N-GPU agreement check, on an initialised communicator (S.comm_init done).  Hard mode raises
    AssertionError on the first disagreement; soft mode collects them.  Returns what was covered.  bench.py runs
    this (soft) at N > 1 outside its timed regions, so the driver's own multi-GPU runs carry the parity evidence
    (`extras.multi_gpu_parity`)."""
    global _SOFT, _CHECKS
    _SOFT = soft
    _CHECKS = 0
    del _FAILED[:]
    rs = np.random.RandomState(7)

    # ---- heat, small grid vs oracle (uneven split); one ghost plane (a step per exchange) and two
    #      (two time steps per pass and per exchange), odd and even step counts
    field = (rs.rand(4 * world + 3, 20, 36) * 100).astype(np.float32)
    for ghost, steps in ((1, 7), (2, 7), (2, 6)):
        want = field.copy()
        for _ in range(steps):
            want = O.heat_step_nd(want, np.float32(0.1))
        lay = S.slab_layout(field.shape[0], world, rank, ghost)
        loc = S.slab_from_global(field, world, rank, ghost)
        a, b = D.from_host(loc), D.from_host(loc)
        fin = S.heat_run_sharded(a, b, 0.1, steps, ghost).to_host()[ghost:-ghost]
        _chk(fin.tobytes() == want[lay["start"]:lay["stop"]].tobytes(), f"rank {rank}: small heat slab (ghost {ghost}, {steps} steps) differs from oracle")

    # ---- heat, larger grid vs a single-GPU run of the same library (rank 0 computes it)
    big = (rs.rand(16 * world, 96, 256) * 100).astype(np.float32)
    for ghost, steps in ((1, 9), (2, 9), (2, 12)):
        lay = S.slab_layout(big.shape[0], world, rank, ghost)
        loc = S.slab_from_global(big, world, rank, ghost)
        a, b = D.from_host(loc), D.from_host(loc)
        fin = S.heat_run_sharded(a, b, 0.1, steps, ghost).to_host()[ghost:-ghost]
        whole = heat.simulate(D.from_host(big), 0.1, steps).to_host()
        _chk(fin.tobytes() == whole[lay["start"]:lay["stop"]].tobytes(), f"rank {rank}: big heat slab (ghost {ghost}, {steps} steps) differs from 1-GPU run")

    # ---- 2-D grid (BASELINE configs[0] generalised), two ghost rows, two steps per pass
    flat2 = (rs.rand(64 * world + 5, 520) * 100).astype(np.float32)
    for ghost, steps in ((2, 8), (2, 5), (1, 3)):
        want = flat2.copy()
        for _ in range(steps):
            want = O.heat_step_nd(want, np.float32(0.1))
        lay = S.slab_layout(flat2.shape[0], world, rank, ghost)
        loc = S.slab_from_global(flat2, world, rank, ghost)
        a, b = D.from_host(loc), D.from_host(loc)
        fin = S.heat_run_sharded(a, b, 0.1, steps, ghost).to_host()[ghost:-ghost]
        _chk(fin.tobytes() == want[lay["start"]:lay["stop"]].tobytes(), f"rank {rank}: 2-D heat slab (ghost {ghost}, {steps} steps) differs from oracle")

    # ---- the same three heat cases with the slabs in peer-mapped memory: the stencil kernel stores the halo
    #      planes straight into the neighbours' ghosts (no NCCL); without P2P these run the NCCL form again
    p2p = S.p2p_ready()
    for name, grid in (("small", field), ("big", big), ("2-D", flat2)):
        for ghost, steps in ((1, 5), (2, 7), (2, 12), (2, 2)):
            want = grid.copy()
            for _ in range(steps):
                want = O.heat_step_nd(want, np.float32(0.1))
            lay = S.slab_layout(grid.shape[0], world, rank, ghost)
            loc = S.slab_from_global(grid, world, rank, ghost)
            a, b = S.symm_from_host(loc), S.symm_from_host(loc)
            for rep in range(2):                                           # a second run re-uses the mapped slabs
                fin = S.heat_run_sharded(a, b, 0.1, steps, ghost)
                got = fin.to_host()
                _chk(got[ghost:-ghost].tobytes() == want[lay["start"]:lay["stop"]].tobytes(), f"rank {rank}: {name} heat slab in peer-mapped memory (ghost {ghost}, {steps} steps, p2p={p2p}) differs")
                # the ghost planes of the final state are the neighbours' edge planes
                if lay["lo_rank"] >= 0:
                    _chk(got[:ghost].tobytes() == want[lay["start"] - ghost:lay["start"]].tobytes(), "lower ghosts stale")
                if lay["hi_rank"] >= 0:
                    _chk(got[-ghost:].tobytes() == want[lay["stop"]:lay["stop"] + ghost].tobytes(), "upper ghosts stale")
                ph.check(ph.load().ph_h2d(a.ptr, loc.ctypes.data, loc.nbytes))          # reset for the second run
                ph.check(ph.load().ph_sync())
            a._buf.free(); b._buf.free()

    # ---- sharded reductions
    data = rs.randint(-8, 9, size=(8 * world + 1, 50, 30)).astype(np.float32)
    data[3, 2, 1] = data[-1, 4, 4] = 99.0                      # tie across shards
    r0, r1 = S.shard_range(data.shape[0], world, rank)
    x = D.from_host(data[r0:r1])
    off = r0 * 50 * 30
    _chk(S.reduce_full_sharded(x, "sum") == np.float32(data.sum(dtype=np.float64)), "tests/mgpu_check.py:110")
    _chk(S.reduce_full_sharded(x, "max") == data.max() and S.reduce_full_sharded(x, "min") == data.min(), "tests/mgpu_check.py:111")
    v, i = S.reduce_full_sharded(x, "argmax", off)
    _chk((v, i) == (np.float32(99.0), 3 * 1500 + 2 * 30 + 1), (v, i))
    for _ in range(5):                                          # call parity of the double-buffered slots
        _chk(S.reduce_full_sharded(x, "sum") == np.float32(data.sum(dtype=np.float64)), "tests/mgpu_check.py:115")
        _chk(S.reduce_full_sharded(x, "argmin", off)[0] == data.min(), "tests/mgpu_check.py:116")
    # every Crystal number type goes through the same combine
    for dt in (np.float64, np.int32, np.int64, np.uint8, np.int16, np.uint64):
        dd = rs.randint(0, 7, size=(3 * world + 2, 40)).astype(dt)
        dd[1, 5] = dd[-1, 7] = 9
        q0, q1 = S.shard_range(dd.shape[0], world, rank)
        xs = D.from_host(dd[q0:q1])
        total = dd.astype(np.int64).sum() if dd.dtype.kind != "f" else dd.sum()
        if dd.dtype.kind == "f" or total <= np.iinfo(dt).max:
            _chk(S.reduce_full_sharded(xs, "sum") == dt(total), "tests/mgpu_check.py:125")
        else:                                                   # UInt8: the global total leaves T -> every rank raises
            try:
                S.reduce_full_sharded(xs, "sum")
                _chk(False, f"{np.dtype(dt)} sum {total} must overflow")
            except ph.CrOverflowError:
                pass
        _chk(S.reduce_full_sharded(xs, "argmax", q0 * 40) == (dt(9), 1 * 40 + 5), "tests/mgpu_check.py:132")
        _chk(S.reduce_full_sharded(xs, "min") == dd.min(), "tests/mgpu_check.py:133")
    # a rank that owns NO rows contributes the identity (and still enters the collective)
    few = (rs.rand(max(1, world - 1), 6) + 1.0).astype(np.float32)                # all positive
    f0, f1 = S.shard_range(few.shape[0], world, rank)
    xf = D.from_host(few[f0:f1]) if f1 > f0 else D([0, 6], np.float32)
    _chk(S.reduce_full_sharded(xf, "min") == few.min() and S.reduce_full_sharded(xf, "max") == few.max(), "tests/mgpu_check.py:138")
    _chk(abs(float(S.reduce_full_sharded(xf, "sum")) - float(few.sum(dtype=np.float64))) <= 1e-4 * float(few.sum(dtype=np.float64)), "tests/mgpu_check.py:139")
    v, i = S.reduce_full_sharded(xf, "argmax", f0 * 6)
    _chk((v, i) == (few.max(), int(np.argmax(few.reshape(-1)))), "tests/mgpu_check.py:141")
    sf = S.ShardedNArray.from_global(few)
    _chk(sf.min(axis=0).to_host().tobytes() == few.min(axis=0).tobytes(), "tests/mgpu_check.py:143")
    _chk(sf.max(axis=0).to_host().tobytes() == few.max(axis=0).tobytes(), "tests/mgpu_check.py:144")
    try:
        S.reduce_full_sharded(D([0, 6], np.float32), "max")
        _chk(False, "max of an array that is empty on every rank must raise")
    except ph.CrEmptyError:
        pass
    _chk(S.reduce_full_sharded(D([0, 6], np.float32), "sum") == 0, "tests/mgpu_check.py:150")
    # integer sums are overflow-checked over the GLOBAL lexicographic fold, on every rank
    big32 = np.zeros((2 * world, 4), np.int32)
    big32[:, 0] = (2**31 - 1) // world                                           # every shard fits, the total fits too
    g0, g1 = S.shard_range(big32.shape[0], world, rank)
    xi = D.from_host(big32[g0:g1])
    try:
        S.reduce_full_sharded(xi, "sum")
        _chk(False, "global Int32 sum overflow must raise on every rank")
    except ph.CrOverflowError:
        pass
    pre = np.zeros((max(2, world), 2), np.int32)                                 # (one rank: both rows are its own)
    pre[0] = [2**31 - 2, 0]; pre[-1] = [5, -10]                                   # a PREFIX leaves Int32, the total does not
    p0, p1 = S.shard_range(pre.shape[0], world, rank)
    xp = D.from_host(pre[p0:p1])
    try:
        S.reduce_full_sharded(xp, "sum")
        _chk(False, "a prefix of the global fold leaves Int32: must raise")
    except ph.CrOverflowError:
        pass
    pre[-1] = [-10, 5]                                                            # same values, no prefix leaves Int32
    _chk(S.reduce_full_sharded(D.from_host(pre[p0:p1]), "sum") == np.int32(2**31 - 2 - 5), "same values, no prefix leaves Int32")
    si = S.ShardedNArray.from_global(big32)
    try:
        si.sum(axis=0)
        _chk(False, "Int32 axis-0 sum across ranks overflows: must raise")
    except ph.CrOverflowError:
        pass
    _chk(S.ShardedNArray.from_global(big32 // 4).sum(axis=0).to_host().tolist() == (big32 // 4).sum(axis=0).tolist(), "tests/mgpu_check.py:177")
    # NaN under max on ONE rank raises ArgumentError on EVERY rank (the flags travel with the partials)
    nn = np.ones((world, 8), np.float32)
    nn[world - 1, 3] = np.nan
    try:
        S.reduce_full_sharded(D.from_host(nn[rank:rank + 1]), "max")
        _chk(False, "NaN under max must raise on every rank")
    except ph.CrArgumentError:
        pass
    _chk(D.take_flags() == 0, "tests/mgpu_check.py:186")
    # axis-0 reduce: allreduce of the [outer*inner] partial
    part = x.sum(axis=0)
    ph.check(ph.load().ph_allreduce(ph.K["PH_SUM"], ph.K["PH_F32"], part.ptr, part.size))
    _chk(part.to_host().tobytes() == data.sum(axis=0, dtype=np.float64).astype(np.float32).tobytes(), "tests/mgpu_check.py:190")
    # ---- f-3: ShardedNArray behaves like the undivided array
    g = rs.randint(-8, 9, size=(6 * world + 2, 12, 10)).astype(np.float32)
    h = rs.randint(-8, 9, size=(6 * world + 2, 12, 10)).astype(np.float32)
    sg, sh = S.ShardedNArray.from_global(g), S.ShardedNArray.from_global(h)
    _chk(((sg * sh + sg) - 2.0).to_global().tobytes() == ((g * h + g) - np.float32(2.0)).tobytes(), "tests/mgpu_check.py:195")
    _chk((sg > sh).to_global().tobytes() == (g > h).tobytes(), "tests/mgpu_check.py:196")
    _chk(sg.sum() == np.float32(g.sum(dtype=np.float64)) and sg.max() == g.max(), "tests/mgpu_check.py:197")
    flat = int(np.argmax(g.reshape(-1)))
    _chk(sg.argmax() == (g.max(), list(np.unravel_index(flat, g.shape))), "tests/mgpu_check.py:199")
    _chk(sg.sum(axis=0).to_host().tobytes() == g.sum(axis=0, dtype=np.float64).astype(np.float32).tobytes(), "tests/mgpu_check.py:200")
    _chk(sg.max(axis=2).to_global().tobytes() == g.max(axis=2).tobytes(), "tests/mgpu_check.py:201")
    _chk(sg[ph.ALL, ph.rng(1, 9, 2), 3].to_global().tobytes() == np.ascontiguousarray(g[:, 1:10:2, 3]).tobytes(), "tests/mgpu_check.py:202")
    # transposes across shards (ph_alltoallv): default pattern = reversed axes, then two others
    _chk(sg.permute().to_global().tobytes() == np.ascontiguousarray(g.transpose(2, 1, 0)).tobytes(), "tests/mgpu_check.py:204")
    _chk(sg.permute(1, 0, 2).to_global().tobytes() == np.ascontiguousarray(g.transpose(1, 0, 2)).tobytes(), "tests/mgpu_check.py:205")
    _chk(sg.permute(0, 2, 1).to_global().tobytes() == np.ascontiguousarray(g.transpose(0, 2, 1)).tobytes(), "tests/mgpu_check.py:206")
    m2 = rs.rand(1000 * world + 3, 517).astype(np.float64)
    t2 = S.ShardedNArray.from_global(m2).permute()
    _chk(t2.shape == [517, 1000 * world + 3] and t2.to_global().tobytes() == np.ascontiguousarray(m2.T).tobytes(), "tests/mgpu_check.py:209")
    _chk(t2.permute().to_global().tobytes() == m2.tobytes(), "tests/mgpu_check.py:210")              # transposing twice is the identity
    if S.p2p_ready():                                                      # the P2P form reuses a result buffer on request
        again = S.ShardedNArray.from_global(m2 + 1.0).permute(out=t2)
        _chk(again is t2 and t2.to_global().tobytes() == np.ascontiguousarray((m2 + 1.0).T).tobytes(), "tests/mgpu_check.py:213")
    m3 = rs.rand(3 * world, 1)                                             # the result has ONE row: every rank but 0 owns nothing
    t3 = S.ShardedNArray.from_global(m3).permute()
    _chk(t3.shape == [1, 3 * world] and t3.to_global().tobytes() == np.ascontiguousarray(m3.T).tobytes(), "tests/mgpu_check.py:216")
    _chk(t3.permute().to_global().tobytes() == m3.tobytes(), "tests/mgpu_check.py:217")
    # slicing ACROSS shards (the result is re-split over the ranks): ranges, steps, reversal, an Int on the sharded axis
    for lit, npkey in (((ph.rng(2, 4 * world),), np.s_[2:4 * world + 1]), ((ph.rng(None, None, -1), ph.rng(1, 9, 2)), np.s_[::-1, 1:10:2]),
                       ((ph.rng(1, None, 3), ph.ALL, 4), np.s_[1::3, :, 4]), ((5, ph.rng(None, None, -1)), np.s_[5, ::-1]),
                       ((6 * world + 1, 3, 2), np.s_[6 * world + 1, 3, 2])):
        got = sg[lit]
        want = np.asarray(g[npkey])
        want = np.ascontiguousarray(want.reshape(want.shape if want.ndim else (1,)))
        _chk(got.shape == list(want.shape) and got.to_global().tobytes() == want.tobytes(), f"sharded slice {lit}")
    # ... and the scatter / fill twins: `narr[region] = sharded array | scalar` (the gather plan run backwards)
    for lit, npkey in (((ph.rng(2, 4 * world),), np.s_[2:4 * world + 1]), ((ph.rng(None, None, -1), ph.rng(1, 9, 2)), np.s_[::-1, 1:10:2]),
                       ((ph.rng(1, None, 3), ph.ALL, 4), np.s_[1::3, :, 4]), ((5, ph.rng(None, None, -1)), np.s_[5, ::-1]),
                       ((ph.ALL, 7), np.s_[:, 7])):
        dst_h = g.copy()
        src_h = np.ascontiguousarray(rs.randint(-8, 9, size=np.asarray(dst_h[npkey]).shape).astype(np.float32))
        dst = S.ShardedNArray.from_global(g)
        dst[lit] = S.ShardedNArray.from_global(src_h)
        dst_h[npkey] = src_h
        _chk(dst.to_global().tobytes() == dst_h.tobytes(), f"sharded scatter {lit}")
        dst[lit] = 2.5
        dst_h[npkey] = np.float32(2.5)
        _chk(dst.to_global().tobytes() == dst_h.tobytes(), f"sharded fill {lit}")
    try:
        S.ShardedNArray.from_global(g)[(ph.rng(0, 1),)] = S.ShardedNArray.from_global(g)
        _chk(False, "a source of the wrong shape must raise ShapeError")
    except ph.ShapeError:
        pass
    sg.set_mask(sg > sh, 0.0)
    _chk(sg.to_global().tobytes() == np.where(g > h, np.float32(0), g).tobytes(), "tests/mgpu_check.py:219")
    return {"ok": not _FAILED, "failed": list(_FAILED), "checks": _CHECKS, "world": world, "p2p": bool(p2p),
            "covered": ["heat 3-D small / big / 2-D slabs vs oracle and vs the 1-GPU run, 1 and 2 ghost planes, odd and even steps",
                        "the same in peer-mapped memory (in-kernel halos), ghost planes of the final state",
                        "sharded full reductions: every dtype, ties across shards, empty shards, integer overflow over the "
                        "global fold, NaN on one rank raising on every rank",
                        "ShardedNArray vs the undivided array: operators, comparisons, per-axis folds, slicing, permute "
                        "(3 patterns, twice = identity, empty shards, reused result), masked store"]}


if __name__ == "__main__":
    main()
