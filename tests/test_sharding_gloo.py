"""CPU, world_size 2, gloo: the N > 1 host logic -- axis-0 shard ranges, slab layout with
ghost planes, halo neighbours, and the combination rules for sharded reductions.  The
compute inside each rank is the ORACLE (this is a test of the partitioning logic, which is
what runs on the host at N > 1); 1-vs-N agreement must be bit-identical for the stencil and
max/argmax, and exact for exactly-summable sums."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ph_core_b200 import sharding as S
    from oracle import ph_oracle as O
    rs = np.random.RandomState(42)
    field = (rs.rand(11, 9, 8) * 100).astype(np.float32)          # 11 planes: uneven split 6 + 5
    steps = 5
    # ---- heat: slab decomposition + halo exchange over gloo send/recv
    lay = S.slab_layout(field.shape[0], world, rank)
    cur = S.slab_from_global(field, world, rank)
    for _ in range(steps):
        # the full-slab oracle step holds plane 0 / -1 fixed: exactly the ghost-plane contract;
        # an owned global-boundary plane (no neighbour) must additionally stay fixed
        nxt = O.heat_step_nd(cur, np.float32(0.1))
        if lay["lo_rank"] < 0:
            nxt[1] = cur[1]
        if lay["hi_rank"] < 0:
            nxt[-2] = cur[-2]
        reqs = []
        lo_recv, hi_recv = torch.zeros(nxt[0].shape), torch.zeros(nxt[0].shape)
        if lay["lo_rank"] >= 0:
            reqs.append(dist.isend(torch.from_numpy(nxt[1].copy()), lay["lo_rank"]))
            reqs.append(dist.irecv(lo_recv, lay["lo_rank"]))
        if lay["hi_rank"] >= 0:
            reqs.append(dist.isend(torch.from_numpy(nxt[-2].copy()), lay["hi_rank"]))
            reqs.append(dist.irecv(hi_recv, lay["hi_rank"]))
        for r in reqs:
            r.wait()
        if lay["lo_rank"] >= 0:
            nxt[0] = lo_recv.numpy()
        if lay["hi_rank"] >= 0:
            nxt[-1] = hi_recv.numpy()
        cur = nxt
    np.save(os.path.join(out_dir, f"heat_{rank}.npy"), cur[1:-1])
    # ---- heat with TWO ghost planes per side: two time steps per 2-plane exchange (the layout the
    #      temporally blocked kernel uses, ph_heat_run_sharded ghost_planes = 2)
    g = 2
    lay = S.slab_layout(field.shape[0], world, rank, ghost=g)
    cur = S.slab_from_global(field, world, rank, ghost=g)
    for _ in range(3):                                             # 3 passes = 6 steps
        nxt = cur
        for _ in range(2):
            prev, nxt = nxt, O.heat_step_nd(nxt, np.float32(0.1))
            if lay["lo_rank"] < 0:
                nxt[:g + 1] = prev[:g + 1]                          # unused ghosts + the fixed global boundary plane
            if lay["hi_rank"] < 0:
                nxt[-g - 1:] = prev[-g - 1:]
        # after two local steps only the owned planes are valid: refresh both ghost planes
        reqs = []
        lo_recv, hi_recv = torch.zeros(nxt[:g].shape), torch.zeros(nxt[:g].shape)
        if lay["lo_rank"] >= 0:
            reqs.append(dist.isend(torch.from_numpy(nxt[g:2 * g].copy()), lay["lo_rank"]))
            reqs.append(dist.irecv(lo_recv, lay["lo_rank"]))
        if lay["hi_rank"] >= 0:
            reqs.append(dist.isend(torch.from_numpy(nxt[-2 * g:-g].copy()), lay["hi_rank"]))
            reqs.append(dist.irecv(hi_recv, lay["hi_rank"]))
        for r in reqs:
            r.wait()
        if lay["lo_rank"] >= 0:
            nxt[:g] = lo_recv.numpy()
        if lay["hi_rank"] >= 0:
            nxt[-g:] = hi_recv.numpy()
        cur = nxt
    np.save(os.path.join(out_dir, f"heat2_{rank}.npy"), cur[g:-g])
    # ---- transpose across shards: the plan of ShardedNArray.permute driven with numpy blocks
    #      and gloo send/recv standing in for ph_alltoallv
    tsrc = rs.randint(0, 1000, size=(7, 5, 6)).astype(np.int32)     # uneven split 4 + 3 on axis 0
    for pat in ([2, 1, 0], [1, 0, 2], [2, 0, 1], [0, 2, 1]):
        plan = S.transpose_plan(tsrc.shape, pat, world, rank)
        r0, r1 = S.shard_range(tsrc.shape[0], world, rank)
        mine = tsrc[r0:r1]
        if plan["local"]:
            res = np.ascontiguousarray(mine.transpose(pat))
        else:
            k, j = plan["k"], plan["j"]
            m0, m1 = plan["my_new_rows"]
            res = np.zeros([m1 - m0] + plan["new_shape"][1:], np.int32)
            reqs, bufs = [], {}
            for q in range(world):
                k0, k1 = plan["send"][q]
                sl = [slice(None)] * 3; sl[k] = slice(k0, k1)
                blk = np.ascontiguousarray(mine[tuple(sl)].transpose(pat))
                assert list(blk.shape) == plan["send_shape"][q]
                if q == rank:
                    bufs[q] = torch.from_numpy(blk.copy())
                else:
                    reqs.append(dist.isend(torch.from_numpy(blk), q))
                    bufs[q] = torch.zeros(plan["recv_shape"][q], dtype=torch.int32)
                    reqs.append(dist.irecv(bufs[q], q))
            for rq in reqs:
                rq.wait()
            for q in range(world):
                p0, p1 = plan["recv"][q]
                sl = [slice(None)] * 3; sl[j] = slice(p0, p1)
                res[tuple(sl)] = bufs[q].numpy()
        np.save(os.path.join(out_dir, f"perm_{''.join(map(str, pat))}_{rank}.npy"), res)
    # ---- slicing across shards: the plan of ShardedNArray#[] (ph_slice_plan_of) driven with numpy strided views,
    #      gloo send/recv standing in for the peer stores of ph_alltoall_strided
    from ph_core_b200 import rng, ALL
    r0, r1 = S.shard_range(tsrc.shape[0], world, rank)
    mine = np.ascontiguousarray(tsrc[r0:r1]).reshape(-1)
    for tag, lit in (("rev", [rng(None, None, -1), rng(0, None, 2)]), ("step", [rng(1, None, 2), ALL, 3]), ("row", [5, rng(None, None, -1)])):
        plan = S.slice_plan(tsrc.shape, lit, world, rank)
        m0, m1 = plan["my_new_rows"]
        res = np.full([m1 - m0] + plan["new_shape"][1:], -1, np.int32)
        flat = res.reshape(-1)
        reqs, inbox = [], {}
        for q in range(world):
            sd = plan["send"][q]
            blk = None
            if sd:
                ext, strd, off = sd
                blk = np.array([mine[off + sum(c * t for c, t in zip(idx, strd))] for idx in np.ndindex(*ext)], np.int32).reshape(ext)
            lo, hi = plan["recv"][q]
            if q == rank:
                inbox[q] = blk
            else:
                if blk is not None:
                    reqs.append(dist.isend(torch.from_numpy(blk.copy()), q))
                if hi > lo:
                    inbox[q] = torch.zeros([hi - lo] + plan["new_shape"][1:], dtype=torch.int32)
                    reqs.append(dist.irecv(inbox[q], q))
        for rq in reqs:
            rq.wait()
        for q in range(world):
            lo, hi = plan["recv"][q]
            if hi > lo:
                got = inbox[q] if isinstance(inbox[q], np.ndarray) else inbox[q].numpy()
                res[lo - m0:hi - m0] = got.reshape([hi - lo] + plan["new_shape"][1:])
        np.save(os.path.join(out_dir, f"slice_{tag}_{rank}.npy"), res)
    # ---- reductions over axis-0 shards
    data = rs.randint(-8, 9, size=(10, 7)).astype(np.float32)
    data[3, 2] = data[8, 1] = 50.0                                 # tie across the two shards
    a, b = S.shard_range(data.shape[0], world, rank)
    local = data[a:b]
    part = torch.tensor([float(local.sum(dtype=np.float64))], dtype=torch.float64)
    dist.all_reduce(part)
    v, i = O.reduce_argmax(local)
    # the same 32-byte record the GPU path allgathers (value @0, LOCAL index @16, row offset @24)
    rec = torch.from_numpy(S.pack_extremum_record(v, int(i), a * data.shape[1], np.float32))
    recs = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(recs, rec)
    vals, idxs = S.parse_extremum_records(np.concatenate([r.numpy() for r in recs]), np.float32, world)
    best = S.combine_extremum(vals, idxs, True)
    axis0 = torch.from_numpy(O.reduce_axis(local, 0, "sum").astype(np.float64))
    dist.all_reduce(axis0)
    np.save(os.path.join(out_dir, f"red_{rank}.npy"), np.array([part.item(), best[0], best[1]] + axis0.tolist()))
    dist.destroy_process_group()


def test_two_rank_partitioning(tmp_path):
    from oracle import ph_oracle as O
    from ph_core_b200 import sharding as S
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(42)
    field = (rs.rand(11, 9, 8) * 100).astype(np.float32)
    want = field.copy()
    for _ in range(5):
        want = O.heat_step_nd(want, np.float32(0.1))
    got = np.concatenate([np.load(tmp_path / f"heat_{r}.npy") for r in range(world)])
    assert got.tobytes() == want.tobytes()                          # slabbing changes no cell's arithmetic
    want = O.heat_step_nd(want, np.float32(0.1))                    # 6 steps for the 2-ghost-plane run
    got2 = np.concatenate([np.load(tmp_path / f"heat2_{r}.npy") for r in range(world)])
    assert got2.tobytes() == want.tobytes()
    tsrc = rs.randint(0, 1000, size=(7, 5, 6)).astype(np.int32)
    for pat in ([2, 1, 0], [1, 0, 2], [2, 0, 1], [0, 2, 1]):
        got = np.concatenate([np.load(tmp_path / f"perm_{''.join(map(str, pat))}_{r}.npy") for r in range(world)])
        assert got.tobytes() == np.ascontiguousarray(tsrc.transpose(pat)).tobytes(), pat
    for tag, key in (("rev", np.s_[::-1, ::2]), ("step", np.s_[1::2, :, 3]), ("row", np.s_[5, ::-1])):
        got = np.concatenate([np.load(tmp_path / f"slice_{tag}_{r}.npy") for r in range(world)])
        assert got.tobytes() == np.ascontiguousarray(tsrc[key]).tobytes(), tag
    data = rs.randint(-8, 9, size=(10, 7)).astype(np.float32)
    data[3, 2] = data[8, 1] = 50.0
    for r in range(world):
        red = np.load(tmp_path / f"red_{r}.npy")
        assert red[0] == data.sum(dtype=np.float64)
        assert (red[1], int(red[2])) == (50.0, 3 * 7 + 2)           # lower flat index wins the tie
        assert red[3:].tolist() == data.sum(axis=0).tolist()


def test_shard_ranges_cover_exactly():
    from ph_core_b200 import sharding as S
    for n in [0, 1, 7, 8, 1000, 2048]:
        for world in [1, 2, 3, 4, 8]:
            ranges = [S.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    lay = S.slab_layout(2048, 8, 0); assert lay["lo_rank"] == -1 and lay["hi_rank"] == 1 and lay["local_planes"] == 258
    lay = S.slab_layout(2048, 8, 7); assert lay["hi_rank"] == -1 and lay["start"] == 1792
    lay = S.slab_layout(2048, 8, 3, ghost=2); assert lay["local_planes"] == 260 and lay["ghost"] == 2
    assert S.combine_extremum([3.0, 9.0, 9.0], [5, 40, 12], True) == (9.0, 12)
    assert S.combine_extremum([3.0, 1.0, 1.0], [5, 40, 12], False) == (1.0, 12)
    assert S.combine_extremum([3.0, 0.0], [5, -1], True) == (3.0, 5)    # empty shard ignored
    raw = np.concatenate([S.pack_extremum_record(7, 11, 0, np.int64), S.pack_extremum_record(0, -1, 40, np.int64),
                          S.pack_extremum_record(7, 2, 40, np.int64)])
    vals, idxs = S.parse_extremum_records(raw, np.int64, 3)
    assert [int(v) for v in vals] == [7, 0, 7] and idxs == [11, -1, 42]
    assert S.combine_extremum(vals, idxs, True) == (7, 11)


def test_transpose_plan_covers_every_element_once():
    """Host plan of the cross-shard permute, all ranks simulated in one process with numpy: for
    random shapes / patterns / world sizes the blocks sent by rank r to rank q are exactly the
    blocks q expects from r, and the assembled result equals numpy's transpose."""
    from ph_core_b200 import sharding as S
    rs = np.random.RandomState(3)
    for _ in range(60):
        nd = int(rs.randint(2, 5))
        shape = [int(rs.randint(1, 7)) for _ in range(nd)]
        pat = list(rs.permutation(nd))
        world = int(rs.randint(1, 6))
        g = rs.randint(0, 10_000, size=shape).astype(np.int32)
        plans = [S.transpose_plan(shape, pat, world, r) for r in range(world)]
        want = np.ascontiguousarray(g.transpose(pat))
        outs = []
        for q in range(world):
            r0, r1 = S.shard_range(shape[0], world, q)
            if plans[q]["local"]:
                outs.append(np.ascontiguousarray(g[r0:r1].transpose(pat)))
                continue
            m0, m1 = plans[q]["my_new_rows"]
            res = np.full([m1 - m0] + plans[q]["new_shape"][1:], -1, np.int32)
            for r in range(world):                                  # what r sends to q
                a0, a1 = S.shard_range(shape[0], world, r)
                k0, k1 = plans[r]["send"][q]
                sl = [slice(None)] * nd; sl[plans[r]["k"]] = slice(k0, k1)
                blk = np.ascontiguousarray(g[a0:a1][tuple(sl)].transpose(pat))
                assert list(blk.shape) == plans[r]["send_shape"][q] == plans[q]["recv_shape"][r]
                p0, p1 = plans[q]["recv"][r]
                assert (p0, p1) == (a0, a1)
                dl = [slice(None)] * nd; dl[plans[q]["j"]] = slice(p0, p1)
                res[tuple(dl)] = blk
            outs.append(res)
        got = np.concatenate(outs, axis=0) if outs else want
        assert got.tobytes() == want.tobytes(), (shape, pat, world)
    with pytest.raises(IndexError):
        S.transpose_plan([3, 4], [0, 0], 2, 0)


# ---- the partition plans are C++ host code (include/ph_host.h); independent Python restatements check them
def _py_shard_range(n, world, rank):
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _py_transpose_plan(shape, pattern, world, rank):
    new_shape = [shape[a] for a in pattern]
    if pattern[0] == 0:
        return {"local": True, "new_shape": new_shape}
    k, j = pattern[0], pattern.index(0)
    r0, r1 = _py_shard_range(shape[0], world, rank)
    m0, m1 = _py_shard_range(shape[k], world, rank)
    plan = {"local": False, "new_shape": new_shape, "k": k, "j": j, "my_rows": (r0, r1), "my_new_rows": (m0, m1),
            "send": [], "recv": [], "send_shape": [], "recv_shape": []}
    for q in range(world):
        k0, k1 = _py_shard_range(shape[k], world, q)
        p0, p1 = _py_shard_range(shape[0], world, q)
        ss = list(new_shape); ss[0] = k1 - k0; ss[j] = r1 - r0
        rs = list(new_shape); rs[0] = m1 - m0; rs[j] = p1 - p0
        plan["send"].append((k0, k1)); plan["recv"].append((p0, p1))
        plan["send_shape"].append(ss); plan["recv_shape"].append(rs)
    return plan


def test_cpp_partition_plans_match_their_restatement():
    from ph_core_b200 import sharding as S
    rs = np.random.RandomState(5)
    for _ in range(400):
        world = int(rs.randint(1, 9))
        n = int(rs.randint(0, 50))
        for rank in range(world):
            assert S.shard_range(n, world, rank) == _py_shard_range(n, world, rank)
            ghost = int(rs.randint(1, 3))
            lay = S.slab_layout(n, world, rank, ghost)
            a, b = _py_shard_range(n, world, rank)
            assert lay == {"start": a, "stop": b, "count": b - a, "ghost": ghost, "local_planes": b - a + 2 * ghost,
                           "lo_rank": rank - 1 if rank > 0 else -1, "hi_rank": rank + 1 if rank < world - 1 else -1}
    for _ in range(300):
        nd = int(rs.randint(1, 5))
        shape = [int(v) for v in rs.randint(0, 12, size=nd)]
        pattern = [int(v) for v in rs.permutation(nd)]
        world = int(rs.randint(1, 9))
        for rank in range(world):
            assert S.transpose_plan(shape, pattern, world, rank) == _py_transpose_plan(shape, pattern, world, rank)
    for bad in ([0, 0], [0, 2], [1], [0, 1, 2]):
        with pytest.raises(IndexError):
            S.transpose_plan([4, 5], bad, 2, 0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int8, np.int16, np.int32, np.int64,
                                   np.uint8, np.uint16, np.uint32, np.uint64])
def test_cpp_extremum_combine_matches_the_python_fold(dtype):
    from ph_core_b200 import sharding as S
    rs = np.random.RandomState(9)
    info = np.iinfo(dtype) if np.dtype(dtype).kind in "iu" else None
    for trial in range(200):
        world = int(rs.randint(1, 9))
        recs, vals, idxs = [], [], []
        offset = 0
        for r in range(world):
            count = int(rs.randint(0, 6))                          # elements this shard owns (0 = empty shard)
            if info is not None:
                pool = [info.min, info.max, 0, 1] if trial % 3 == 0 else [3, 5, 7]
                v = dtype(pool[rs.randint(len(pool))])
            else:
                v = dtype([-0.0, 0.0, 1.5, -2.5, np.inf][rs.randint(5)])
            local = int(rs.randint(0, count)) if count else -1
            recs.append(S.pack_extremum_record(v, local, offset, dtype))
            vals.append(v); idxs.append(local + offset if local >= 0 else -1)
            offset += count
        raw = np.concatenate(recs)
        for is_max in (True, False):
            want = S.combine_extremum(vals, idxs, is_max)
            got = S.combine_extremum_records(raw, dtype, world, is_max)
            if want[0] is None:
                assert got == (None, None)
            else:
                assert got[1] == want[1] and got[0] == want[0] and np.signbit(np.float64(got[0])) == np.signbit(np.float64(want[0]))
        pv, pi = S.parse_extremum_records(raw, dtype, world)
        assert pi == idxs and all(a == b for a, b in zip(pv, vals))
