"""Multi-GPU partitioning of the hot path: one process per GPU.

Host logic only (which rows / planes a rank owns, who its neighbours are, how per-rank partial
results combine) plus thin wrappers that drive the NCCL entry points of libphgpu.  The
reference has no distributed layer; only what BASELINE.json's north_star partitions is here:
elementwise ops and reductions split along the leading axis, the heat grid is slab-decomposed
along axis 0 with one (or, for the two-steps-per-pass stencil, two) ghost planes per side.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import K, check
from .region import ALL as ALL_


def _host_check(status: int) -> None:
    from .narray import host_check
    host_check(status)


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `n` leading-axis indices: the first n % world ranks get one extra
    (ph_shard_range, include/ph_host.h -- the plans below are C++ host code; this file marshals)."""
    a, b = C.c_int64(), C.c_int64()
    _host_check(_lib.load().ph_shard_range(int(n), int(world), int(rank), C.byref(a), C.byref(b)))
    return a.value, b.value


def slab_layout(n0: int, world: int, rank: int, ghost: int = 1) -> dict:
    """Slab of a grid split along axis 0 (ph_slab_layout): owned planes [start, stop) live at local planes
    [ghost, ghost + count); `ghost` planes on either side are ghosts (1: one time step per halo
    exchange; 2: two, for the temporally blocked stencil).  A rank at either end of the grid has
    no neighbour there: its first / last owned plane is the fixed global boundary."""
    out = _lib.PhSlab()
    _host_check(_lib.load().ph_slab_layout(int(n0), int(world), int(rank), int(ghost), C.byref(out)))
    return {"start": out.start, "stop": out.stop, "count": out.count, "ghost": out.ghost,
            "local_planes": out.local_planes, "lo_rank": out.lo_rank, "hi_rank": out.hi_rank}


def slab_from_global(field: np.ndarray, world: int, rank: int, ghost: int = 1) -> np.ndarray:
    """Local slab (with ghost planes filled from the neighbouring planes of `field`)."""
    lay = slab_layout(field.shape[0], world, rank, ghost)
    loc = np.zeros((lay["local_planes"],) + field.shape[1:], dtype=field.dtype)
    loc[ghost:ghost + lay["count"]] = field[lay["start"]:lay["stop"]]
    if lay["lo_rank"] >= 0:
        loc[:ghost] = field[lay["start"] - ghost:lay["start"]]
    if lay["hi_rank"] >= 0:
        loc[ghost + lay["count"]:] = field[lay["stop"]:lay["stop"] + ghost]
    return loc


def transpose_plan(shape: Sequence[int], pattern: Sequence[int], world: int, rank: int) -> dict:
    """Host plan of `permute(pattern)` on an array sharded along axis 0 whose result is sharded
    along ITS axis 0 (= old axis k = pattern[0]).  For every peer q:
      send[q] = (k0, k1): the slice of old axis k this rank cuts out of its rows for q (q's share
                of the new leading axis); the block is sent already permuted, so its shape is
                send_shape[q];
      recv[q] = (p0, p1): the old-axis-0 rows q owns; its block lands at positions p0..p1-1 of new
                axis j (where old axis 0 ends up), shape recv_shape[q].
    pattern[0] == 0 needs no exchange (`local` = True)."""
    shape = [int(v) for v in shape]
    pattern = [int(v) for v in pattern]
    if len(pattern) != len(shape):
        raise IndexError(f"permute pattern {pattern} is not a permutation of the axes of a rank-{len(shape)} array")
    plan_c = _lib.PhTransposePlan()
    peers = (_lib.PhTransposePeer * max(1, int(world)))()
    st = _lib.load().ph_transpose_plan_of((C.c_int64 * len(shape))(*shape), len(shape), (C.c_int32 * len(pattern))(*pattern),
                                          int(world), int(rank), C.byref(plan_c), peers)
    if st == K["PH_HOST_INDEX_ERROR"]:
        raise IndexError(f"permute pattern {pattern} is not a permutation of the axes of a rank-{len(shape)} array")
    _host_check(st)
    nd = len(shape)
    new_shape = [int(plan_c.new_shape[i]) for i in range(nd)]
    if plan_c.local:
        return {"local": True, "new_shape": new_shape}
    plan = {"local": False, "new_shape": new_shape, "k": plan_c.k, "j": plan_c.j,
            "my_rows": (plan_c.my_rows[0], plan_c.my_rows[1]), "my_new_rows": (plan_c.my_new_rows[0], plan_c.my_new_rows[1]),
            "send": [], "recv": [], "send_shape": [], "recv_shape": []}
    for q in range(world):
        p = peers[q]
        plan["send"].append((p.send0, p.send1)); plan["recv"].append((p.recv0, p.recv1))
        plan["send_shape"].append([int(p.send_shape[i]) for i in range(nd)])
        plan["recv_shape"].append([int(p.recv_shape[i]) for i in range(nd)])
    return plan


def slice_plan(shape: Sequence[int], literal: Sequence, world: int, rank: int) -> dict:
    """Host plan of `narr[region_literal]` (IndexRegion.new, src/index_region.cr:192-224) on an array sharded along
    axis 0, result sharded along ITS axis 0 (ph_slice_plan_of, include/ph_host.h -- the compiled host layers use the
    same plan).  `local` = True when the literal leaves axis 0 whole (nothing moves).  Otherwise, for every peer q:
      send[q] = (extents, strides, offset) of the block this rank owes q, as a strided view of ITS local shard
                (element units; None when there is nothing) -- an arithmetic progression of its rows;
      land[q] = (extents, strides, offset) of where that block lands in q's shard of the result (a contiguous
                range of q's rows);
      recv[q] = (j_lo, j_hi): the rows of the result (global numbering) q holds for this rank.
    The leading axis of the result is the first axis the literal does not index with an Int: axis 0 itself (its
    selected rows are re-split over the ranks) or, when axis 0 is ONE row, a later axis (the row's owner deals it out)."""
    from .narray import make_region, host_check, _i64
    shape = [int(v) for v in shape]
    reg = make_region(list(literal), shape, True)                         # raises IndexError / DimensionError like the reference
    plan_c = _lib.PhSlicePlan()
    peers = (_lib.PhSlicePeer * max(1, int(world)))()
    host_check(_lib.load().ph_slice_plan_of(_i64(shape), len(shape), C.byref(reg), int(world), int(rank), C.byref(plan_c), peers))
    new_shape = [int(plan_c.new_shape[i]) for i in range(plan_c.dims)]
    if plan_c.local:
        return {"local": True, "new_shape": new_shape}
    plan = {"local": False, "new_shape": new_shape, "my_new_rows": (int(plan_c.my_new_rows[0]), int(plan_c.my_new_rows[1])),
            "send": [], "land": [], "recv": []}

    def unpack(d):
        n = d.rank
        ext = [int(d.extent[i]) for i in range(n)]
        return (ext, [int(d.stride[i]) for i in range(n)], int(d.offset)) if all(ext) else None

    for q in range(world):
        plan["send"].append(unpack(peers[q].send))
        plan["land"].append(unpack(peers[q].land))
        plan["recv"].append((int(peers[q].recv0), int(peers[q].recv1)))
    return plan


def combine_extremum(values: Sequence, indices: Sequence[int], is_max: bool = True):
    """Per-rank (value, global lex index) pairs -> the FIRST extremum: best value, then the
    lowest index (README.md:56-61 semantics across shards)."""
    best_v, best_i = None, None
    for v, i in zip(values, indices):
        if i < 0:
            continue                      # empty shard
        if best_v is None or (v > best_v if is_max else v < best_v) or (v == best_v and i < best_i):
            best_v, best_i = v, i
    return best_v, best_i


# ---------------------------------------------------------------- NCCL-backed wrappers (GPU)
_comm = {"world": 1, "rank": 0, "ready": False}


EXTREMUM_RECORD_BYTES = 32


def pack_extremum_record(value, local_index: int, row_offset_elems: int, dtype) -> np.ndarray:
    """Host twin of the 32-byte record a rank contributes to the argmax / argmin allgather:
    value @0 (<= 8 bytes), local flat index @16 (int64, -1 = empty shard), elements owned by lower
    ranks @24 (int64).  The device fills @0 and @16 (ph_reduce_full_dev), the host @24."""
    rec = np.zeros(EXTREMUM_RECORD_BYTES, dtype=np.uint8)
    dt = np.dtype(dtype)
    rec[:dt.itemsize] = np.array([value], dtype=dt).view(np.uint8)
    rec[16:24] = np.array([local_index], dtype=np.int64).view(np.uint8)
    rec[24:32] = np.array([row_offset_elems], dtype=np.int64).view(np.uint8)
    return rec


def parse_extremum_records(raw: np.ndarray, dtype, world: int):
    """The gathered records -> (values, GLOBAL flat indices) per rank; an empty shard keeps -1."""
    dt = np.dtype(dtype)
    vals, idxs = [], []
    for r in range(world):
        chunk = raw[EXTREMUM_RECORD_BYTES * r: EXTREMUM_RECORD_BYTES * (r + 1)]
        vals.append(chunk[:dt.itemsize].view(dt)[0])
        i = int(chunk[16:24].view(np.int64)[0])
        idxs.append(i + int(chunk[24:32].view(np.int64)[0]) if i >= 0 else -1)
    return vals, idxs


def combine_extremum_records(raw: np.ndarray, dtype, world: int, is_max: bool = True):
    """The gathered 32-byte records -> (value, GLOBAL flat index) of the FIRST extremum: best value, then the
    lowest global index (ph_combine_extremum_records, include/ph_host.h).  (None, None) if every shard is empty."""
    from .narray import dtype_code
    dt = np.dtype(dtype)
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    winner, gidx = C.c_int32(-1), C.c_int64(-1)
    _host_check(_lib.load().ph_combine_extremum_records(raw.ctypes.data, int(world), dtype_code(dt), int(bool(is_max)),
                                                        C.byref(winner), C.byref(gidx)))
    if winner.value < 0:
        return None, None
    rec = raw[EXTREMUM_RECORD_BYTES * winner.value: EXTREMUM_RECORD_BYTES * (winner.value + 1)]
    return rec[:dt.itemsize].view(dt)[0], gidx.value


def comm_init(dist=None) -> Tuple[int, int]:
    """Create libphgpu's communicator: rank 0 makes the NCCL unique id, `torch.distributed`
    (any backend) only carries those 128 bytes to the other ranks."""
    import torch
    lib = _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    _lib.init()
    ident = (C.c_uint8 * 128)()
    if world > 1:
        if dist is None:
            import torch.distributed as dist
        if rank == 0:
            check(lib.ph_comm_unique_id(ident))
        t = torch.tensor(list(ident), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
    check(lib.ph_comm_init(world, rank, ident))
    _comm.update(world=world, rank=rank, ready=True)
    return world, rank


def world_rank() -> Tuple[int, int]:
    return _comm["world"], _comm["rank"]


def p2p_ready() -> bool:
    """True when every rank mapped its peers' memory (CUDA IPC): sharded reductions combine inside the
    reduction kernel and ph_heat_run_sharded delivers its own halos; False = the NCCL forms run."""
    out = C.c_int32(0)
    check(_lib.load().ph_comm_p2p_ready(C.byref(out)))
    return bool(out.value)


class _SymmBuffer:
    """Peer-mapped device memory (ph_symm_alloc).  Allocation and release are COLLECTIVE, so release is
    explicit (`free()`, every rank, same order) and never left to the garbage collector; whatever is
    still allocated is released by ph_comm_destroy."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(_lib.load().ph_symm_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self) -> None:
        if self.ptr:
            check(_lib.load().ph_symm_free(self.ptr))
            self.ptr = None


def symm_empty(shape, dtype):
    """An uninitialised DeviceNArray in peer-mapped memory (collective: every rank calls it)."""
    from .narray import DeviceNArray
    dt = np.dtype(dtype)
    n = int(np.prod([int(v) for v in shape], dtype=np.int64)) if len(shape) else 0
    return DeviceNArray(shape, dt, _SymmBuffer(max(1, n) * dt.itemsize))


def symm_from_host(arr: np.ndarray):
    """`DeviceNArray.from_host` into peer-mapped memory (collective)."""
    arr = np.ascontiguousarray(arr)
    out = symm_empty(arr.shape, arr.dtype)
    if arr.size:
        check(_lib.load().ph_h2d(out.ptr, arr.ctypes.data, arr.nbytes))
        check(_lib.load().ph_sync())
    return out


_RED_OUT = ((C.c_uint64 * 2)(), C.c_int64(-1), C.c_uint32(0))


def reduce_full_sharded(local, name: str, row_offset_elems: int = 0):
    """Full reduction of an array sharded along axis 0 (ph_reduce_full_sharded; collective -- every rank
    calls it, also with an EMPTY shard, which contributes the identity).  sum / min / max return the value;
    argmax / argmin return (value, GLOBAL flat index of the first extremum).  `row_offset_elems` = number of
    elements owned by lower ranks.  Raises like the undivided array would: EmptyError for min / max of
    nothing, OverflowError for an integer sum leaving T over the global lexicographic fold, ArgumentError
    for NaN under max -- on EVERY rank (the flags travel with the partials)."""
    from .narray import dtype_code, _RED, raise_for_flags, CrEmptyError
    dt = np.dtype(np.uint8) if local.dtype == np.dtype(np.bool_) else local.dtype
    val, idx, flags = _RED_OUT                               # reused ctypes cells: the call itself is ~0.1 ms
    st = _lib.load().ph_reduce_full_sharded(K[_RED[name]], dtype_code(dt), local.ptr, C.byref(local.desc()),
                                            int(row_offset_elems), val, C.byref(idx), C.byref(flags))
    if st:
        check(st)
    if flags.value:
        raise_for_flags(flags.value)
    value = np.frombuffer(val, dtype=dt, count=1)[0]
    if name == "sum":
        return value
    if idx.value < 0:
        raise CrEmptyError("Empty enumerable")
    return (value, idx.value) if name.startswith("arg") else value


def heat_run_sharded(slab, other, coeff, steps: int, ghost: int = 1):
    """`steps` steps on this rank's slab (`ghost` ghost planes per side included) with the halo
    exchange overlapped with the interior update; returns the buffer holding the final state.
    ghost = 2 lets a rank-3 grid advance two time steps per pass over HBM and per exchange."""
    from .narray import dtype_code
    lib = _lib.load()
    c = np.array(coeff, dtype=slab.dtype)
    ext = (C.c_int64 * len(slab.shape))(*[int(s) for s in slab.shape])
    final_is_b = C.c_int32(0)
    check(lib.ph_heat_run_sharded(dtype_code(slab.dtype), len(slab.shape), ext, c.ctypes.data, int(ghost), slab.ptr,
                                  other.ptr, int(steps), C.byref(final_is_b)))
    return other if final_is_b.value else slab


# ---------------------------------------------------------------- f-3: a sharded NArray
class ShardedNArray:
    """An NArray distributed along axis 0 over the ranks of the job: every
    rank holds the contiguous row range `shard_range(shape[0], world, rank)` as an ordinary
    DeviceNArray.  Elementwise ops, comparisons and masked stores are purely local; full
    reductions combine per-GPU partials (allreduce / allgather of (value, index) pairs);
    per-axis reductions allreduce only when the reduced axis is the sharded one; slicing is
    local as long as axis 0 is taken whole; `permute` moves data between shards with one
    personalised all-to-all (ph_alltoallv)."""

    def __init__(self, global_shape: Sequence[int], local):
        self.shape = [int(s) for s in global_shape]
        self.local = local
        self.world, self.rank = world_rank()
        self.row0, self.row1 = shard_range(self.shape[0], self.world, self.rank)
        if list(local.shape) != [self.row1 - self.row0] + self.shape[1:]:
            raise ValueError(f"local shard has shape {local.shape}, expected {[self.row1 - self.row0] + self.shape[1:]}")

    @property
    def dtype(self):
        return self.local.dtype

    @classmethod
    def from_global(cls, host: np.ndarray) -> "ShardedNArray":
        """Every rank passes the same host array and keeps its own rows."""
        from .narray import DeviceNArray
        world, rank = world_rank()
        a, b = shard_range(host.shape[0], world, rank)
        return cls(host.shape, DeviceNArray.from_host(np.ascontiguousarray(host[a:b])))

    def to_global(self) -> np.ndarray:
        """Assemble the whole array on every rank's host (allgather of the shards)."""
        from .narray import _Buffer
        lib = _lib.load()
        mine = self.local.to_host()
        if self.world == 1:
            return mine
        row_elems = int(np.prod(self.shape[1:], dtype=np.int64)) if len(self.shape) > 1 else 1
        rows_max = -(-self.shape[0] // self.world)
        nbytes = rows_max * row_elems * self.dtype.itemsize
        send = _Buffer(max(1, nbytes))
        recv = _Buffer(max(1, nbytes * self.world))
        if mine.nbytes:
            check(lib.ph_d2d(send.ptr, self.local.ptr, mine.nbytes))
        check(lib.ph_allgather(send.ptr, recv.ptr, nbytes))
        raw = np.zeros(nbytes * self.world, dtype=np.uint8)
        check(lib.ph_d2h(raw.ctypes.data, recv.ptr, raw.nbytes))
        parts = []
        for r in range(self.world):
            a, b = shard_range(self.shape[0], self.world, r)
            n = (b - a) * row_elems * self.dtype.itemsize
            parts.append(raw[r * nbytes: r * nbytes + n].view(self.dtype).reshape([b - a] + self.shape[1:]))
        return np.concatenate(parts, axis=0)

    # ---- elementwise / compare: local, shapes checked on the GLOBAL shape ------------------
    def _wrap(self, local, shape=None):
        return ShardedNArray(self.shape if shape is None else shape, local)

    def _peer(self, other):
        from .narray import ShapeError
        if isinstance(other, ShardedNArray):
            if other.shape != self.shape:
                raise ShapeError(f"The shape of this MultiIndexable ({self.shape}) does not match the shape of the "
                                 f"one provided ({other.shape}).")
            return other.local
        return other

    def __add__(self, o): return self._wrap(self.local + self._peer(o))
    def __sub__(self, o): return self._wrap(self.local - self._peer(o))
    def __mul__(self, o): return self._wrap(self.local * self._peer(o))
    def __truediv__(self, o): return self._wrap(self.local / self._peer(o))
    def __gt__(self, o): return self._wrap(self.local > self._peer(o))
    def __lt__(self, o): return self._wrap(self.local < self._peer(o))
    def __ge__(self, o): return self._wrap(self.local >= self._peer(o))
    def __le__(self, o): return self._wrap(self.local <= self._peer(o))
    def eq(self, o): return self._wrap(self.local.eq(self._peer(o)))

    def set_mask(self, mask: "ShardedNArray", value) -> None:
        self.local.set_mask(mask.local, self._peer(value))

    def __getitem__(self, key):
        """`narr[region_literal]` on the distributed array (gather, src/multi_indexable.cr:338-356, across
        shards).  A literal that leaves axis 0 whole is local: the result stays sharded the same way.  Anything
        else changes who owns what -- the result is sharded along ITS axis 0 -- and is one redistribution
        (`slice_plan`): every (source rank, destination rank) block is an arithmetic progression of the source's
        rows, i.e. ONE strided descriptor, and lands as a contiguous row range of the destination's shard.  With
        peer-mapped memory the gather kernels store the blocks straight into their owners (ph_alltoall_strided,
        one pass); otherwise the blocks are gathered locally and exchanged with ph_alltoallv, received in place."""
        from .narray import DeviceNArray, DeviceView
        from ._lib import PhDesc
        if not isinstance(key, tuple):
            key = (key,)
        plan = slice_plan(self.shape, list(key), self.world, self.rank)
        new_shape = plan["new_shape"]
        if plan["local"]:
            loc = self.local.get_chunk([ALL_] + list(key[1:])) if self.row1 > self.row0 else \
                DeviceNArray([0] + new_shape[1:], self.dtype)
            return ShardedNArray(new_shape, loc)
        lib = _lib.load()
        isz = self.dtype.itemsize
        m0, m1 = plan["my_new_rows"]
        my_shape = [m1 - m0] + new_shape[1:]
        rnk = len(new_shape)
        if self.world > 1 and p2p_ready() and not os.environ.get("PH_PERMUTE_NCCL"):
            res = symm_empty(my_shape, self.dtype)                          # collective
            srcs, dsts = (PhDesc * self.world)(), (PhDesc * self.world)()
            for q in range(self.world):
                sd, dd = plan["send"][q], plan["land"][q]
                srcs[q] = PhDesc.make(*sd) if sd else PhDesc.make([0] * rnk, [0] * rnk, 0)
                dsts[q] = PhDesc.make(*dd) if dd else PhDesc.make([0] * rnk, [0] * rnk, 0)
            check(lib.ph_alltoall_strided(isz, self.local.ptr, srcs, res.ptr, dsts))
            return ShardedNArray(new_shape, res)
        res = DeviceNArray(my_shape, self.dtype)
        row_bytes = int(np.prod(new_shape[1:], dtype=np.int64)) * isz
        vp = C.c_void_p
        sends, sp, sb, rp, rb = [], [], [], [], []
        for q in range(self.world):
            sd = plan["send"][q]
            if sd:
                blk = DeviceView(self.local._buf, PhDesc.make(*sd), list(sd[0]), self.dtype).to_narr()
                sends.append(blk); sp.append(blk.ptr); sb.append(blk.size * isz)
            else:
                sp.append(None); sb.append(0)
            lo, hi = plan["recv"][q]                                        # what q holds for me: a contiguous row range
            rp.append(res.ptr + (lo - m0) * row_bytes if hi > lo else None)
            rb.append((hi - lo) * row_bytes)
        check(lib.ph_alltoallv((vp * self.world)(*sp), (C.c_int64 * self.world)(*sb), (vp * self.world)(*rp), (C.c_int64 * self.world)(*rb)))
        del sends
        return ShardedNArray(new_shape, res)

    def set_chunk(self, literal: Sequence, value) -> None:
        """`narr[region_literal] = value` across shards (scatter / fill, src/multi_writable.cr:55-84).  A scalar
        fills this rank's part of the region (no exchange).  A ShardedNArray of the region's shape is the gather
        run backwards with the same plan: the rows of `value` this rank holds leave as contiguous blocks
        (ph_alltoallv) and every block received is scattered into the arithmetic progression of local rows it
        belongs to (one strided copy per peer)."""
        from .narray import DeviceNArray, ShapeError
        from ._lib import PhDesc
        literal = list(literal)
        plan = slice_plan(self.shape, literal, self.world, self.rank)
        lib = _lib.load()
        isz = self.dtype.itemsize
        if isinstance(value, ShardedNArray):
            if value.shape != plan["new_shape"]:                       # multi_writable.cr:58-60
                raise ShapeError(f"Cannot substitute: the given array has shape {value.shape}, but the region has "
                                 f"shape {plan['new_shape']}.")
            if value.dtype != self.dtype:
                raise TypeError("device path: source and destination must share a dtype")
            if plan["local"]:
                if self.row1 > self.row0:
                    self.local.set_chunk([ALL_] + literal[1:], value.local)
                return
            m0, _ = plan["my_new_rows"]
            row_bytes = int(np.prod(plan["new_shape"][1:], dtype=np.int64)) * isz
            vp = C.c_void_p
            temps, sp, sb, rp, rb = [], [], [], [], []
            for q in range(self.world):
                lo, hi = plan["recv"][q]                               # rows of `value` I hold that q's shard receives
                sp.append(value.local.ptr + (lo - m0) * row_bytes if hi > lo and row_bytes else None)
                sb.append((hi - lo) * row_bytes if hi > lo else 0)
                sd = plan["send"][q]                                   # where q's rows land in MY shard
                t = DeviceNArray(list(sd[0]), self.dtype) if sd else None
                temps.append(t)
                rp.append(t.ptr if t is not None else None)
                rb.append(t.size * isz if t is not None else 0)
            check(lib.ph_alltoallv((vp * self.world)(*sp), (C.c_int64 * self.world)(*sb), (vp * self.world)(*rp), (C.c_int64 * self.world)(*rb)))
            for q, t in enumerate(temps):
                if t is not None:
                    d = PhDesc.make(*plan["send"][q])
                    check(lib.ph_copy_strided(isz, t.ptr, C.byref(t.desc()), self.local.ptr, C.byref(d)))
            return
        scalar = np.array(value, dtype=self.dtype)
        if plan["local"]:
            if self.row1 > self.row0:
                self.local.set_chunk([ALL_] + literal[1:], value)
            return
        for q in range(self.world):                                    # my cells of the region, one block per destination
            sd = plan["send"][q]
            if sd:
                d = PhDesc.make(*sd)
                check(lib.ph_fill_region(isz, self.local.ptr, C.byref(d), scalar.ctypes.data))

    def __setitem__(self, key, value) -> None:
        if isinstance(key, ShardedNArray):                             # narr[mask] = value
            self.set_mask(key, value)
            return
        self.set_chunk(list(key) if isinstance(key, tuple) else [key], value)

    # ---- transposes across shards: the one real exchange step (all-to-all) --------------------
    def permute(self, *pattern, out: Optional["ShardedNArray"] = None) -> "ShardedNArray":
        """MultiIndexable#permute (src/multi_indexable.cr:795-803; default = reversed axes,
        transforms.cr:236-238) on the distributed array; the result is sharded along its own axis 0.
        With peer-mapped memory (p2p_ready) the exchange is ONE pass of peer stores: for every peer the
        block this rank owes it -- a permuted VIEW of its rows -- is copied by the transpose kernel
        straight into that peer's shard of the result over NVLink (ph_alltoall_strided): no staging, no
        NCCL, no scatter.  The result then lives in peer-mapped memory (collective allocation); pass a
        previous result as `out` to reuse it.  Otherwise (NCCL form): every rank permutes each block
        locally (strided gather), exchanges them with ph_alltoallv and scatters what it receives."""
        from .narray import DeviceNArray
        from ._lib import PhDesc
        from .region import rng, ALL
        nd = len(self.shape)
        pat = list(pattern[0]) if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)) else list(pattern)
        if not pat:
            pat = list(reversed(range(nd)))
        plan = transpose_plan(self.shape, pat, self.world, self.rank)
        if plan["local"]:
            return ShardedNArray(plan["new_shape"], self.local.permute(*pat))      # axis 0 stays put: no exchange
        lib = _lib.load()
        k, j = plan["k"], plan["j"]
        m0, m1 = plan["my_new_rows"]
        new_shape = plan["new_shape"]
        if self.world > 1 and p2p_ready() and not os.environ.get("PH_PERMUTE_NCCL"):
            my_shape = [m1 - m0] + new_shape[1:]
            if out is not None:
                if out.shape != new_shape or out.dtype != self.dtype or not isinstance(out.local._buf, _SymmBuffer):
                    raise ValueError("permute(out=): `out` must be an earlier P2P result of the same shape and dtype")
                res = out.local
            else:
                res = symm_empty(my_shape, self.dtype)                           # collective
            r0, r1 = self.row0, self.row1
            srcs, dsts = (PhDesc * self.world)(), (PhDesc * self.world)()
            for q in range(self.world):
                k0, k1 = plan["send"][q]
                q0, q1 = shard_range(new_shape[0], self.world, q)                # q's rows of the result
                q_shape = [q1 - q0] + new_shape[1:]
                if k1 - k0 <= 0 or r1 - r0 <= 0:
                    srcs[q] = PhDesc.make([0] * nd, [0] * nd, 0)
                    dsts[q] = PhDesc.make([0] * nd, [0] * nd, 0)
                    continue
                lit = [ALL] * nd
                lit[k] = rng(k0, k1 - 1)
                srcs[q] = self.local.view(*lit).permute(*pat).desc()             # my rows x q's slice, in q's axis order
                strides = [int(np.prod(q_shape[i + 1:], dtype=np.int64)) for i in range(nd)]
                ext = list(q_shape)
                ext[j] = r1 - r0                                                 # old axis 0 (my rows) lands on new axis j
                dsts[q] = PhDesc.make(ext, strides, r0 * strides[j])
            check(lib.ph_alltoall_strided(self.dtype.itemsize, self.local.ptr, srcs, res.ptr, dsts))
            return out if out is not None else ShardedNArray(new_shape, res)
        out = DeviceNArray([m1 - m0] + plan["new_shape"][1:], self.dtype)
        isz = self.dtype.itemsize
        sends, recvs = [], []
        for q in range(self.world):
            k0, k1 = plan["send"][q]
            n_send = int(np.prod(plan["send_shape"][q], dtype=np.int64))
            if q == self.rank:
                # my own block never leaves the GPU: one permuting copy straight into the result
                # (a transposed view scattered into a region) instead of gather + self-send + scatter
                if n_send:
                    lit = [ALL] * nd
                    lit[k] = rng(k0, k1 - 1)
                    p0, p1 = plan["recv"][q]
                    dst = [ALL] * nd
                    dst[j] = rng(p0, p1 - 1)
                    out.set_chunk(dst, self.local.view(*lit).permute(*pat))
                sends.append(None)
                recvs.append(None)
                continue
            if n_send:
                lit = [ALL] * nd
                lit[k] = rng(k0, k1 - 1)
                blk = self.local.view(*lit).permute(*pat).to_narr()       # contiguous, already permuted
            else:
                blk = None
            n_recv = int(np.prod(plan["recv_shape"][q], dtype=np.int64))
            sends.append(blk)
            recvs.append(DeviceNArray(plan["recv_shape"][q], self.dtype) if n_recv else None)
        vp = C.c_void_p
        sp = (vp * self.world)(*[b.ptr if b is not None else None for b in sends])
        rp = (vp * self.world)(*[b.ptr if b is not None else None for b in recvs])
        sb = (C.c_int64 * self.world)(*[b.size * isz if b is not None else 0 for b in sends])
        rb = (C.c_int64 * self.world)(*[b.size * isz if b is not None else 0 for b in recvs])
        check(lib.ph_alltoallv(sp, sb, rp, rb))
        for q in range(self.world):
            if recvs[q] is None:
                continue
            p0, p1 = plan["recv"][q]
            lit = [ALL] * nd
            lit[j] = rng(p0, p1 - 1)
            out.set_chunk(lit, recvs[q])
        return ShardedNArray(plan["new_shape"], out)

    # ---- reductions -----------------------------------------------------------------------
    def _row_elems(self):
        return int(np.prod(self.shape[1:], dtype=np.int64)) if len(self.shape) > 1 else 1

    def sum(self, axis=None):
        return self._reduce("sum", axis)

    def max(self, axis=None):
        return self._reduce("max", axis)

    def min(self, axis=None):
        return self._reduce("min", axis)

    def argmin(self):
        v, i = reduce_full_sharded(self.local, "argmin", self.row0 * self._row_elems())
        coord = []
        for length in reversed(self.shape):
            coord.append(i % length)
            i //= length
        return v, list(reversed(coord))

    def argmax(self):
        v, i = reduce_full_sharded(self.local, "argmax", self.row0 * self._row_elems())
        coord = []
        for length in reversed(self.shape):
            coord.append(i % length)
            i //= length
        return v, list(reversed(coord))

    def _reduce(self, name, axis):
        from .narray import dtype_code, _RED, DeviceNArray, CrEmptyError, CrIndexError, _Buffer
        if axis is None:
            return reduce_full_sharded(self.local, name, self.row0 * self._row_elems())
        if axis < 0 or axis >= len(self.shape):
            raise CrIndexError(f"axis {axis} is not present in a {len(self.shape)}-dimensional MultiIndexable")
        if axis != 0:
            part = getattr(self.local, name)(axis=axis)
            return ShardedNArray([self.shape[0]] + part.shape[1:], part)   # kept axis 0: still sharded
        # axis 0 is the sharded one: partial over my rows, then a combine across ranks.  Every decision that
        # can raise is taken on the GLOBAL shape first, so all ranks reach the collective (or none does).
        if self.shape[0] == 0 and name != "sum":
            raise CrEmptyError("Empty enumerable")
        out_shape = self.shape[1:] or [1]
        if self.row1 - self.row0 > 0:
            part = self.local._reduce_axis(name, 0, raise_now=self.world == 1)   # flags are read once, after the combine
        else:                                                            # an empty shard contributes the identity
            dt = self.dtype
            if name == "sum":
                ident = 0
            elif dt.kind == "f":
                ident = -np.inf if name == "max" else np.inf
            else:
                info = np.iinfo(dt)
                ident = info.min if name == "max" else info.max
            part = DeviceNArray.fill(out_shape, ident, dt)
        lib = _lib.load()
        if name == "sum" and self.dtype.kind in "iu" and self.world > 1 and not p2p_ready():
            # integer sums are overflow-CHECKED: an ncclSum would wrap silently.  The per-rank partials
            # ([world, inner], rank order = row order) are gathered and folded by the checked axis-0 sum.
            # (With peer-mapped memory ph_allreduce itself folds in rank order with checked adds.)
            nbytes = part.size * part.dtype.itemsize
            gathered = DeviceNArray([self.world] + out_shape, part.dtype, _Buffer(max(1, nbytes * self.world)))
            check(lib.ph_allgather(part.ptr, gathered.ptr, nbytes))
            return gathered.sum(axis=0)
        check(lib.ph_allreduce(K[_RED[name]], dtype_code(part.dtype), part.ptr, part.size))
        if self.world > 1:
            DeviceNArray.raise_pending()       # per-axis folds are raise points (narray._reduce_axis): the cross-rank fold too
        return part                            # replicated DeviceNArray
