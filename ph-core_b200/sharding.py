"""Multi-GPU partitioning of the hot path: one process per GPU.

Host logic only (which rows / planes a rank owns, who its neighbours are, how per-rank partial
results combine) plus thin wrappers that drive the NCCL entry points of libphgpu.  The
reference has no distributed layer; only what BASELINE.json's north_star partitions is here:
elementwise ops and reductions split along the leading axis, the heat grid is slab-decomposed
along axis 0 with one ghost plane per side.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import K, check


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `n` leading-axis indices: the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def slab_layout(n0: int, world: int, rank: int) -> dict:
    """Slab of a grid split along axis 0: owned planes [start, stop) live at local planes
    [1, 1 + count); local planes 0 and count + 1 are ghosts.  A rank at either end of the
    grid has no neighbour there: its first / last owned plane is the fixed global boundary."""
    start, stop = shard_range(n0, world, rank)
    return {"start": start, "stop": stop, "count": stop - start, "local_planes": stop - start + 2,
            "lo_rank": rank - 1 if rank > 0 else -1, "hi_rank": rank + 1 if rank < world - 1 else -1}


def slab_from_global(field: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Local slab (with ghost planes filled from the neighbouring planes of `field`)."""
    lay = slab_layout(field.shape[0], world, rank)
    loc = np.zeros((lay["local_planes"],) + field.shape[1:], dtype=field.dtype)
    loc[1:-1] = field[lay["start"]:lay["stop"]]
    if lay["lo_rank"] >= 0:
        loc[0] = field[lay["start"] - 1]
    if lay["hi_rank"] >= 0:
        loc[-1] = field[lay["stop"]]
    return loc


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


def reduce_full_sharded(local, name: str, row_offset_elems: int = 0):
    """Full reduction of an array sharded along axis 0: local two-pass reduce, then
    allreduce of the per-GPU partial (sum/min/max) or allgather of (value, index) pairs
    (argmax/argmin).  `row_offset_elems` = number of elements owned by lower ranks."""
    from .narray import DeviceNArray, dtype_code, _Buffer, _RED
    lib = _lib.load()
    world, rank = world_rank()
    dt = local.dtype
    res = _Buffer(64)
    idx_ptr = res.ptr + 16
    check(lib.ph_reduce_full_dev(K[_RED[name]], dtype_code(dt), local.ptr, C.byref(local.desc()), res.ptr, idx_ptr))
    if name in ("sum", "min", "max"):
        check(lib.ph_allreduce(K[_RED[name]], dtype_code(dt), res.ptr, 1))
        out = np.zeros(1, dtype=dt)
        check(lib.ph_d2h(out.ctypes.data, res.ptr, dt.itemsize))
        DeviceNArray.raise_pending()
        return out[0]
    gathered = _Buffer(32 * world)
    check(lib.ph_allgather(res.ptr, gathered.ptr, 32))
    raw = np.zeros(32 * world, dtype=np.uint8)
    check(lib.ph_d2h(raw.ctypes.data, gathered.ptr, raw.nbytes))
    DeviceNArray.raise_pending()
    vals, idxs = [], []
    offs = _allgather_host_int(row_offset_elems)
    for r in range(world):
        chunk = raw[32 * r: 32 * (r + 1)]
        vals.append(chunk[:dt.itemsize].view(dt)[0])
        i = int(chunk[16:24].view(np.int64)[0])
        idxs.append(i + offs[r] if i >= 0 else -1)
    return combine_extremum(vals, idxs, is_max=(name == "argmax"))


def _allgather_host_int(v: int) -> List[int]:
    world, rank = world_rank()
    if world == 1:
        return [int(v)]
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v)], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [int(o.item()) for o in outs]


def heat_run_sharded(slab, other, coeff, steps: int):
    """`steps` steps on this rank's slab (ghost planes included) with the halo exchange
    overlapped with the interior update; returns the buffer holding the final state."""
    from .narray import dtype_code
    lib = _lib.load()
    c = np.array(coeff, dtype=slab.dtype)
    ext = (C.c_int64 * len(slab.shape))(*[int(s) for s in slab.shape])
    check(lib.ph_heat_run_sharded(dtype_code(slab.dtype), len(slab.shape), ext, c.ctypes.data, slab.ptr, other.ptr,
                                  int(steps)))
    return other if steps % 2 else slab
