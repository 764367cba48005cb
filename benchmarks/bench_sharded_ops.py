#!/usr/bin/env python
"""Sharded NArray operations under torchrun (one rank per GPU), device timers, max over ranks:
  * BASELINE configs[3]: full sum / max / argmax and axis-0 / axis-1 sums of a 1e9-element f32
    array sharded along axis 0 (allreduce of a scalar, allgather of (value, index) pairs,
    allreduce of the [1000,1000] partial, no collective);
  * ShardedNArray.permute of a 16384^2 f64 matrix (transposed copy across
    shards = local permuting gathers + ph_alltoallv + scatters).
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bench_sharded_ops.py"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, sharding as S

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ph.init(local)
lib = ph.load()
world, rank = S.comm_init(dist)


def fill(d, value):
    v = np.array(value, d.dtype)
    ph.check(lib.ph_fill_region(d.dtype.itemsize, d.ptr, C.byref(d.desc()), v.ctypes.data))


def timed(fn, reps=args.reps, warm=2):
    for _ in range(warm):
        fn()
    best = None
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        ms = C.c_float()
        ph.check(lib.ph_timer_start())
        fn()
        ph.check(lib.ph_timer_stop(C.byref(ms)))
        t = torch.tensor([ms.value], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    return best


def emit(name, nbytes_global, ms, **extra):
    if rank == 0:
        print(json.dumps({"op": name, "gpus": world, "ms": round(ms, 4), "global_bytes": int(nbytes_global),
                          "aggregate_gbs": round(nbytes_global / (ms * 1e-3) / 1e9, 1),
                          "per_gpu_gbs": round(nbytes_global / world / (ms * 1e-3) / 1e9, 1), **extra}), flush=True)


# ---- config 3: 1e9 f32 as [1000,1000,1000], axis-0 shards
R = [1000, 1000, 1000] if not args.quick else [64 * world, 1000, 1000]
r0, r1 = S.shard_range(R[0], world, rank)
loc = D([r1 - r0] + R[1:], np.float32)
fill(loc, 0.5)
x = S.ShardedNArray(R, loc)
nb = float(np.prod(R)) * 4
want = 0.5 * float(np.prod(R))
got = x.sum()
ok = abs(float(got) - want) <= 1e-4 * want
emit("sharded full sum f32 %s (allreduce of one scalar)" % R, nb, timed(lambda: x.sum()), result_ok=bool(ok))
emit("sharded full max f32 (allreduce)", nb, timed(lambda: x.max()))
emit("sharded full argmax f32 (allgather of (value, index))", nb, timed(lambda: x.argmax()))
emit("sharded axis-0 sum f32 (allreduce of the [1000,1000] partial)", nb, timed(lambda: x.sum(axis=0)))
emit("sharded axis-0 max f32 (allreduce of the [1000,1000] partial)", nb, timed(lambda: x.max(axis=0)))
emit("sharded axis-1 sum f32 (no collective)", nb, timed(lambda: x.sum(axis=1)))
del x, loc

# ---- f-3: transposed copy across shards, 16384^2 f64
N = 16384 if not args.quick else 4096
m0, m1 = S.shard_range(N, world, rank)
mat = D([m1 - m0, N], np.float64)
fill(mat, float(rank + 1))
sm = S.ShardedNArray([N, N], mat)
t = sm.permute()
# every rank's rows of the transpose hold, in column block q, the fill value of rank q
probe = [float(t.local.get(0, S.shard_range(N, world, q)[0])) for q in range(world)]
p2p_perm = S.p2p_ready() and not os.environ.get("PH_PERMUTE_NCCL")        # the P2P form writes into a reused symmetric result
emit("sharded permute (transpose across shards) f64 %dx%d" % (N, N), 2.0 * N * N * 8,
     timed((lambda: sm.permute(out=t)) if p2p_perm else (lambda: sm.permute()), reps=3, warm=1),
     result_ok=bool(probe == [float(q + 1) for q in range(world)]),
     transport="peer stores (ph_alltoall_strided)" if p2p_perm else "ncclSend/ncclRecv (ph_alltoallv)",
     note="algorithmic bytes = read + write of the matrix once")
dist.destroy_process_group()
