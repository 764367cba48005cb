#!/usr/bin/env python
"""Slab-decomposed 3-D heat run under torchrun: Gcell-updates/s for a given grid, steps and ghost
width (max over ranks, device timers).  python -m torch.distributed.run --nproc-per-node N
benchmarks/bench_heat_sharded.py --grid 2048,2048,2048 --steps 10 --ghost 2"""
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
ap.add_argument("--grid", default="2048,2048,2048")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--ghost", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ph.init(local)
lib = ph.load()
world, rank = S.comm_init(dist)
G = [int(v) for v in args.grid.split(",")]
lay = S.slab_layout(G[0], world, rank, ghost=args.ghost)
shape = [lay["local_planes"]] + G[1:]
a, b = D(shape, np.float32), D(shape, np.float32)
one = np.array(1.0, np.float32)
for buf in (a, b):
    ph.check(lib.ph_fill_region(4, buf.ptr, C.byref(buf.desc()), one.ctypes.data))
S.heat_run_sharded(a, b, 0.1, 2, ghost=args.ghost)
best = None
for _ in range(args.reps):
    dist.barrier(); torch.cuda.synchronize()
    ms = C.c_float()
    ph.check(lib.ph_timer_start())
    S.heat_run_sharded(a, b, 0.1, args.steps, ghost=args.ghost)
    ph.check(lib.ph_timer_stop(C.byref(ms)))
    t = torch.tensor([ms.value], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    v = float(t.item()) / args.steps
    best = v if best is None else min(best, v)
if rank == 0:
    cells = float(np.prod(G))
    print(json.dumps({"grid": G, "gpus": world, "ghost": args.ghost, "steps": args.steps, "ms_per_step": round(best, 4),
                      "gcell_per_s": round(cells / (best * 1e-3) / 1e9, 1),
                      "no_overlap": bool(os.environ.get("PH_HEAT_NO_OVERLAP"))}), flush=True)
dist.destroy_process_group()
