#!/usr/bin/env python
"""Sustained 3-D stencil rate: `--warm` + `--steps` time steps of a Z,Y,X f32 grid between two resident buffers,
timed with CUDA events over the whole run (long enough for the power cap to settle, unlike the 10-step rows of
bench_kernels.py); SM clock sampled meanwhile.  One JSON line.  The plan is picked by the PH_HEAT_* variables
(read once per process), so a sweep runs this once per variant:
    PH_HEAT_TB_CFG=5 python benchmarks/bench_heat_sustained.py --shape 2048,2048,2048"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, heat

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="2048,2048,2048")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--warm", type=int, default=10)
args = ap.parse_args()
shape = [int(v) for v in args.shape.split(",")]
ph.init(0)
lib = ph.load()
a, b = D(shape, np.float32), D(shape, np.float32)
# a non-constant field: plane z holds z mod 7 + a ramp along x (cheap to build on the device)
ramp = D.from_host((np.arange(shape[2], dtype=np.float32) % 13).reshape(1, 1, shape[2]))
zs = D.from_host((np.arange(shape[0], dtype=np.float32) % 7).reshape(shape[0], 1, 1))
ph.check(lib.ph_fill_region(4, a.ptr, C.byref(a.desc()), np.array(0.0, np.float32).ctypes.data))
a = a.broadcast_op("+", ramp).broadcast_op("+", zs)
ph.check(lib.ph_sync())
clocks = []
stop = threading.Event()


def sample():
    while not stop.is_set():
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            clocks.append((float(o[0]), float(o[1])))
        except Exception:
            pass
        time.sleep(0.05)


cur = heat.simulate_into(a, b, 0.1, args.warm)
other = b if cur is a else a
ph.check(lib.ph_sync())
th = threading.Thread(target=sample, daemon=True)
th.start()
ms = C.c_float()
ph.check(lib.ph_timer_start())
cur = heat.simulate_into(cur, other, 0.1, args.steps)
ph.check(lib.ph_timer_stop(C.byref(ms)))
stop.set()
th.join()
cells = int(np.prod(shape))
knobs = {k: v for k, v in os.environ.items() if k.startswith("PH_HEAT")}
mid = sorted(c[0] for c in clocks)[len(clocks) // 2] if clocks else None
pw = sorted(c[1] for c in clocks)[len(clocks) // 2] if clocks else None
print(json.dumps({"shape": shape, "steps": args.steps, "warm": args.warm, "knobs": knobs, "ms_per_step": round(ms.value / args.steps, 4),
                  "gcell_per_s": round(cells * args.steps / (ms.value * 1e-3) / 1e9, 1), "sm_mhz_median": mid, "power_w_median": pw,
                  "checksum": f"{cur.checksum64(0):016x}"}), flush=True)
