#!/usr/bin/env python
"""Per-axis reductions on axis-0 SHARDS of the 1e9-element f32 array ([1000/N, 1000, 1000], the per-GPU
shapes of BASELINE configs[3] at N = 1, 2, 4, 8): kernel time of ph_reduce_axis, 5 back-to-back launches
between two CUDA events.  PH_AXIS_E / PH_AXIS_U / PH_AXIS_BLOCK override the strip kernel's plan (sweeps)."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D
ph.init(0); lib = ph.load()
axes = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2").split(",")]
tag = {k: os.environ[k] for k in ("PH_AXIS_E", "PH_AXIS_U", "PH_AXIS_BLOCK") if k in os.environ}
for rows in (1000, 500, 250, 125):
    x = D([rows, 1000, 1000], np.float32)
    ph.check(lib.ph_fill_region(4, x.ptr, C.byref(x.desc()), np.array(0.5, np.float32).ctypes.data))
    for axis in axes:
        oshape = [s for i, s in enumerate(x.shape) if i != axis]
        for name in ("sum", "argmax"):
            out = D(oshape, np.int64 if name == "argmax" else np.float32)
            red = ph.K["PH_SUM"] if name == "sum" else ph.K["PH_ARGMAX"]
            call = lambda: ph.check(lib.ph_reduce_axis(red, ph.K["PH_F32"], x.ptr, C.byref(x.desc()), axis, out.ptr, C.byref(out.desc())))
            call(); ts = []
            for _ in range(5):
                ms = C.c_float(); ph.check(lib.ph_timer_start())
                for _ in range(5):
                    call()
                ph.check(lib.ph_timer_stop(C.byref(ms))); ts.append(ms.value / 5)
            t = sorted(ts)[len(ts) // 2]
            print(json.dumps({"shape": [rows, 1000, 1000], "axis": axis, "red": name, "ms": round(t, 4),
                              "gbs": round(x.size * 4 / t / 1e6, 1), **tag}), flush=True)
    del x
