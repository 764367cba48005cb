#!/usr/bin/env python
"""Chunk schedules of pipeline.RowPipeline on the BASELINE elementwise config with HOST operands (the e2e leg of
bench.py): a*b+c, b the [1,8192] row vector, 8192x8192 f32, pinned host buffers, uploads + kernels + download
inside every timed step.  One JSON row per (chunks, taper); `link_ms` is the same bytes with no kernels.
    python benchmarks/bench_pipeline.py > gpurun_out/pipeline.jsonl"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ph_core_b200 as ph

ph.init(0)
lib = ph.load()
R, COLS = 8192, 8192
rs = np.random.RandomState(7)
a_pin = ph.pinned_from((rs.rand(R, COLS) * 2 - 1).astype(np.float32))
c_pin = ph.pinned_from((rs.rand(R, COLS) * 2 - 1).astype(np.float32))
b_pin = ph.pinned_from((rs.rand(1, COLS) * 2 - 1).astype(np.float32))
out_pin = ph.pinned_empty((R, COLS), np.float32)
BYTES = 5 * R * COLS * 4 + COLS * 4
expr = lambda x, z, y: x.broadcast_op("*", y) + z
ms = C.c_float()
ref = None
for chunks, taper, ups, ramp in [(16, 0, 1, 0), (4, 7, 1, 0), (4, 7, 1, 3), (4, 7, 1, 5), (8, 6, 1, 3), (3, 7, 1, 4)]:
    pipe = ph.pipeline.RowPipeline(chunks=chunks, taper=taper, uploaders=ups, ramp=ramp)
    step = lambda: pipe.map_rows(expr, rows=[a_pin, c_pin], out=out_pin, shared=[b_pin], wait=False)
    step(); ph.sync()
    if ref is None:
        ref = out_pin.copy()
    ts = []
    for _ in range(3):
        ph.check(lib.ph_timer_start())
        for _ in range(4):
            step()
        ph.check(lib.ph_timer_stop(C.byref(ms)))
        ts.append(ms.value / 4)
    ph.sync()
    ok = bool(out_pin.tobytes() == ref.tobytes())
    print(json.dumps({"chunks": chunks, "taper": taper, "uploaders": ups, "ramp": ramp, "n_chunks": len(ph.pipeline.row_chunks(R, chunks, taper, ramp)),
                      "ms_best": round(min(ts), 4), "ms_median": round(sorted(ts)[1], 4),
                      "gbs": round(BYTES / (sorted(ts)[1] * 1e-3) / 1e9, 2), "same_result": ok}), flush=True)
    pipe.close()
