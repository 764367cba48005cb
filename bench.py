#!/usr/bin/env python
"""bench.py -- headline benchmark of the ph-core device path on B200.

Workload (BASELINE.json configs[1]): elementwise a*b+c with row-vector broadcasting on
8192x8192 Float32 NArrays.  One "step" = the reference-faithful execution of `a * b + c`:
two kernels with a materialised, individually rounded temporary
    t = a * b      (b is [1, 8192]; algorithmic bytes 2*N*4 + 32 KiB = 536.9 MB)
    out = t + c    (3*N*4 = 805.3 MB)
=> 1342.2 MB of algorithmic traffic per step per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
N > 1 is launched by torchrun, one rank per GPU; elementwise work shards along the leading
axis with no data-path collective (weak scaling: every rank owns an 8192x8192 shard).

Prints ONE JSON line (rank 0).  `value` = whole-job GB/s with inputs resident in HBM (two CUDA
events around the K steps); `e2e` = the same metric through the public ARRAY API with HOST buffers
(ph_core_b200.pipeline.RowPipeline: from_host_async of a, b, c, the two operators, to_host_async of the
result, chunked over the library's streams; `e2e.naive` = plain from_host -> operators -> to_host beside it,
`e2e.host_link_ceiling` = the measured rate of moving the same bytes with no kernels at all); `roofline` describes the
dominant kernel (its own duration from events on every 8th step of the same timed region, DRAM
traffic from the committed ncu capture); `cpu_baseline` = the oracle's C port of the REFERENCE'S
structure (single thread: ph-core has no threads) timed on this box, with the flat OpenMP loop on
all cores reported beside it; `extras` = the 2048^3 heat stencil and the 1e9-element sum (BASELINE
configs 3 and 4), each with its own CPU baseline at N = 1.
`--impl reference` times that same reference-structured port as the reference arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS = 8192, 8192
N = ROWS * COLS
BYTES_MUL = 2 * N * 4 + COLS * 4          # read a, write t, read b once
BYTES_ADD = 3 * N * 4                     # read t, read c, write out
BYTES_STEP = BYTES_MUL + BYTES_ADD
SEED = 20261017
METRIC = "f32 elementwise HBM GB/s"   # BASELINE.json metric, first clause (the heat clause is in extras)
WORKLOAD = ("elementwise a*b+c, b=[1,8192] row-vector broadcast, 8192x8192 f32 "
            "(two reference-faithful kernels)")
# dram__bytes_read.sum + dram__bytes_write.sum of one `out = t + c` launch of this very command,
# captured in ONE ncu pass with the caches left alone (profiles/r02_bench_dram_warm.csv):
# 477.6 MB read (59 MB of the temporary come from L2: alternating traversal + L2 eviction priorities;
# round 1 without the priorities: 504.0 MB) + 269.8 MB written.  The cold-cache `--set full` capture
# (profiles/r01_ncu_bench_kernels.csv) reads 536.9 MB = the algorithmic bytes.
NCU_TRAFFIC_ADD = 747322368


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_inputs(rank: int, rows: int = ROWS):
    """Philox counter RNG, seed 20261017, stream = tensor id."""
    def gen(stream, shape):
        g = np.random.Generator(np.random.Philox(key=SEED, counter=[0, 0, stream, rank]))
        return (g.random(shape, dtype=np.float32) * 2 - 1).astype(np.float32)
    return gen(1, (rows, COLS)), gen(2, (1, COLS)), gen(3, (rows, COLS))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = threading.Event()

    def _start_nvml(self) -> bool:
        """NVML in-process, one sample every ~5 ms: the multi-GPU legs last tens of milliseconds, too short for a
        freshly started nvidia-smi loop (its first line arrives after ~100 ms)."""
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = [("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap)]

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                        pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
                        r = int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if r & b else "Not Active" for _, b in bits])
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.nvml = N
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return True
        except Exception:
            self.nvml = None
            return False

    def start(self):
        if self._start_nvml():
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvml, 5 ms" if self.nvml is not None else "nvidia-smi -lms 50"}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def bind_to_gpu_numa(local: int) -> bool:
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE the pinned host
    buffers are allocated, so they land on the GPU's NUMA node: with 8 ranks on a two-socket host
    the e2e leg is otherwise limited by cross-socket traffic to wherever the pages happened to land."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local)))
        return True
    except Exception:
        return False


# ----------------------------------------------------------------------------- CPU legs
def cpu_flat_sample(rows: int, reps: int):
    """oracle C port, flat loops + OpenMP on all host cores, on `rows` of the workload."""
    from oracle import c_oracle as CO
    CO.use_all_cores()
    a, b, c = make_inputs(0, rows)
    out, tmp = np.empty_like(a), np.empty_like(a)
    CO.flat_mul_rowvec_add_f32(a, b, c, out, tmp)        # warm (page faults)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        CO.flat_mul_rowvec_add_f32(a, b, c, out, tmp)
        ts.append(time.perf_counter() - t0)
    bytes_sample = BYTES_STEP * rows / ROWS
    return bytes_sample / statistics.median(ts) / 1e9, CO.num_threads(), ts


def cpu_ref_sample(rows: int):
    """oracle C port keeping the reference's structure (1 thread, per-element coordinate
    iterator + dot-product indexing, tile + two operators)."""
    from oracle import c_oracle as CO
    a, b, c = make_inputs(0, rows)
    t0 = time.perf_counter()
    CO.ref_mul_rowvec_add_f32(a, b, c)
    dt = time.perf_counter() - t0
    return (BYTES_STEP * rows / ROWS) / dt / 1e9, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  ph-core is Crystal
    (no compiler in this image), so this is the oracle's C port that keeps the reference's
    STRUCTURE (kind "port"): `b.tile` + two `map_with` operators, each a per-element lexicographic
    coordinate iterator with a coordinate->index dot product per operand and a materialised
    temporary (src/multi_indexable.cr:818-827, 931-952, 1080-1102).  The reference has no threads,
    fibers or SIMD dispatch anywhere, so one host thread is every thread it
    can use.  The flat OpenMP loop on all cores -- a CPU implementation the reference does not
    have -- is timed too and reported beside it, labelled, never as the headline."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # One step = the WHOLE workload (8192 rows, ~1 s on one core) whenever the run then ends within a few
    # minutes (K <= 32; the driver passes K = 20); beyond that a bounded sample of rows per step, stated in
    # config.sample, and ms_per_step is the time of what was actually run -- never a scaled figure.
    rows = ROWS if args.steps <= 32 else (1024 if args.steps <= 200 else (256 if args.steps <= 2000 else 64))
    from oracle import c_oracle as CO
    a, b, c = make_inputs(0, rows)
    for _ in range(max(1, min(args.warmup, 3))):
        CO.ref_mul_rowvec_add_f32(a, b, c)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        CO.ref_mul_rowvec_add_f32(a, b, c)
    dt = time.perf_counter() - t0
    gbs = (BYTES_STEP * rows / ROWS) * args.steps / dt / 1e9
    flat_gbs, flat_cores, _ = cpu_flat_sample(2048, 5)
    sample = ("the full workload per step" if rows == ROWS else
              f"{rows} of {ROWS} rows per step (value = GB/s of the rows run; ms_per_step = time of those rows)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "shape_per_gpu": [rows, COLS], "bytes_per_step_per_gpu": int(BYTES_STEP * rows / ROWS),
                   "sample": sample, "seed": SEED},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": f"{rows}x{COLS} f32 rows per step; C port of the reference's structure "
                                   "(tile + two per-element map_with passes), single thread like the reference",
                         "flat_openmp_all_cores": {"value": round(flat_gbs, 3), "unit": "GB/s", "cores": flat_cores,
                                                   "note": "flat loops + OpenMP: NOT the reference's implementation "
                                                           "(it is single-threaded, per-element iterators); most generous CPU bound"}},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import ph_core_b200 as ph
    from ph_core_b200 import DeviceNArray as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the device path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa_bound = bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ph.init(local)
    lib = ph.load()
    stream = torch.cuda.ExternalStream(lib.ph_stream(), device=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs resident in HBM (this rank's 8192x8192 shard of the [world*8192, 8192] array)
    a_h, b_h, c_h = make_inputs(rank)
    a, b, c = D.from_host(a_h), D.from_host(b_h), D.from_host(c_h)
    t = D(a.shape, np.float32)
    out = D(a.shape, np.float32)
    da, dt_, dc, do = a.desc(), t.desc(), c.desc(), out.desc()
    db = b.bcast_desc(a.shape)
    F32, MUL, ADD = ph.K["PH_F32"], ph.K["PH_MUL"], ph.K["PH_ADD"]

    def step():
        ph.check(lib.ph_ewise_binary(MUL, F32, a.ptr, C.byref(da), b.ptr, C.byref(db), t.ptr, C.byref(dt_)))
        ph.check(lib.ph_ewise_binary(ADD, F32, t.ptr, C.byref(dt_), c.ptr, C.byref(dc), out.ptr, C.byref(do)))

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = lib.ph_launch_count()
    # Two events bracket the K steps (the headline).  Every SAMPLE-th step additionally carries
    # three events so the two kernels' own durations are measured inside the same timed region
    # (an event after every launch would put a gap behind each of the 2K kernels).
    SAMPLE = 8
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampled = [i for i in range(args.steps) if i % SAMPLE == SAMPLE // 2 or args.steps < SAMPLE]
    ev = {i: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for i in sampled}
    barrier()
    with torch.cuda.stream(stream):
        ev_start.record(stream)
        for i in range(args.steps):
            e = ev.get(i)
            if e:
                e[0].record(stream)
            ph.check(lib.ph_ewise_binary(MUL, F32, a.ptr, C.byref(da), b.ptr, C.byref(db), t.ptr, C.byref(dt_)))
            if e:
                e[1].record(stream)
            ph.check(lib.ph_ewise_binary(ADD, F32, t.ptr, C.byref(dt_), c.ptr, C.byref(dc), out.ptr, C.byref(do)))
            if e:
                e[2].record(stream)
        ev_end.record(stream)
    barrier()
    launches = lib.ph_launch_count() - launches0
    total_ms = ev_start.elapsed_time(ev_end)
    mul_ms = [ev[i][0].elapsed_time(ev[i][1]) for i in sampled]
    add_ms = [ev[i][1].elapsed_time(ev[i][2]) for i in sampled]
    if dist is not None:
        tt = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = BYTES_STEP * world / (ms_per_step * 1e-3) / 1e9

    # ---- fused single-pass variant, reported beside the headline
    for _ in range(3):
        ph.check(lib.ph_ewise_mul_add(F32, a.ptr, C.byref(da), b.ptr, C.byref(db), c.ptr, C.byref(dc), out.ptr, C.byref(do)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            ph.check(lib.ph_ewise_mul_add(F32, a.ptr, C.byref(da), b.ptr, C.byref(db), c.ptr, C.byref(dc), out.ptr, C.byref(do)))
        e1.record(stream)
    barrier()
    fused_ms = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the public ARRAY API with HOST buffers: every step uploads a, b, c from pinned
    # host memory, runs the two operators and downloads the result (all inside the timed region).
    # RowPipeline cuts the rows into chunks on the library's own streams so upload, kernels and download
    # overlap; the expression is written with the array API (no hand-built descriptors, no torch streams).
    a_pin, b_pin, c_pin = ph.pinned_from(a_h), ph.pinned_from(b_h), ph.pinned_from(c_h)
    out_pin = ph.pinned_empty(a_h.shape, np.float32)
    # 13 chunks: 512, 512, 1024 rows (the first download starts after 512 rows), 2 x 2048, then 1024, 512, ... 16, 16 (the
    # tail nothing overlaps is a 16-row chunk).  Measured inside this harness (profiles/r02_pipeline_schedules.jsonl):
    # 16 equal chunks 120.5 GB/s, "4,7,0" 124.2, "4,7,2" 124.8, "4,7,5" 124.8 (and once 100.7: every chunk costs ~0.35 ms of
    # Python to enqueue, so the Python mirror keeps the chunk count moderate; the compiled host layer uses the deeper ramp:
    # 126.2).  PH_E2E_SCHEDULE=chunks,taper,ramp overrides.
    sched = [int(v) for v in os.environ.get("PH_E2E_SCHEDULE", "4,7,2").split(",")]
    pipe = ph.pipeline.RowPipeline(chunks=sched[0], taper=sched[1], ramp=sched[2])
    expr = lambda x, z, y: x.broadcast_op("*", y) + z          # (a * b) + c, b the [1, COLS] row vector

    def e2e_step():
        pipe.map_rows(expr, rows=[a_pin, c_pin], out=out_pin, shared=[b_pin], wait=False)

    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    ph.sync()
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        e1.record(stream)
    barrier()
    ph.sync()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    e2e_out = out_pin.copy()

    # the same through the plainest calls a user can make: pageable numpy in, numpy out, one stream
    def naive_step():
        return (D.from_host(a_h).broadcast_op("*", D.from_host(b_h)) + D.from_host(c_h)).to_host()
    naive_step()
    barrier()
    t0 = time.perf_counter()
    naive_out = naive_step()
    naive_ms = (time.perf_counter() - t0) * 1e3
    naive_ok = bool(naive_out.tobytes() == e2e_out.tobytes())

    # what the host link alone allows, measured with every rank active at once: the step's uploads on one
    # stream and its download on another, no kernels -- the ceiling the e2e figure is judged against
    s_up, s_dn = ph.Stream(), ph.Stream()
    def link_step():
        s_up.wait(None); s_dn.wait(None)
        with s_up:
            ph.check(lib.ph_h2d(a.ptr, a_pin.ctypes.data, a_pin.nbytes))
            ph.check(lib.ph_h2d(c.ptr, c_pin.ctypes.data, c_pin.nbytes))
            ph.check(lib.ph_h2d(b.ptr, b_pin.ctypes.data, b_pin.nbytes))
        with s_dn:
            ph.check(lib.ph_d2h_async(out_pin.ctypes.data, out.ptr, out_pin.nbytes))
        ph.narray.main_stream_wait(s_up); ph.narray.main_stream_wait(s_dn)
    link_step()
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(3):
            link_step()
        e1.record(stream)
    barrier()
    link_ms = e0.elapsed_time(e1) / 3
    if dist is not None:
        tt = torch.tensor([e2e_ms, link_ms, naive_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, link_ms, naive_ms = [float(v) for v in tt.tolist()]
    e2e_value = BYTES_STEP * world / (e2e_ms * 1e-3) / 1e9
    link_ceiling = BYTES_STEP * world / (link_ms * 1e-3) / 1e9
    clocks = sampler.stop()          # sampled across the headline, fused and e2e timed regions

    # ---- parity spot check of the timed result (oracle = checker only)
    from oracle import c_oracle as CO
    chk_rows = 64
    want = CO.flat_mul_rowvec_add_f32(a_h[:chk_rows].copy(), b_h, c_h[:chk_rows].copy())
    parity_ok = bool(e2e_out[:chk_rows].tobytes() == want.tobytes())

    os.sched_setaffinity(0, all_cpus)               # the CPU baselines may use every host core
    extras = None
    if not args.no_extras:
        try:
            extras = run_extras(ph, lib, dist, world, rank, torch, stream)
        except Exception as e:                       # extras never invalidate the headline
            extras = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = measured_peak()
        dom_ms = statistics.mean(add_ms)
        achieved = BYTES_ADD / (dom_ms * 1e-3) / 1e9
        if world == 1:
            # reference-structured port, single thread (the reference has no threads), ~10 s of CPU work
            ref_runs = [cpu_ref_sample(2048) for _ in range(5)]
            cpu_ref = statistics.median(r[0] for r in ref_runs)
            cpu_flat, cores, _ = cpu_flat_sample(2048, 5)
            cpu_baseline = {"value": round(cpu_ref, 4), "unit": "GB/s", "cores": 1, "kind": "port",
                            "sample": "2048 of 8192 rows x 5 (median, %.1f s in total); oracle C port keeping the "
                                      "reference's structure: tile + two per-element map_with passes, one thread"
                                      % sum(r[1] for r in ref_runs),
                            "flat_openmp_all_cores": {"value": round(cpu_flat, 3), "unit": "GB/s", "cores": cores,
                                                      "sample": "2048 of 8192 rows, flat loops + OpenMP, median of 5",
                                                      "note": "not the reference's implementation; most generous CPU bound"}}
        else:
            cpu_baseline = None                     # timed on rank 0 at N=1 only (contract)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "shape_per_gpu": [ROWS, COLS], "bytes_per_step_per_gpu": BYTES_STEP,
                       "parallelism": f"axis0-shard x{world}, no collective",
                       "l2": "no flush: each 256 MiB operand exceeds the 126 MB L2; consecutive launches traverse in opposite "
                             "directions, so `t + c` finds the last-written part of the temporary t in L2 "
                             "(3.7 % of that kernel's time; PH_FLAT_NO_ALTERNATE=1 gives 6653 GB/s instead of 6798); inputs are read "
                             "with L2 evict_first and results stored evict_last (PH_FLAT_L2_HINT=0: 6786 instead of 6843)",
                       "seed": SEED},
            "pct_of_peak": {"of_measured_copy": round(value / world / peak, 4), "of_nominal_8000": round(value / world / 8000, 4)},
            "roofline": {"bound": "hbm", "kernel": "map_flat_kernel<BinaryOp<float,ADD>,8,2> (out = t + c, all operands contiguous)",
                         "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": NCU_TRAFFIC_ADD, "peak_source": peak_src,
                         "traffic_source": "profiles/r02_bench_dram_warm.csv (ncu, one pass, --cache-control none: dram__bytes_read.sum 477.6 MB "
                                           "+ dram__bytes_write.sum 269.8 MB per launch; cold-cache --set full capture: 536.9 MB read)",
                         "algorithmic_bytes_per_launch": BYTES_ADD, "avg_launch_ms": round(dom_ms, 5),
                         "launches_timed": len(add_ms),
                         "other_kernels": {"map_flat_kernel<BinaryOp<float,MUL>,8,2> (t = a * b, b periodic: the [1,8192] row vector)": {
                             "algorithmic_bytes_per_launch": BYTES_MUL, "avg_launch_ms": round(statistics.mean(mul_ms), 5),
                             "achieved": round(BYTES_MUL / (statistics.mean(mul_ms) * 1e-3) / 1e9, 2)}}},
            "fused_single_pass": {"ms_per_step": round(fused_ms, 5), "algorithmic_bytes": BYTES_ADD + COLS * 4,
                                  "gbs": round((BYTES_ADD + COLS * 4) / (fused_ms * 1e-3) / 1e9, 2)},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": round(e2e_value, 3), "unit": "GB/s", "ms_per_step": round(e2e_ms, 4),
                    "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes + c_h.nbytes), "d2h_bytes_per_step": int(a_h.nbytes),
                    "steps": e2e_steps,
                    "how": "array API: pipeline.RowPipeline.map_rows (from_host_async of pinned a, b, c -> a.broadcast_op('*', b) + c "
                           "-> to_host_async), 13 row chunks (512, 512, 1024, 2 x 2048, then halving down to 16: downloads start early and the "
                           "un-overlapped tail is one small chunk) on upload / "
                           "compute / download streams of the library so H2D, kernels and D2H overlap",
                    "naive": {"value": round(BYTES_STEP * world / (naive_ms * 1e-3) / 1e9, 3), "ms_per_step": round(naive_ms, 3),
                              "how": "from_host(pageable numpy) x3 -> operators -> to_host, one stream, wall clock", "same_result": naive_ok},
                    "host_link_ceiling": {"value": round(link_ceiling, 3), "unit": "GB/s", "ms_per_step": round(link_ms, 4),
                                          "how": "the step's 537 MB up and 268 MB down on two streams at once, no kernels, all ranks concurrently"},
                    "frac_of_host_link_ceiling": round(e2e_value / link_ceiling, 4),
                    "host_buffers_numa_local": numa_bound},
            "gpu_launches": int(launches), "clocks": clocks, "parity_spot_check": parity_ok,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _gather_objects(dist, world, obj):
    if dist is None:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def run_extras(ph, lib, dist, world, rank, torch, stream):
    """Secondary BASELINE.json configs reported beside the headline (same JSON line, key `extras`), each
    SELF-VERIFYING so that the driver's N = 1/2/4/8 runs check the multi-GPU paths:
      * heat3d_2048_f32: 2048^3 f32 stencil, slab-decomposed over the N ranks (STRONG scaling).  Seeded
        non-constant field (analytic bump + per-plane Philox noise, identical for every N), 10 warm-up +
        100 timed steps; `subcube_vs_oracle` = a 48^3 crop straddling the slab boundary at plane 1024
        replayed through the oracle for the first 2 steps, bit for bit; `field_hash` = position-weighted
        64-bit checksum of the final field -- equal at N = 1/2/4/8 iff the fields are bit-identical.
      * reduce_sum_1e9_f32 / reduce_argmax_1e9_f32: full reductions of [1000,1000,1000] f32 sharded along
        axis 0; integer-valued data (every partial sum exact => the result must EQUAL the exact total) with
        a planted maximum and a planted tie in another shard (argmax must return the lower global index).
      * sharded_permute_16384_f64 (N > 1): transpose across shards, compared with the closed form."""
    import math
    from ph_core_b200 import DeviceNArray as D, sharding as S, heat
    from oracle import ph_oracle as O
    out = {}
    S.comm_init(dist)
    p2p = S.p2p_ready()
    dev = torch.device("cuda", torch.cuda.current_device())
    ms = C.c_float()

    def max_over_ranks(v):
        t = torch.tensor([v], device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- heat 3-D 2048^3
    G, GH, COEFF, WARM, STEPS = 2048, 2, 0.1, 10, 100
    lay = S.slab_layout(G, world, rank, ghost=GH)
    pbytes = G * G * 4
    if world == 1:
        a, b = D([G, G, G], np.float32), D([G, G, G], np.float32)
        first_local, own0, own1 = 0, 0, G                       # local plane of global plane 0; owned global planes
        run = lambda x, y, n: heat.simulate_into(x, y, COEFF, n)
    else:
        shape = [lay["local_planes"], G, G]
        a, b = S.symm_empty(shape, np.float32), S.symm_empty(shape, np.float32)     # peer-mapped when P2P is up
        first_local, own0, own1 = GH - lay["start"], lay["start"], lay["stop"]
        run = lambda x, y, n: S.heat_run_sharded(x, y, COEFF, n, ghost=GH)
    # field: plane z = 100 * g(z) * g(y) g(x) + U[0,1) noise seeded by z alone (so every N builds the same grid)
    with torch.cuda.stream(stream):
        idx = torch.arange(G, device=dev, dtype=torch.float32)
        cen, sig = (G - 1) / 2.0, G / 6.0
        g1 = torch.exp(-((idx - cen) ** 2) / (2 * sig * sig))
        bump_xy = 100.0 * g1[:, None] * g1[None, :]
        gen = torch.Generator(device=dev)
        for z in range(own0, own1):
            gen.manual_seed(SEED * 4096 + z)
            plane = bump_xy * float(np.float32(math.exp(-((z - cen) ** 2) / (2 * sig * sig)))) + \
                torch.rand(G, G, generator=gen, device=dev, dtype=torch.float32)
            ph.check(lib.ph_d2d(a.ptr + (z + first_local) * pbytes, plane.data_ptr(), pbytes))
        del plane
    torch.cuda.synchronize()

    CZ, CY, CX, CN = G // 2 - 24, 700, 1100, 48                # crop straddles plane 1024: a slab boundary at N = 2, 4, 8
    def crop_of(x):
        z0, z1 = max(CZ, own0), min(CZ + CN, own1)
        if z0 >= z1:
            return None
        part = x[ph.rng(z0 + first_local, z1 - 1 + first_local), ph.rng(CY, CY + CN - 1), ph.rng(CX, CX + CN - 1)].to_host()
        return (z0, part)
    def gather_crop(x):
        parts = [p for p in _gather_objects(dist, world, crop_of(x)) if p is not None]
        return np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])], axis=0)

    crop0 = gather_crop(a)
    cur = run(a, b, 2)
    crop2 = gather_crop(cur)
    want = crop0
    for _ in range(2):
        want = O.heat_step_nd(want, np.float32(COEFF))
    # cells at distance >= 2 from the crop's faces depend only on cells inside the crop
    subcube_ok = bool(crop2.shape == (CN, CN, CN) and crop2[2:-2, 2:-2, 2:-2].tobytes() == want[2:-2, 2:-2, 2:-2].tobytes()
                      and not np.array_equal(crop2, crop0))
    other = b if cur is a else a
    ph.check(lib.ph_timer_start())
    cur2 = run(cur, other, WARM - 2)
    ph.check(lib.ph_timer_stop(C.byref(ms)))
    warm_ms = max_over_ranks(ms.value) / (WARM - 2)
    other = cur if cur2 is not cur else other
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    heat_clocks = ClockSampler(physical_gpu_index(torch.cuda.current_device()))
    heat_clocks.start()
    nccl0 = lib.ph_nccl_call_count()
    ph.check(lib.ph_timer_start())
    fin = run(cur2, other, STEPS)
    ph.check(lib.ph_timer_stop(C.byref(ms)))
    heat_nccl = max(_gather_objects(dist, world, int(lib.ph_nccl_call_count() - nccl0)))
    heat_clk = heat_clocks.stop()
    heat_ms = max_over_ranks(ms.value) / STEPS
    h = C.c_uint64(0)
    ph.check(lib.ph_checksum64(fin.ptr + (own0 + first_local) * pbytes, (own1 - own0) * pbytes, own0 * (pbytes // 8), C.byref(h)))
    field_hash = sum(_gather_objects(dist, world, int(h.value))) % (1 << 64)
    out["heat3d_2048_f32"] = {
        "gcell_updates_per_s": round(G ** 3 / (heat_ms * 1e-3) / 1e9, 2), "ms_per_step": round(heat_ms, 4),
        "grid": [G, G, G], "steps": STEPS, "warmup": WARM, "scaling": "strong",
        "field": "100 * gaussian bump + U[0,1) noise (torch Philox seeded per global plane), fixed boundary, C = 0.1",
        "field_hash": f"{field_hash:016x}", "field_hash_steps": WARM + STEPS, "subcube_vs_oracle": subcube_ok,
        "subcube": f"48^3 crop at [{CZ},{CY},{CX}] (straddles plane 1024), first 2 steps replayed by oracle.heat_step_nd, inner 44^3 bit-exact",
        "halo": ("none (1 GPU)" if world == 1 else
                 ("in-kernel P2P: the stencil kernel stores its edge planes into the neighbours' ghost planes (peer-mapped slabs), "
                  "flag words + stream waits, no NCCL" if p2p else "NCCL send/recv of 2-plane halos every 2 steps, overlapped with the interior")),
        "decomposition": f"axis-0 slabs x{world}, two time steps per pass over HBM (temporal blocking, bit-identical)",
        "algorithmic_bytes_per_cell_update": 8,
        "hbm_gbs_per_gpu": round(8 * G ** 3 / world / (heat_ms * 1e-3) / 1e9, 1),
        "clocks": heat_clk,
        "nccl_calls_in_timed_region": heat_nccl,
        "first_steps": {"steps": WARM - 2, "gcell_updates_per_s": round(G ** 3 / (warm_ms * 1e-3) / 1e9, 2),
                        "note": "the warm-up steps right after set-up, timed the same way: the kernel's rate before the "
                                "GPU's power limit settles (the 100 timed steps run under it: see clocks.reasons)"}}
    if world == 1:
        # the reference's CPU path beside it: the slice-arithmetic step through ph-core's operator
        # structure (C port, one thread) on a 160^3 sample, and the flat OpenMP loop nest on 512^3
        from oracle import c_oracle as CO
        rs = np.random.RandomState(1)
        small = (rs.rand(160, 160, 160) * 100).astype(np.float32)
        t0 = time.perf_counter(); CO.ref_heat_step_3d_f32(small, 0.1); dt_ref = time.perf_counter() - t0
        CO.use_all_cores()
        mid = (rs.rand(512, 512, 512) * 100).astype(np.float32)
        tmp = np.empty_like(mid)
        CO.flat_heat_step_3d_f32(mid, 0.1, tmp)
        t0 = time.perf_counter()
        for _ in range(3):
            CO.flat_heat_step_3d_f32(mid, 0.1, tmp)
        dt_flat = (time.perf_counter() - t0) / 3
        out["heat3d_2048_f32"]["cpu_baseline"] = {
            "value": round(small.size / dt_ref / 1e9, 5), "unit": "Gcell-updates/s", "cores": 1, "kind": "port",
            "sample": f"one step of a 160^3 f32 grid through the reference's operator structure ({dt_ref:.2f} s)",
            "flat_openmp_all_cores": {"value": round(mid.size / dt_flat / 1e9, 3), "cores": CO.num_threads(),
                                      "sample": "512^3 f32, flat loop nest + OpenMP, mean of 3"}}
        del small, mid, tmp
    if world > 1:
        a._buf.free(); b._buf.free()
    del a, b, cur, cur2, fin, other

    # ---------------------------------------------------------------- full reductions of 1e9 f32 sharded along axis 0
    R0, INNER = 1000, 1000 * 1000
    n_total = R0 * INNER
    r0, r1 = S.shard_range(R0, world, rank)
    x = D([max(r1 - r0, 0), 1000, 1000], np.float32)
    P1, P2 = 123_456_789, 876_543_210                        # planted maximum and its tie (another shard at N >= 2)
    exact = torch.zeros((), dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        gen = torch.Generator(device=dev)
        for i in range(r0, r1):
            gen.manual_seed(SEED * 8192 + i)
            row = torch.randint(-8, 9, (INNER,), generator=gen, device=dev, dtype=torch.int32)
            for p in (P1, P2):
                if p // INNER == i:
                    row[p % INNER] = 99
            exact += row.sum(dtype=torch.int64)
            rowf = row.to(torch.float32)
            ph.check(lib.ph_d2d(x.ptr + (i - r0) * INNER * 4, rowf.data_ptr(), INNER * 4))
        del row, rowf
    torch.cuda.synchronize()
    if dist is not None:
        dist.all_reduce(exact)
    exact = int(exact.item())
    off = r0 * INNER
    reps = 20
    for name, key in (("sum", "reduce_sum_1e9_f32"), ("argmax", "reduce_argmax_1e9_f32")):
        got = S.reduce_full_sharded(x, name, off)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        nccl0 = lib.ph_nccl_call_count()
        ph.check(lib.ph_timer_start())
        for _ in range(reps):
            got = S.reduce_full_sharded(x, name, off)
        ph.check(lib.ph_timer_stop(C.byref(ms)))
        red_nccl = max(_gather_objects(dist, world, int(lib.ph_nccl_call_count() - nccl0)))
        red_ms = max_over_ranks(ms.value) / reps
        ok = (float(got) == float(exact)) if name == "sum" else (float(got[0]) == 99.0 and int(got[1]) == P1)
        ok = all(_gather_objects(dist, world, bool(ok)))
        out[key] = {"gbs": round(4 * n_total / (red_ms * 1e-3) / 1e9, 1), "ms": round(red_ms, 4), "result_ok": ok,
                    "result": (float(got) if name == "sum" else [float(got[0]), int(got[1])]),
                    "expected": (exact if name == "sum" else [99.0, P1]),
                    "check": ("integers in {-8..8} (+ two planted 99s): every partial sum is exact in f32, the result must equal the exact total"
                              if name == "sum" else "planted maximum 99 at two global indices in different shards: argmax must return the lower one"),
                    "collective": ("none (1 GPU)" if world == 1 else
                                   ("in-kernel one-shot combine over peer-mapped slots (no NCCL, one launch, one sync)" if p2p
                                    else "ncclAllGather of one 64-byte record per GPU + a combine launch")),
                    "nccl_calls_in_timed_region": red_nccl, "reps": reps}
    if world == 1:
        from oracle import c_oracle as CO
        xs = np.random.RandomState(2).rand(200_000_000).astype(np.float32)
        t0 = time.perf_counter(); CO.ref_sum_f32(xs); dt_ref = time.perf_counter() - t0
        CO.use_all_cores()
        CO.flat_sum_f32(xs)
        t0 = time.perf_counter(); CO.flat_sum_f32(xs); dt_flat = time.perf_counter() - t0
        out["reduce_sum_1e9_f32"]["cpu_baseline"] = {
            "value": round(xs.nbytes / dt_ref / 1e9, 3), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"2e8 of 1e9 elements, Enumerable#sum as a sequential f32 left fold ({dt_ref:.2f} s)",
            "flat_openmp_all_cores": {"value": round(xs.nbytes / dt_flat / 1e9, 2), "cores": CO.num_threads(),
                                      "sample": "2e8 elements, OpenMP reduction in f64"}}
        del xs
    del x

    # ---------------------------------------------------------------- transpose across shards (f-3), N > 1
    if world > 1:
        n = 16384
        q0, q1 = S.shard_range(n, world, rank)
        rows_v = D.from_host((np.arange(q0, q1, dtype=np.float64) * n).reshape(q1 - q0, 1))
        cols_v = D.from_host(np.arange(n, dtype=np.float64).reshape(1, n))
        src = S.ShardedNArray([n, n], rows_v.broadcast_op("+", cols_v))          # value = flat index
        t = src.permute()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        preps = 3
        nccl0 = lib.ph_nccl_call_count()
        ph.check(lib.ph_timer_start())
        p2p_perm = S.p2p_ready() and not os.environ.get("PH_PERMUTE_NCCL")
        for _ in range(preps):
            t = src.permute(out=t) if p2p_perm else src.permute()
        ph.check(lib.ph_timer_stop(C.byref(ms)))
        perm_nccl = max(_gather_objects(dist, world, int(lib.ph_nccl_call_count() - nccl0)))
        perm_ms = max_over_ranks(ms.value) / preps
        # transposed[i, j] = j * n + i on my rows i in [q0, q1)
        expect = D.from_host(np.arange(q0, q1, dtype=np.float64).reshape(q1 - q0, 1)).broadcast_op(
            "+", D.from_host((np.arange(n, dtype=np.float64) * n).reshape(1, n)))
        ok = all(_gather_objects(dist, world, bool(t.local.equals(expect))))
        hsum = sum(_gather_objects(dist, world, t.local.checksum64(q0 * n))) % (1 << 64)
        out["sharded_permute_16384_f64"] = {"ms": round(perm_ms, 4), "gbs_aggregate": round(2 * n * n * 8 / (perm_ms * 1e-3) / 1e9, 1),
                                            "result_ok": ok, "checksum": f"{hsum:016x}", "nccl_calls_in_timed_region": perm_nccl,
                                            "how": ("ShardedNArray.permute: ONE pass of peer stores -- the transpose kernel writes every block "
                                                    "straight into its owner's shard over NVLink (ph_alltoall_strided), result buffer reused"
                                                    if p2p_perm else
                                                    "ShardedNArray.permute: per-peer permuting gathers, ncclSend/ncclRecv all-to-all, scatters")}
        # ------------------------------------------------------------ N-GPU parity, in the driver's own multi-GPU run
        # (the GPU test box has ONE GPU, so tests/test_gpu_multi.py is skipped there): every agreement check of
        # tests/mgpu_check.py -- slabbed stencil vs oracle and vs the 1-GPU run, in-kernel halos, sharded reductions
        # of every dtype, ShardedNArray vs the undivided array -- outside every timed region.  A disagreement is
        # reported, never swallowed: "ok": false and the assertion text.
        if not os.environ.get("PH_BENCH_NO_PARITY"):
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import mgpu_check
            t0 = time.perf_counter()
            par = mgpu_check.run_checks(world, rank, soft=True)       # soft: every rank walks every collective
            oks = _gather_objects(dist, world, bool(par["ok"]))
            par["ok"] = all(oks)
            par["ranks_ok"] = oks
            par["seconds"] = round(time.perf_counter() - t0, 1)
            out["multi_gpu_parity"] = par
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the heat / reduction extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
