"""Chunked host -> device -> host pipelines written with the array API.

One elementwise expression over HOST-resident operands is bound by the host link, not by any kernel:
uploading the operands, running the operators and downloading the result one after the other leaves the
link idle two thirds of the time.  `map_rows` cuts the leading axis into chunks and runs them through an
upload stream (`DeviceNArray.from_host_async`, pinned source), a compute stream (the caller's expression on
device arrays) and a download stream (`to_host_async`, pinned destination), so the upload of chunk i+1, the
kernels of chunk i and the download of chunk i-1 overlap (PCIe is full duplex).  Everything goes through the public array API and the
stream entry points of include/ph_gpu.h (ph_stream_create / ph_set_stream / ph_stream_wait); nothing here
builds descriptors by hand.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from .narray import DeviceNArray, Stream, ShapeError, main_stream_wait, sync


def row_chunks(n: int, chunks: int, taper: int = 0, ramp: int = 0):
    """[r0, r1) row ranges (ph_row_chunks, include/ph_host.h -- the schedule is C++ host code shared with the
    compiled host layer; this function marshals): `chunks` equal chunks; with taper = t the LAST one is cut again
    into halves t times (per/2, per/4, ..., per/2^t, per/2^t).  What is left when the last upload ends is one
    chunk's kernels and download -- nothing overlaps that tail -- so the final chunks are small while the early ones
    stay large.  With ramp = r the FIRST chunk is the mirror image (per/2^r, per/2^r, ..., per/2): the first download
    can only start after the first chunk."""
    from . import _lib
    from .narray import host_check
    cap = max(1, int(chunks)) + max(0, int(taper)) + max(0, int(ramp)) + 2
    bounds = (C.c_int64 * (2 * cap))()
    count = C.c_int32(0)
    host_check(_lib.load().ph_row_chunks(int(n), int(chunks), int(taper), int(ramp), bounds, cap, C.byref(count)))
    return [(int(bounds[2 * i]), int(bounds[2 * i + 1])) for i in range(count.value)]


class RowPipeline:
    """Three streams by ROLE, reused across calls: one uploads, one computes, one downloads.  Uploads of all
    chunks are queued back to back (the host-to-device engine never waits for a kernel or a download of an
    earlier chunk); chunk k's operators wait for its upload, its download for its operators.  (Spreading the
    uploads of a chunk over TWO streams is slower: 4 + 7 halvings 10.74 -> 11.77 ms, 16 equal chunks 11.18 -> 11.40,
    profiles/r02_pipeline_schedules.jsonl -- concurrent host-to-device copies share the link and the chunk is
    complete later.)"""

    def __init__(self, chunks: int = 16, streams: int = 3, taper: int = 0, uploaders: int = 1, ramp: int = 0):
        self.chunks = int(chunks)
        self.taper = int(taper)
        self.ramp = int(ramp)
        self.up, self.comp, self.down = Stream(), Stream(), Stream()
        # uploaders > 1: the row operands of a chunk go up on different streams, so one copy's set-up runs under
        # another copy's transfer (experiment knob; 1 = every upload on `up`)
        self.ups = [self.up] + [Stream() for _ in range(max(0, int(uploaders) - 1))]
        self.streams = self.ups + [self.comp, self.down]

    def map_rows(self, fn: Callable, rows: Sequence[np.ndarray], out: np.ndarray, shared: Sequence[np.ndarray] = (),
                 wait: bool = True) -> None:
        """out[r0:r1] = fn(*rows[k][r0:r1] as device arrays, *shared as device arrays) for every row chunk.
        `rows` and `out` are pinned host arrays (narray.pinned_empty) with the same leading extent; `shared`
        operands (e.g. a broadcast row vector) are uploaded once.  With wait=True the call returns when `out`
        is complete and raises pending data-dependent errors; wait=False leaves that to the caller's sync()."""
        n = out.shape[0]
        for r in rows:
            if r.shape[0] != n:
                raise ShapeError("map_rows: every row operand needs the leading extent of the output")
        up, comp, down = self.up, self.comp, self.down
        for s in self.streams:
            s.wait(None)                                           # behind whatever the main stream has queued
        with up:
            shared_dev = [DeviceNArray.from_host_async(x) for x in shared]
        keep = []
        try:
            for r0, r1 in row_chunks(n, self.chunks, self.taper, self.ramp):
                ins = []
                for i, r in enumerate(rows):
                    with self.ups[i % len(self.ups)]:
                        ins.append(DeviceNArray.from_host_async(r[r0:r1]))
                for u in self.ups:
                    comp.wait(u)                                       # chunk k's operands (and the shared ones) have landed
                with comp:
                    res = fn(*ins, *shared_dev)
                down.wait(comp)
                with down:
                    res.to_host_async(out[r0:r1])
                keep.append((ins, res))
                del ins, res
        except BaseException:
            # the caller's expression raised half way: chunks already queued still use the temporaries, so nothing
            # is released before the streams have drained
            for s in self.streams:
                s.synchronize()
            raise
        # device temporaries are released on the streams they were allocated on: order each of those behind
        # every consumer before letting go, then join the main stream
        for u in self.ups:
            u.wait(comp); u.wait(down)
        comp.wait(down)
        for s in self.streams:
            main_stream_wait(s)
        del keep, shared_dev
        if wait:
            sync()

    def close(self) -> None:
        for s in self.streams:
            s.close()
        self.streams = []
