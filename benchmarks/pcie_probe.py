#!/usr/bin/env python
"""Host link probe behind DESIGN.md's e2e discussion: pinned H2D alone, D2H alone, and both at once
(256 MiB buffers, CUDA events), i.e. the ceiling of bench.py's e2e leg on this box."""
import json
import torch

n = 256 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


for s in (s1, s2):
    s.wait_stream(torch.cuda.current_stream())
t_h, t_d, t_b = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"h2d_gbs": round(n / t_h / 1e6, 1), "d2h_gbs": round(n / t_d / 1e6, 1),
                  "both_each_gbs": round(n / t_b / 1e6, 1), "both_total_gbs": round(2 * n / t_b / 1e6, 1)}))
