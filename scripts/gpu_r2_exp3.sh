#!/bin/bash
mkdir -p gpurun_out
timeout 600 python benchmarks/bench_kernels.py --only "reduce axis=0" 2>> gpurun_out/exp3.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['gbs'], d['frac_measured'], d['kernel'])"
PH_AXIS_STAGED=0 timeout 600 python benchmarks/bench_kernels.py --only "reduce axis=0 sum f32 [16384" 2>> gpurun_out/exp3.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('unstaged', d['gbs'], d['frac_measured'], d['kernel'])"
tail -3 gpurun_out/exp3.err
