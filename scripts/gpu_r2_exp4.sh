#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reduce.py -q -x --timeout 600 2>&1 | tail -2
timeout 600 python benchmarks/bench_kernels.py --only "reduce axis=0" 2>> gpurun_out/exp4.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['gbs'], d['frac_measured'], d['kernel'])"
timeout 600 python benchmarks/bench_kernels.py --only "view reduce" 2>> gpurun_out/exp4.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['gbs'], d['frac_measured'], d['kernel'])"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x -k "staged or strided_views" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | head -3
tail -3 gpurun_out/exp4.err
