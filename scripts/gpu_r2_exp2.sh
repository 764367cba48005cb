#!/bin/bash
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_fullsize.py -x -q --timeout 600 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python benchmarks/bench_kernels.py --only "reduce axis=0" > gpurun_out/reduce_rows3.jsonl 2>> gpurun_out/exp2.err
python - <<'PY'
import json
for l in open("gpurun_out/reduce_rows3.jsonl"):
    d = json.loads(l); print(f'{d["gbs"]:8.1f} {d["frac_measured"]:.3f}  {d["kernel"]}')
PY
echo "== axis=0 sum/max with U=8 B=128"
PH_AXIS_U=8 PH_AXIS_BLOCK=128 CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_kernels.py --only "reduce axis=0" 2>> gpurun_out/exp2.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['gbs'], d['kernel'])"
bash scripts/gpu_r2_multi.sh 4 p2ponly
