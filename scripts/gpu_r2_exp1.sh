#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_fullsize.py -x -q --timeout 600 2>&1 | tail -3
timeout 600 python benchmarks/bench_kernels.py --only "reduce" > gpurun_out/reduce_rows2.jsonl 2>> gpurun_out/exp1.err
python - <<'PY'
import json
for l in open("gpurun_out/reduce_rows2.jsonl"):
    d = json.loads(l); print(f'{d["gbs"]:8.1f} {d["frac_measured"]:.3f}  {d["kernel"]}')
PY
echo "== axis=0 argmax plan sweep"
for kv in "PH_AXIS_U=8" "PH_AXIS_U=16" "PH_AXIS_BLOCK=128" "PH_AXIS_U=8 PH_AXIS_BLOCK=128" "PH_AXIS_E=16" "PH_AXIS_E=16 PH_AXIS_U=8"; do
  echo "-- $kv"; env $kv timeout 300 python benchmarks/bench_kernels.py --only "reduce axis=0 argmax f32" 2>> gpurun_out/exp1.err | python -c "import sys,json; [print(json.loads(l)['gbs']) for l in sys.stdin]"
done
echo "== 1024^3 march sweep"
for m in 32 64 96 128 192 256 512; do PH_HEAT_TB_MARCH=$m timeout 300 python benchmarks/bench_heat_sustained.py --shape 1024,1024,1024 2>> gpurun_out/exp1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['knobs'], d['gcell_per_s'], d['sm_mhz_median'])"; done
echo "== 512^3 / 2048x1024x1024"
for sh in 512,512,512 2048,1024,1024 256,4096,4096; do timeout 300 python benchmarks/bench_heat_sustained.py --shape $sh 2>> gpurun_out/exp1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['shape'], d['gcell_per_s'], d['sm_mhz_median'])"; done
tail -3 gpurun_out/exp1.err
