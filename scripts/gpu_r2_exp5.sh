#!/bin/bash
# Round-2 experiment 5 (one GPU): the two-columns-per-lane staged strip kernel (4-byte elements) -- parity, then A/B.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x --timeout 250 2>&1 | tail -4
for cpt in 1 2; do
  echo "== PH_AXIS_STAGED_CPT=$cpt"
  PH_AXIS_STAGED_CPT=$cpt timeout 120 python benchmarks/bench_kernels.py --only "reduce axis=0" 2>gpurun_out/exp5.err | grep "f32 \[" | cut -c1-260 | tee -a gpurun_out/exp5_cpt$cpt.jsonl
done
for m in 64 128 256; do
  echo "== PH_HEAT_TB_MARCH=$m 1024^3"
  PH_HEAT_TB_MARCH=$m timeout 120 python benchmarks/bench_kernels.py --heat-shape 1024,1024,1024 2>>gpurun_out/exp5.err | cut -c1-260 | tee -a gpurun_out/exp5_heat1024.jsonl
done
tail -3 gpurun_out/exp5.err
