#!/bin/bash
# 2-GPU job: stencil parity tests + sustained rates of the shuffle form, C++ sharded spec on 2 ranks (P2P and NCCL)
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_heat.py tests/test_gpu_fullsize.py tests/test_gpu_example.py -x -q --timeout 600 2>&1 | tail -4
: > gpurun_out/heat_sweep2.jsonl
run() { env "$@" CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py >> gpurun_out/heat_sweep2.jsonl 2>> gpurun_out/heat_sweep.err; }
run PH_X=0; run PH_HEAT_TB_CFG=1; run PH_HEAT_TB_CFG=2; run PH_HEAT_TB_CFG=5; run PH_HEAT_TB_MARCH=256
env CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py --steps 10 --warm 2 >> gpurun_out/heat_sweep2.jsonl
env CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py --shape 1024,1024,1024 >> gpurun_out/heat_sweep2.jsonl
cat gpurun_out/heat_sweep2.jsonl
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PH_NO_P2P=1; else unset PH_NO_P2P; fi
  rm -f /tmp/ph_id_$mode
  for r in 0 1; do RANK=$r WORLD_SIZE=2 PH_ID_FILE=/tmp/ph_id_$mode timeout 300 tests/cpp/sharded_spec > gpurun_out/sharded_spec_${mode}_r$r.log 2>&1 & done
  wait
  echo "== sharded_spec $mode"; tail -12 gpurun_out/sharded_spec_${mode}_r0.log; tail -2 gpurun_out/sharded_spec_${mode}_r1.log
done
unset PH_NO_P2P
timeout 120 tests/cpp/sharded_spec > gpurun_out/sharded_spec_single.log 2>&1; echo "== sharded_spec single rc=$?"; tail -3 gpurun_out/sharded_spec_single.log
