#!/bin/bash
# sustained (100-step) 2048^3 stencil rate for the tile / ring / arithmetic variants of the two-step kernel
mkdir -p gpurun_out; : > gpurun_out/heat_sweep.jsonl
run() { env "$@" timeout 300 python benchmarks/bench_heat_sustained.py >> gpurun_out/heat_sweep.jsonl 2>> gpurun_out/heat_sweep.err; }
run PH_X=0
for cfg in 1 2 3 4 5 6 7; do run PH_HEAT_TB_CFG=$cfg; done
for m in 64 256; do run PH_HEAT_TB_MARCH=$m; done
run PH_HEAT_TB_CFG=5 PH_HEAT_TB_MARCH=64
run PH_HEAT_NO_FUSE2=1
cat gpurun_out/heat_sweep.jsonl; tail -5 gpurun_out/heat_sweep.err
