#!/bin/bash
mkdir -p gpurun_out
for cfg in 0 1 2; do for m in 48 64 96 128; do
echo "cfg $cfg march $m: $(PH_HEAT_TB_CFG=$cfg PH_HEAT_TB_MARCH=$m timeout 300 python benchmarks/bench_kernels.py --heat-shape 2048,2048,2048 2>&1 | grep -o 'gcell_per_s.*')"
done; done
echo "1024: $(timeout 300 python benchmarks/bench_kernels.py --heat-shape 1024,1024,1024 2>&1 | grep -o 'gcell_per_s.*')"
echo "256 slab: $(timeout 300 python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 2>&1 | grep -o 'gcell_per_s.*')"
