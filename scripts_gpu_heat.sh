#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heat.py -m gpu -q --timeout 300 -x -k "2d" 2>&1 | tail -3
for cfg in 0 1 2 3 4; do
echo "2D cfg $cfg"
PH_HEAT2D_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_heat.py -m gpu -q --timeout 300 -x -k "2d_two_step" 2>&1 | tail -1
PH_HEAT2D_CFG=$cfg timeout 300 python benchmarks/bench_kernels.py --only "heat 2-D" 2>&1 | cut -c1-220
done
