#!/bin/bash
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for grid in 512,2048,2048 2048,2048,2048; do
  timeout 300 $TR --master-port 29801 benchmarks/bench_heat_sharded.py --grid $grid --steps 10 2>&1 | grep "^{"
  PH_HEAT_NO_OVERLAP=1 timeout 300 $TR --master-port 29802 benchmarks/bench_heat_sharded.py --grid $grid --steps 10 2>&1 | grep "^{"
done
timeout 300 $TR --master-port 29803 benchmarks/bench_heat_sharded.py --grid 512,2048,2048 --steps 10 --ghost 1 2>&1 | grep "^{"
timeout 200 python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 2>&1 | cut -c1-200
timeout 200 python benchmarks/bench_kernels.py --heat-shape 1024,2048,2048 2>&1 | cut -c1-200
