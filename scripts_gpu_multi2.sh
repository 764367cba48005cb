#!/bin/bash
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 500 2>&1 | tail -3
for grid in 512,2048,2048 2048,2048,2048; do
  timeout 300 $TR --master-port 29801 benchmarks/bench_heat_sharded.py --grid $grid --steps 10 2>&1 | grep "^{"
done
PH_HEAT_NO_OVERLAP=1 timeout 300 $TR --master-port 29802 benchmarks/bench_heat_sharded.py --grid 512,2048,2048 --steps 10 2>&1 | grep "^{"
timeout 300 $TR --master-port 29803 benchmarks/bench_heat_sharded.py --grid 512,2048,2048 --steps 10 --ghost 1 2>&1 | grep "^{"
