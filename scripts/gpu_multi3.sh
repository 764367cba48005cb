#!/bin/bash
# 2-GPU job: N-GPU agreement test + sharded NArray operations (config 3 reductions with collectives, f-3 permute)
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 benchmarks/bench_sharded_ops.py --reps 3 > gpurun_out/sharded_ops_n$N.jsonl 2> gpurun_out/sharded_ops.err
grep "^{" gpurun_out/sharded_ops_n$N.jsonl | cut -c1-250; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/sharded_ops.err | tail -5 | cut -c1-300
