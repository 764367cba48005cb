#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
