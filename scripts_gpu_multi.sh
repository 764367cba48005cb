#!/bin/bash
# N-GPU job: agreement tests + bench at N GPUs (N = number of visible devices)
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_heat.py tests/test_gpu_multi.py -m gpu -q -x --timeout 600 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; python - <<PY
import json
for l in open("gpurun_out/bench_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["e2e"]["value"], json.dumps(d["extras"]))
PY
