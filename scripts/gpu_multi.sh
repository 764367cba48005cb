#!/bin/bash
# N-GPU job: agreement tests + bench at N GPUs (N = number of visible devices; on 8 also N = 4)
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 2>&1 | tail -4
for n in 4 $N; do
  if [ $n -le $N ] && { [ $n -eq $N ] || [ $N -eq 8 ]; }; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    tail -2 gpurun_out/bench_n$n.err
    python - <<PY
import json
for l in open("gpurun_out/bench_n$n.json"):
    if l.startswith("{"):
        d = json.loads(l); print($n, d["value"], d["e2e"]["value"], json.dumps(d["extras"]))
PY
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29799 benchmarks/bench_heat_sharded.py --grid 2048,2048,2048 --steps 20 --reps 3 2>&1 | grep "^{"
PH_HEAT_NO_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29798 benchmarks/bench_heat_sharded.py --grid 2048,2048,2048 --steps 20 --reps 3 2>&1 | grep "^{"
