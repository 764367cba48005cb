#!/bin/bash
# 1-GPU job: parity tests, then the bench line (default K/W of the driver)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_n1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "e2e", json.dumps(d["e2e"]))
        print("extras", json.dumps(d["extras"], indent=1)[:6000])
PY
