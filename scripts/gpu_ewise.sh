#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ewise.py tests/test_gpu_index.py -m gpu -q --timeout 300 -x 2>&1 | tail -2
timeout 300 python bench.py --no-extras --steps 1000 > gpurun_out/bench_alt.json 2>gpurun_out/bench_alt.err
PH_FLAT_NO_ALTERNATE=1 timeout 300 python bench.py --no-extras --steps 1000 > gpurun_out/bench_noalt.json 2>>gpurun_out/bench_alt.err
python - <<PY
import json
for f in ("bench_alt", "bench_noalt"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["other_kernels"], d["fused_single_pass"], d["e2e"]["value"], d["clocks"])
PY
timeout 300 python benchmarks/bench_kernels.py --only "ewise" 2>&1 | cut -c1-200
