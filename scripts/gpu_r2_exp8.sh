#!/bin/bash
# Round-2 experiment 8 (one GPU): C++ host-layer spec incl. the pipeline block, C++ host bench (e2e row), pipeline GPU test,
# bench.py headline + e2e with the tapered schedule
mkdir -p gpurun_out
timeout 100 ./tests/cpp/device_narray_spec > gpurun_out/cpp_spec.log 2>&1; echo "cpp spec exit $?"; grep -E "FAIL|expectations" gpurun_out/cpp_spec.log | head
timeout 200 ./tests/cpp/device_narray_bench > gpurun_out/cpp_bench.jsonl 2>&1; cut -c1-230 gpurun_out/cpp_bench.jsonl
timeout 300 python -m pytest tests/test_gpu_streams_slices.py tests/test_cpp_host_layer.py -m gpu -q --timeout 250 2>&1 | tail -3
timeout 300 python bench.py --no-extras > gpurun_out/bench_noextras.json 2> gpurun_out/bench_noextras.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_noextras.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["frac_of_host_link_ceiling"], d["e2e"]["host_link_ceiling"]["value"], d["e2e"]["naive"]["same_result"], d["parity_spot_check"])
P
tail -2 gpurun_out/bench_noextras.err
timeout 200 python benchmarks/bench_pipeline.py 2>&1 | tee gpurun_out/pipeline3.jsonl
