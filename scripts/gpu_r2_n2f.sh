#!/bin/bash
# 2-GPU job: bench line with the NCCL data-path call counter per leg, peer-memory forms (default) and NCCL forms
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_p2p.json 2> gpurun_out/bench_n2_p2p.err
PH_NO_P2P=1 PH_BENCH_NO_PARITY=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
python - <<'P'
import json
for f in ("p2p","nccl"):
    try:
        d=json.loads(open(f"gpurun_out/bench_n2_{f}.json").read().strip().splitlines()[-1])
        x=d["extras"]
        print(f, d["value"], {k:(v.get("nccl_calls_in_timed_region"), v.get("ms") or v.get("ms_per_step"), v.get("result_ok", v.get("field_hash"))) for k,v in x.items() if isinstance(v,dict) and "nccl_calls_in_timed_region" in v}, x.get("multi_gpu_parity",{}).get("ok"))
    except Exception as e:
        print(f, "ERR", e)
P
tail -3 gpurun_out/bench_n2_p2p.err
