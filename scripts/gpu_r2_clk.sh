#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2975$N \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_p2p.json 2> gpurun_out/bench_n${N}_p2p.err
echo "rc=$?"; grep -v "^W0\|^\*\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_p2p.err | tail -5
python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_p2p.json"):
    if l.startswith("{"):
        d = json.loads(l); h = d["extras"]["heat3d_2048_f32"]
        print("value", d["value"], "clocks", d["clocks"])
        print("heat", h["gcell_updates_per_s"], h["clocks"], h["first_steps"])
        print({k: (v.get("gbs") or v.get("ms") or v.get("ok")) for k, v in d["extras"].items() if k != "heat3d_2048_f32"})
PY
