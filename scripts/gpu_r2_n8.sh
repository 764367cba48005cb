#!/bin/bash
# 8-GPU job: C++ sharded spec on 8 ranks, bench at N=8 with the P2P and the NCCL forms (the parity leg runs inside bench.py)
mkdir -p gpurun_out
N=${1:-8}
rm -f /tmp/ph_id_8
for r in $(seq 0 $((N-1))); do RANK=$r WORLD_SIZE=$N PH_ID_FILE=/tmp/ph_id_8 timeout 300 tests/cpp/sharded_spec > gpurun_out/sharded_spec_n${N}_r$r.log 2>&1 & done
wait
echo "== sharded_spec N=$N"; tail -9 gpurun_out/sharded_spec_n${N}_r0.log; tail -1 gpurun_out/sharded_spec_n${N}_r$((N-1)).log
bash scripts/gpu_r2_multi.sh $N ab
python - <<PY
import json
for mode in ("p2p", "nccl"):
    for l in open(f"gpurun_out/bench_n${N}_{mode}.json"):
        if l.startswith("{"):
            d = json.loads(l); p = d["extras"].get("multi_gpu_parity", {})
            print(mode, "parity", {k: p.get(k) for k in ("ok", "failed", "checks", "seconds", "p2p")})
            print(mode, "permute", d["extras"].get("sharded_permute_16384_f64"))
PY
