#!/bin/bash
# 2-GPU job: heat variant sweep on GPU 0, C++ sharded spec on 2 ranks (P2P and NCCL), mgpu check + bench at N=2
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 bash scripts/gpu_r2_heat_sweep.sh
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PH_NO_P2P=1; else unset PH_NO_P2P; fi
  rm -f /tmp/ph_id_$mode
  for r in 0 1; do RANK=$r WORLD_SIZE=2 PH_ID_FILE=/tmp/ph_id_$mode timeout 300 tests/cpp/sharded_spec > gpurun_out/sharded_spec_${mode}_r$r.log 2>&1 & done
  wait
  echo "== sharded_spec $mode"; tail -12 gpurun_out/sharded_spec_${mode}_r0.log; tail -2 gpurun_out/sharded_spec_${mode}_r1.log
done
unset PH_NO_P2P
timeout 120 tests/cpp/sharded_spec > gpurun_out/sharded_spec_single.log 2>&1; echo "== sharded_spec single rc=$?"; tail -3 gpurun_out/sharded_spec_single.log
bash scripts/gpu_r2_multi.sh 2 p2ponly
python - <<'PY'
import json
for l in open("gpurun_out/bench_n2_p2p.json"):
    if l.startswith("{"):
        d = json.loads(l); print(json.dumps(d["extras"].get("multi_gpu_parity"))[:600])
PY
