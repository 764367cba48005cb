#!/bin/bash
# 8-GPU job: agreement checks (both transports) after the ordered all-reduce, C++ sharded spec, sharded-ops table
mkdir -p gpurun_out
N=${1:-8}
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PH_NO_P2P=1; else unset PH_NO_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N \
    tests/mgpu_check.py > gpurun_out/mgpu_$mode.out 2> gpurun_out/mgpu_$mode.err
  echo "== mgpu $mode rc=$?"; tail -1 gpurun_out/mgpu_$mode.out; grep -E "Error|assert" gpurun_out/mgpu_$mode.err | head -5
done
unset PH_NO_P2P
rm -f /tmp/ph_id_8
for r in $(seq 0 $((N-1))); do RANK=$r WORLD_SIZE=$N PH_ID_FILE=/tmp/ph_id_8 timeout 300 tests/cpp/sharded_spec > gpurun_out/sharded_spec_n${N}_r$r.log 2>&1 & done
wait
echo "== sharded_spec N=$N"; grep -E "FAIL|expectations" gpurun_out/sharded_spec_n${N}_r0.log gpurun_out/sharded_spec_n${N}_r$((N-1)).log
for tag in p2p nccl_allreduce; do
  if [ $tag = nccl_allreduce ]; then export PH_ALLREDUCE_NCCL=1; else unset PH_ALLREDUCE_NCCL; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N \
    benchmarks/bench_sharded_ops.py 2> gpurun_out/sharded_ops_$tag.err | grep "^{" > gpurun_out/sharded_ops_n${N}_$tag.jsonl
  echo "== sharded ops $tag"; python -c "
import json
for l in open('gpurun_out/sharded_ops_n${N}_$tag.jsonl'):
    d=json.loads(l); print(d['ms'], d['aggregate_gbs'], d['op'][:80], d.get('result_ok',''))"
done
