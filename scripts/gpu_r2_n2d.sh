#!/bin/bash
# 2-GPU job: N-GPU agreement (both transports) incl. slicing across shards; C++ sharded spec (both transports + single)
mkdir -p gpurun_out
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PH_NO_P2P=1; else unset PH_NO_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 \
    tests/mgpu_check.py > gpurun_out/mgpu_$mode.out 2> gpurun_out/mgpu_$mode.err
  echo "== mgpu $mode rc=$?"; tail -2 gpurun_out/mgpu_$mode.out; grep -E "Error|assert" gpurun_out/mgpu_$mode.err | head -5
  rm -f /tmp/ph_id_$mode
  for r in 0 1; do RANK=$r WORLD_SIZE=2 PH_ID_FILE=/tmp/ph_id_$mode timeout 300 tests/cpp/sharded_spec > gpurun_out/sharded_spec_${mode}_r$r.log 2>&1 & done
  wait
  echo "== sharded_spec $mode"; grep -E "FAIL|expectations" gpurun_out/sharded_spec_${mode}_r0.log gpurun_out/sharded_spec_${mode}_r1.log
done
unset PH_NO_P2P
timeout 120 tests/cpp/sharded_spec > gpurun_out/sharded_spec_single.log 2>&1; echo "== sharded_spec single rc=$?"; grep -E "FAIL|expectations" gpurun_out/sharded_spec_single.log
