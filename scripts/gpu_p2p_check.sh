#!/bin/bash
# 2+ GPU job: the N-GPU agreement script under both transports (peer-mapped memory / NCCL)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PH_NO_P2P=1; else unset PH_NO_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N \
    tests/mgpu_check.py > gpurun_out/mgpu_$mode.out 2> gpurun_out/mgpu_$mode.err
  echo "== $mode rc=$?"; tail -3 gpurun_out/mgpu_$mode.out; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/mgpu_$mode.err | tail -15
done
