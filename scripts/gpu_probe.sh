#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -2
timeout 120 python benchmarks/probe_axis.py 0,1,2 > gpurun_out/axis_shards.jsonl 2>&1; cut -c1-140 gpurun_out/axis_shards.jsonl
