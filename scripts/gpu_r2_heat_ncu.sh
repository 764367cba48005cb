#!/bin/bash
# ncu --set full of the two-step stencil on a (256,2048,2048) slab: shuffle form (default) vs the round-1 scalar-LDS form (cfg 8)
mkdir -p gpurun_out
for cfg in 0 8; do
  PH_HEAT_TB_CFG=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:"heat_tma2" -s 1 -c 1 -o /tmp/heat_cfg$cfg -f \
    python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 --reps 1 --warm 0 --inner 1 > gpurun_out/ncu_heat_cfg$cfg.log 2>&1
  python benchmarks/ncu_summary.py /tmp/heat_cfg$cfg.ncu-rep gpurun_out/ncu_heat_cfg$cfg.csv
  ncu -i /tmp/heat_cfg$cfg.ncu-rep --page raw --csv > gpurun_out/ncu_heat_cfg${cfg}_raw.csv 2>/dev/null
done
python - <<'PY'
import csv
for cfg in (0, 8):
    rows = list(csv.reader(open(f"gpurun_out/ncu_heat_cfg{cfg}_raw.csv")))
    h, v = rows[0], rows[2]
    want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
            "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active"]
    print("cfg", cfg)
    for w in want:
        if w in h: print("  ", w, v[h.index(w)])
PY
CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py --steps 10 --warm 2
PH_HEAT_TB_CFG=8 CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py --steps 10 --warm 2
PH_HEAT_TB_CFG=8 CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py
CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/bench_heat_sustained.py
timeout 300 python benchmarks/bench_kernels.py --only "heat 3-D" 
