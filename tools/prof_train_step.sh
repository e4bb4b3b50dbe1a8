#!/bin/bash
# per-shape GEMM table (CUDA events) and ncu launch list of one 32-clip training step
cd "$(dirname "$0")/.."
GEMM_TABLE=1 timeout 600 python tools/train_bench.py 32 3 2>&1 | grep -v Warning | tail -45
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b32.csv python tools/train_bench.py 32 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_train_b32.csv 3 > gpurun_out/r2_launches_train_b32_summary.txt
head -36 gpurun_out/r2_launches_train_b32_summary.txt
