cd /root/repo
GEMM_TABLE=1 timeout 600 python tools/train_bench.py 32 3 2>&1 | tail -60
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b32.csv python tools/train_bench.py 32 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_train_b32.csv 3 > gpurun_out/r2_launches_train_b32_summary.txt
head -40 gpurun_out/r2_launches_train_b32_summary.txt
