cd /root/repo
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b32.csv python tools/train_bench.py 32 1 > gpurun_out/prof4.log 2>&1
tail -2 gpurun_out/prof4.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/r2_launches_train_b32.csv 3 top > gpurun_out/r2_launches_train_b32_summary.txt; head -64 gpurun_out/r2_launches_train_b32_summary.txt
