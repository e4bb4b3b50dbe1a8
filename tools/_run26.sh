cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_eval_step.csv python tools/profile_step.py 32 mixed 2 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_eval_step.csv 2 top > gpurun_out/r2_launches_eval_step_summary.txt
head -32 gpurun_out/r2_launches_eval_step_summary.txt
timeout 600 python tools/train_bench.py 32 5 2>&1 | tail -2
