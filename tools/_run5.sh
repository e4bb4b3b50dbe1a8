cd /root/repo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_eval_step_b32.csv python tools/profile_step.py 32 mixed 2 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
python tools/summarize_launches.py gpurun_out/r2_launches_eval_step_b32.csv 2 top > gpurun_out/r2_launches_eval_step_b32_summary.txt
cat gpurun_out/r2_launches_eval_step_b32_summary.txt | head -70
