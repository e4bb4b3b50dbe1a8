cd /root/repo
# (1) whole-step launch list of the captured evaluation graph (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_eval_step.csv python tools/profile_step.py 32 mixed 2 > gpurun_out/prof.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_eval_step.csv 2 top > gpurun_out/r2_launches_eval_step_summary.txt; head -32 gpurun_out/r2_launches_eval_step_summary.txt
# (2) ncu --set full of the dominant GEMM and of the fused XLNet attention
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_gemm_heads -f python tools/one_gemm.py heads mixed > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xl_attn_kernel -s 2 -c 1 -o gpurun_out/r2_xl_attn -f python tools/one_gemm.py xl mixed > gpurun_out/ncu2.log 2>&1
for f in r2_gemm_heads r2_xl_attn; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep
# (3) GPU busy time of a batch-2 training step (sum of kernel durations)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b2.csv python tools/train_bench.py 2 1 > gpurun_out/prof3.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_train_b2.csv 3 > gpurun_out/r2_launches_train_b2_summary.txt; head -30 gpurun_out/r2_launches_train_b2_summary.txt
