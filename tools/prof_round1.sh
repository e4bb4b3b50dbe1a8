set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_eval_final.csv python tools/profile_step.py 8 bf16x3 2 > gpurun_out/prof_eval.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_train.csv python tools/train_bench.py 4 1 > gpurun_out/prof_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 1 --launch-skip 2 -f -o gpurun_out/r1_gemm_heads python tools/one_gemm.py heads > gpurun_out/prof_heads.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 1 --launch-skip 2 -f -o gpurun_out/r1_gemm_wgrad python tools/one_gemm.py wgrad > gpurun_out/prof_wgrad.log 2>&1
ncu -i gpurun_out/r1_gemm_heads.ncu-rep --page raw --csv > gpurun_out/r1_gemm_heads_raw.csv 2>/dev/null
ncu -i gpurun_out/r1_gemm_wgrad.ncu-rep --page raw --csv > gpurun_out/r1_gemm_wgrad_raw.csv 2>/dev/null
ls -la gpurun_out | tail -15
tail -2 gpurun_out/prof_train.log
