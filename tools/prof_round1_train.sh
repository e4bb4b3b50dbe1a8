set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/ln_bwd_bench.py > gpurun_out/r1_bandwidth_kernels.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_train.csv python tools/train_bench.py 4 1 > gpurun_out/prof_train.log 2>&1
cat gpurun_out/r1_bandwidth_kernels.txt
