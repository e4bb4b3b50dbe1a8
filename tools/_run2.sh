set -x
cd /root/repo
timeout 600 python tools/gemm2_probe.py check > gpurun_out/r2_gemm2_check.log 2>&1; echo "check rc=$?"
tail -3 gpurun_out/r2_gemm2_check.log
grep -c "^OK" gpurun_out/r2_gemm2_check.log; grep "^BAD" gpurun_out/r2_gemm2_check.log | head -40
CLIPS=32 timeout 600 python tools/gemm2_probe.py time 2>&1 | grep "impl0" > gpurun_out/r2_gemm2_time.log; cat gpurun_out/r2_gemm2_time.log
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py tests/test_gpu_nms.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -40
VILCO_GEMM_TABLE=1 timeout 900 python bench.py --steps 5 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
cat gpurun_out/r2_bench_a.json | cut -c1-1500
head -60 gpurun_out/r2_bench_a.err
