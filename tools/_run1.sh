set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python tools/gemm2_probe.py check > gpurun_out/r2_gemm2_check.log 2>&1; echo "check rc=$?"
tail -5 gpurun_out/r2_gemm2_check.log
grep -c "^OK" gpurun_out/r2_gemm2_check.log; grep "^BAD" gpurun_out/r2_gemm2_check.log | head -40
timeout 600 python tools/gemm2_probe.py time > gpurun_out/r2_gemm2_time.log 2>&1; echo "time rc=$?"
cat gpurun_out/r2_gemm2_time.log | tail -60
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s 2>&1 | tail -30
