cd /root/repo
timeout 600 python tools/gemm2_probe.py check 2>&1 | grep -v "^OK" | tail -12
timeout 600 python tools/gemm2_probe.py check 2>&1 | grep "major" | head -10
timeout 1200 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python tools/train_bench.py 32 5 2>&1 | tail -2 | cut -c1-250
