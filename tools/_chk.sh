cd /root/repo
timeout 600 python tools/gemm2_probe.py check 2>&1 | grep -v "^OK" | tail -2
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_bwd_ops.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/train_bench.py 32 5 2>&1 | tail -2
VILCO_FUSED_ATTN_BWD=0 timeout 600 python tools/train_bench.py 32 5 2>&1 | tail -1
timeout 600 python tools/gemm2_probe.py sweep 2>&1 | cut -c1-110 | head -12
