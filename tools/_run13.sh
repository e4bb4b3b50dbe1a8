cd /root/repo
GEMM_TABLE=1 timeout 600 python tools/train_bench.py 32 1 2>&1 | tail -32 | cut -c1-200
timeout 600 python tools/train_bench.py 32 5 2>&1 | tail -3 | cut -c1-250
