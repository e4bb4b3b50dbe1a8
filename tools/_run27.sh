cd /root/repo
timeout 800 python -m pytest tests/test_gpu_zz_nlq.py -x -q -m gpu -s 2>&1 | grep "NLQ\|passed\|failed\|Error" | head
timeout 600 python tools/nlq_bench.py 16 10 2>&1 | tail -2
timeout 600 python tools/nlq_bench.py 1 20 2>&1 | tail -1
