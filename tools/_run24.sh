cd /root/repo
timeout 600 python tools/_sync_probe.py 8 2>&1 | grep "distinct\|File \"/root/repo" | sort | uniq -c | sort -rn | head -30
timeout 600 python tools/train_bench.py 32 5 2>&1 | tail -2
timeout 600 python tools/train_bench.py 2 10 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -2
