cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/sync_probe.py 8 2>&1 | grep "distinct\|File \"/root/repo" | sort | uniq -c | sort -rn | head -14
for f in 0 1 0 1; do echo "NOSYNC=$f"; VILCO_NOSYNC_STEP=$f timeout 600 python tools/train_bench.py 32 6 2>&1 | tail -2; done
VILCO_NOSYNC_STEP=1 timeout 600 python tools/train_bench.py 2 10 2>&1 | tail -2
