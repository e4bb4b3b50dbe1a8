cd /root/repo
for t in 512 384 256; do echo "== threads $t"; NMS_B=32 VILCO_NMS_THREADS=$t timeout 300 python tools/nms_bench.py 2>&1 | tail -1 | cut -c1-120; done
echo "== auto"; NMS_B=32 timeout 300 python tools/nms_bench.py 2>&1 | tail -1 | cut -c1-120
timeout 900 python -m pytest tests/test_gpu_nms.py -q -m gpu 2>&1 | tail -2
