cd /root/repo
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "window_attention or local_masked" -s 2>&1 | grep -v "^$" | tail -25
for tq in 32 64 128; do echo TQ $tq; VILCO_LT_TQ=$tq timeout 300 python tools/local_attn_bench.py mixed; done
echo two planes; timeout 300 python tools/local_attn_bench.py fp16x3
