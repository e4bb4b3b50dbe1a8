cd /root/repo
timeout 300 python tools/local_attn_bench.py mixed
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "torch_custom or local_masked" 2>&1 | tail -3
