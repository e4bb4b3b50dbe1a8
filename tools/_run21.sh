cd /root/repo
timeout 900 python tools/gemm2_probe.py check 2>&1 | grep -v "^OK" | tail -3
timeout 600 python tools/gemm2_probe.py time 2>&1 | grep "impl0"
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
