cd /root/repo
timeout 600 python tools/gemm2_probe.py check 2>&1 | grep -v "^OK" | tail -2
timeout 600 python tools/gemm2_probe.py sweep 2>&1 | cut -c1-110
