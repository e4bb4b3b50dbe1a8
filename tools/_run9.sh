cd /root/repo
timeout 600 python tools/gemm2_probe.py check 2>&1 | tail -2
for pf in 0 8 16 32; do echo "== prefetch $pf"; VILCO_GEMM_PREFETCH=$pf CLIPS=32 timeout 600 python tools/gemm2_probe.py time 2>&1 | grep "impl0" | grep "planes 1/1"; done
