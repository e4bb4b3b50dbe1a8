cd /root/repo
timeout 600 python tools/sync_probe.py 8 2>&1 | grep -v "^$" | grep -A12 "^SYNC" | grep "File \"/root/repo\|^SYNC" | cut -c1-150 | head -30
for f in 0 1; do echo "FUSED_ATTN_BWD=$f"; VILCO_FUSED_ATTN_BWD=$f timeout 600 python tools/train_bench.py 32 4 2>&1 | grep "mem GB\|ms/step" | cut -c1-200; done
