cd /root/repo
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "window_attention or local_masked" 2>&1 | tail -2
for tq in 32 64 128; do echo TQ $tq; VILCO_LT_TQ=$tq timeout 300 python tools/local_attn_bench.py mixed; done
timeout 300 python tools/local_attn_bench.py fp16x3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_attn_tc -c 1 -o gpurun_out/r2_local_attn python tools/local_attn_bench.py mixed > /dev/null 2>&1
ncu -i gpurun_out/r2_local_attn.ncu-rep --page raw --csv > gpurun_out/r2_local_attn_raw.csv 2>/dev/null
