cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py -q -m gpu 2>&1 | grep -v "^$" | grep -v "Warning\|warnings.warn\|^  " | tail -15
VILCO_GEMM_TABLE=1 timeout 900 python bench.py --steps 10 --warmup 3 --train-batch 0 --no-cpu-baseline --no-verify > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_b.json'))
print({k:j[k] for k in ('value','ms_per_step','infer_e2e_value','gpu_launches','latency_b1_ms')}, j['roofline']['frac'], j['roofline']['all_gemm_launches'], j['clocks'])
PY
