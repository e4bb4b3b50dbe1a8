cd /root/repo
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -s -k "single_pass_self" 2>&1 | grep "self-attention\|passed\|failed\|Error" | head -20
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 300 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_ms_per_step')})
print(d['verify']); print(d['roofline']['us_per_launch'], d['roofline']['frac'], d['roofline']['all_gemm_launches'])
PY
