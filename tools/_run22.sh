cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 600 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','infer_e2e_value','train_value','train_ms_per_step','train_batch2_ms_per_step','verify')})
print(d['roofline'])
PY
