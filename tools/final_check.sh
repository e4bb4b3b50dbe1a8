#!/bin/bash
# what the driver runs at round end, in one go: GPU test suite, smoke(), bench (own arm + reference arm)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 300 gpurun_out/r2_bench_ref.json
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 200 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('metric','value','ms_per_step','steps','warmup','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_ms_per_step','gpu_launches','dtype')})
print('verify',d.get('verify')); print('roofline',{k:d['roofline'][k] for k in ('us_per_launch','achieved','frac','traffic')}); print('cpu_baseline',d.get('cpu_baseline')); print('clocks',d.get('clocks')); print('next_rows',json.dumps(d.get('next_rows'))[:900])
PY
