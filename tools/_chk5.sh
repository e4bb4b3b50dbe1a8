cd /root/repo
timeout 900 python -m pytest tests/test_gpu_nms.py -x -q -m gpu 2>&1 | tail -2
for b in 8 32; do NMS_B=$b timeout 300 python tools/nms_bench.py 2>&1 | cut -c1-120; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-verify > gpurun_out/r2_bench_e.json 2> /dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_e.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','infer_e2e_value','train_value','train_ms_per_step')})
PY
