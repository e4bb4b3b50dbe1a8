cd /root/repo
timeout 900 python tools/cl_run.py --tasks 3 --clips-per-task 32 --val-clips 16 --out gpurun_out/r2_cl_1gpu_small.json 2>&1 | grep -v Warning | tail -8 | cut -c1-600
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify 2>gpurun_out/b.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_value','train_batch2_ms_per_step')})
"
tail -3 gpurun_out/b.err
