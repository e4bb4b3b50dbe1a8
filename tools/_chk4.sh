cd /root/repo
timeout 900 python tools/cl_run.py --tasks 2 --clips-per-task 16 --val-clips 8 --memory-size 40 2>&1 | grep '"task"' | cut -c1-260
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/b2.err; tail -c 300 gpurun_out/b2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_ms_per_step','n_gpus')})
PY
