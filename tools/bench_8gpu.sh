#!/bin/bash
# bench.py on the 8 GPUs of one box, launched the way the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps ${STEPS:-5} --warmup 3 --no-verify > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/b8.err
tail -c 300 gpurun_out/b8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_value','train_batch2_ms_per_step','n_gpus')}, d['config'].get('videos_per_step_per_gpu'))
PY
