cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/cl_run.py --tasks 2 --clips-per-task 32 --val-clips 8 2>&1 | grep -v "Warning\|warn" | tail -4 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/b2.err; echo rc=$?
python -c "
import json
j=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('n_gpus','value','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_value','train_batch2_ms_per_step')})
print(j['train']['allreduce'], j['train']['grad_bytes_allreduced_per_step'])
"
grep -v "Warning\|warn" gpurun_out/b2.err | tail -5
