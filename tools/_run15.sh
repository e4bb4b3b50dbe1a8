cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/cl_run.py --out gpurun_out/r2_cl_8gpu.json 2>&1 | grep -v "Warning\|warn" | tail -7 | cut -c1-500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/b8.err; echo rc=$?
python -c "
import json
j=json.loads(open('gpurun_out/r2_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('n_gpus','value','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_value','train_batch2_ms_per_step')})
print(j['train']['allreduce'], j['train']['grad_bytes_allreduced_per_step'])
"
tail -3 gpurun_out/b8.err
