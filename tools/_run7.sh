cd /root/repo
(time timeout 1500 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err); echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_c.json'))
for k in ('metric','value','ms_per_step','infer_e2e_value','train_value','train_ms_per_step','train_e2e_value','train_batch2_value','train_batch2_ms_per_step','gpu_launches','latency_b1_ms','verify','cpu_baseline','eager_gpu_baseline','clocks'):
    print(k, j.get(k))
print('roofline', {k:j['roofline'][k] for k in ('shape','us_per_launch','achieved','frac')})
t=j['train']; print('train', {k:t[k] for k in ('value','ms_per_step','gpu_launches_per_step','peak_mem_gib','last_loss','batch2')})
print(len(json.dumps(j)))
PY
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_c.err); echo "ref rc=$?"
cut -c1-600 gpurun_out/r2_bench_ref.json
