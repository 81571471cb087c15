mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2e_bench_8gpu.json 2> gpurun_out/r2e_bench_8gpu.err
tail -4 gpurun_out/r2e_bench_8gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2e_bench_8gpu.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'parity_check')}, d['e2e'])
for k in ('strong', 'configs', 'sweep'):
    for e in d.get(k, []): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in e.items() if a in ('config','prf','n','batch','axis','value','ms_per_step','e2e','parity_ok','speedup_vs_1gpu','error')})
PY
