mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -4 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_2gpu.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'parity_check')}, d['e2e'])
for k in ('strong', 'configs', 'sweep'):
    for e in d.get(k, []): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in e.items()})
PY
