mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_4gpu.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'parity_check')}, d['e2e'])
for k in ('strong', 'configs', 'sweep'):
    for e in d.get(k, []): print(k, e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --impl reference --steps 1 --warmup 0 | cut -c1-160
