# round 2, final 8-GPU numbers: the driver's bench line at N=8 (weak + strong + configs 4/5, all parity-checked) and the
# single-process multi-device context
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2e_bench_8gpu.json 2> gpurun_out/r2e_bench_8gpu.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2e_bench_8gpu.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'parity_check')}, d['e2e'])
    for k in ('strong', 'configs', 'sweep'):
        for e in d.get(k, []): print(k, e)
except Exception as exc:
    print('bench 8gpu failed', exc); print(open('gpurun_out/r2e_bench_8gpu.err').read()[-2000:])
PY
timeout 600 python tools/gpu_multi_single_process.py 2>&1 | tee gpurun_out/r2e_multi_single_process.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
