mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu-baseline --entry 128 --batch-per-gpu 1024 --steps 3 > gpurun_out/c5_new.json 2> gpurun_out/c5_new.err
python -c "
import json; d=json.loads(open('gpurun_out/c5_new.json').read()); print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['gpu_launches'])"
