timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$i bench.py --gpus 2 --strong --entries 65536 --steps 30 --no-sweep 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('strong n=2^16 2 GPUs', d['config']['parallelism'], round(d['ms_per_step'],4), 'ms', d['parity_check']['ok'])"; done
