mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err; echo rc=$?
cut -c1-1800 gpurun_out/scale_${N}.json; tail -5 gpurun_out/scale_${N}.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
