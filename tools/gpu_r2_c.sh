# round 2, call C (2 GPUs): balanced tree-top + gather path on one GPU, then the multi-GPU paths
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q --durations=5 -k "not 2e24 and not config4 and not config5" ) > gpurun_out/r2c_pytest.txt 2>&1; tail -12 gpurun_out/r2c_pytest.txt
for prf in 3 1; do python tools/gpu_phase_timing.py 16384 512 $prf; done 2>&1 | tee gpurun_out/r2c_phase_timing.txt
python tools/gpu_phase_timing.py 65536 512 3 2>&1 | tee -a gpurun_out/r2c_phase_timing.txt
python tools/gpu_phase_timing.py 16384 256 3 2>&1 | tee -a gpurun_out/r2c_phase_timing.txt
python tools/gpu_hostpath.py 2>&1 | tee gpurun_out/r2c_hostpath.txt
python tools/gpu_multi_single_process.py 2>&1 | tee gpurun_out/r2c_multi_single_process.txt
for n in 65536 1048576; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --strong --entries $n --steps 20 --no-sweep 2>/dev/null | cut -c1-700
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 2> gpurun_out/r2c_bench2.err > gpurun_out/r2c_bench2.json; cut -c1-300 gpurun_out/r2c_bench2.json; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench2.json')); print(d.get('parity_check')); print(d.get('strong')); print(d.get('sweep'))"
tail -5 gpurun_out/r2c_bench2.err
