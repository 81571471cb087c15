mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vs_reference_gpu.py -m gpu -x -q 2>&1 | tail -5
: > gpurun_out/refgpu.jsonl
for prf in aes128 salsa20 chacha20; do for n in 16384 65536 262144; do
  timeout 600 python bench.py --impl reference-gpu --prf $prf --entries $n --steps 10 --warmup 1 2>/dev/null >> gpurun_out/refgpu.jsonl
done; done
for prf in aes128 salsa20 chacha20; do
  timeout 900 python bench.py --impl reference-gpu --prf $prf --steps 5 --warmup 1 2>/dev/null >> gpurun_out/refgpu.jsonl
done
python -c "
import json
for l in open('gpurun_out/refgpu.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:44], round(d['value'],1), round(d['ms_per_step'],2), 'init_s', round(d['config']['eval_init_s'],1))
"
python tools/benchmark_like_reference.py 2>&1 | tee gpurun_out/final_benchmark_py.txt
