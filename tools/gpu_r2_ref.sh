mkdir -p gpurun_out
python tools/benchmark_like_reference.py 2>&1 | tee gpurun_out/r2_benchmark_py_equivalent.txt
: > gpurun_out/r2_refgpu.jsonl
for prf in aes128 salsa20 chacha20; do for n in 16384 65536; do
  timeout 600 python bench.py --impl reference-gpu --prf $prf --entries $n --steps 10 --warmup 1 2>/dev/null >> gpurun_out/r2_refgpu.jsonl
done; done
python -c "
import json
for l in open('gpurun_out/r2_refgpu.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:44], round(d['value'],1), round(d['ms_per_step'],2))
"
