# final round-1 measurement set (1 GPU): tests, smoke, default bench (+reference arm), sweeps, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; cut -c1-300 gpurun_out/final_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_reference.json
: > gpurun_out/final_sweep.jsonl
for prf in aes128 salsa20 chacha20; do for n in 16384 65536 262144 1048576; do
  python bench.py --prf $prf --entries $n --no-cpu-baseline 2>/dev/null >> gpurun_out/final_sweep.jsonl
done; done
python bench.py --prf salsa20 --entries 16777216 --steps 2 --no-cpu-baseline 2>/dev/null >> gpurun_out/final_sweep.jsonl
python bench.py --entry 128 --steps 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/final_sweep.jsonl
python bench.py --entries 16384 --batch-per-gpu 256 --steps 50 --no-cpu-baseline 2>/dev/null >> gpurun_out/final_sweep.jsonl
python -c "
import json
for l in open('gpurun_out/final_sweep.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:52], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
python tools/benchmark_like_reference.py 2>&1 | tee gpurun_out/final_benchmark_py.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches_aes.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 6 -c 2 -o gpurun_out/final_prof_aes python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_aes.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 6 -c 2 -o gpurun_out/final_prof_salsa python bench.py --prf salsa20 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_salsa.log 2>&1
ls -la gpurun_out | tail -12
