mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dpf_mac_tma_kernel -s 6 -c 1 -o gpurun_out/final_prof_mac_tma python bench.py --entry 128 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_mac.log 2>&1
python bench.py --entries 16384 --steps 50 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n=2^14', round(d['value'],1), round(d['e2e']['value'],1))"
python tools/benchmark_like_reference.py 2>&1 | head -3
