mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c1-200 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
for prf in salsa20 chacha20; do python bench.py --prf $prf --steps 20 --no-cpu-baseline --no-sweep 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$prf', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['pipe']['frac'], d['parity_check']['ok'])"; done | tee gpurun_out/r2_bench_salsa_chacha.txt
