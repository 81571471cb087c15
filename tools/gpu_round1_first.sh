mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_aes_2e20.json 2> gpurun_out/bench_aes_2e20.err
tail -c 600 gpurun_out/bench_aes_2e20.err
for prf in chacha20 salsa20; do python bench.py --prf $prf --no-cpu-baseline > gpurun_out/bench_${prf}_2e20.json 2>/dev/null; done
for n in 16384 65536 262144; do python bench.py --entries $n --no-cpu-baseline > gpurun_out/bench_aes_$n.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_aes.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 3 -c 1 -o gpurun_out/prof_r1_aes python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_aes.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 3 -c 1 -o gpurun_out/prof_r1_chacha python bench.py --prf chacha20 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_chacha.log 2>&1
cat gpurun_out/bench_*.json
