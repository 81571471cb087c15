# round 2: ncu --set full captures of the evaluation kernel, n=2^20 B=512 (two reports per call: 64 MiB return limit)
mkdir -p gpurun_out
for prf in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_${prf}_n2e20 python bench.py --prf $prf --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > gpurun_out/r2_ncu_${prf}.log 2>&1
  tail -1 gpurun_out/r2_ncu_${prf}.log
done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
