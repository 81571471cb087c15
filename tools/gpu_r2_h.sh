# round 2, call H (1 GPU): ncu full captures -- few key groups (n=2^16 B=64) vs many (B=512), and n=2^14 B=512
mkdir -p gpurun_out
for cfg in "65536 64" "65536 512"; do set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 2 -c 1 -f -o gpurun_out/r2h_prof_aes_n$1_b$2 \
    python bench.py --entries $1 --batch-per-gpu $2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > gpurun_out/r2h_ncu_$1_$2.log 2>&1
  tail -2 gpurun_out/r2h_ncu_$1_$2.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -5
