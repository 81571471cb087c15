# round 2, call A (1 GPU): full GPU test suite incl. the BASELINE-shape tests and the unmodified reference
# scripts, integer-pipe microbenchmark, default bench (sweep + parity), pipe-instruction counts per PRF,
# one full ncu capture of the ChaCha kernel (post-frontier).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2a_pytest.txt 2>&1; tail -25 gpurun_out/r2a_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
tools/microbench/bin/int_pipes > gpurun_out/r2_int_pipes.jsonl 2>&1; cat gpurun_out/r2_int_pipes.jsonl | cut -c1-260
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; cut -c1-400 gpurun_out/r2a_bench_default.json; tail -3 gpurun_out/r2a_bench_default.err
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,sm__cycles_active.avg
for prf in aes128 salsa20 chacha20; do
  ncu --metrics $M --clock-control none -k regex:dpf_eval_kernel -s 8 -c 2 --csv --log-file gpurun_out/r2a_pipes_${prf}_n20.csv \
    python bench.py --prf $prf --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > /dev/null 2>&1
  ncu --metrics $M --clock-control none -k regex:dpf_eval_kernel -s 8 -c 2 --csv --log-file gpurun_out/r2a_pipes_${prf}_n14.csv \
    python bench.py --prf $prf --entries 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:dpf_eval_kernel -s 9 -c 1 -o gpurun_out/r2a_prof_chacha \
  python bench.py --prf chacha20 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > gpurun_out/r2a_ncu_chacha.log 2>&1
ls -la gpurun_out | grep r2
