# round 2, call B (1 GPU): single-launch pipeline + packed keys -- tests, A/B against the legacy
# 4-operation pipeline, int-pipe microbenchmark (fixed), per-pipe instruction counts.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2b_pytest.txt 2>&1; tail -15 gpurun_out/r2b_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
tools/microbench/bin/int_pipes > gpurun_out/r2_int_pipes.jsonl 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2_int_pipes.jsonl'):
    d=json.loads(l)
    if d['kind']=='pipe': print(d['name'], d['steps_per_clk_sm'], 'alu', d['alu_inst_per_clk_sm'], 'fma', d['fma_inst_per_clk_sm'], 'lsu', d['lsu_inst_per_clk_sm'], d['err'])
    elif d['kind']=='quarter_round': print(d['cipher'], d['fma_rot_mask'], d['clk_per_qr_sm'], d['err'])
PY
: > gpurun_out/r2b_ab.jsonl
for one in 1 0; do for prf in aes128 salsa20 chacha20; do for n in 16384 65536 1048576; do
  B200DPF_ONE_LAUNCH=$one python bench.py --prf $prf --entries $n --steps 20 --no-cpu-baseline --no-sweep --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'one_launch':$one,'prf':'$prf','n':$n,'value':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'launches':d['gpu_launches']}))" >> gpurun_out/r2b_ab.jsonl
done; done; done
cat gpurun_out/r2b_ab.jsonl
B200DPF_ONE_LAUNCH=1 python bench.py --entries 16384 --batch-per-gpu 256 --steps 40 --no-cpu-baseline --no-sweep 2>/dev/null | cut -c1-600
python tools/gpu_hostpath.py 2>&1 | tail -20
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,sm__cycles_active.avg
for prf in aes128 salsa20 chacha20; do for n in 1048576 16384; do
  ncu --metrics $M --clock-control none -k regex:dpf_eval_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2b_pipes_${prf}_n${n}.csv \
    python bench.py --prf $prf --entries $n --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > gpurun_out/r2b_pipes_${prf}_n${n}.log 2>&1
done; done
tail -3 gpurun_out/r2b_pipes_aes128_n1048576.csv | cut -c1-400
ls gpurun_out | grep r2b | head -30
