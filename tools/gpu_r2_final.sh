# round 2, final measurement set (1 GPU): sanitizer, AES block-size A/B, default bench (+reference arm), ncu launch list,
# ncu full captures, DRAM traffic per sweep config
mkdir -p gpurun_out
echo "== compute-sanitizer"; bash tools/gpu_sanitize.sh 2>&1 | tail -12
python bench.py --entries 16384 --steps 50 --no-cpu-baseline --no-sweep --no-parity --no-e2e > /dev/null 2>&1   # wake the GPU up after the sanitizer runs
echo "== AES block size A/B (384 vs 256 threads), device-timed bench"
for var in default aes256; do for cfg in "16384 512" "16384 256" "65536 64" "65536 512" "262144 512" "1048576 512"; do set -- $cfg
  if [ $var = aes256 ]; then export LD_LIBRARY_PATH=$PWD/gpu-dpf_b200/variants/aes256; else unset LD_LIBRARY_PATH; fi
  python bench.py --entries $1 --batch-per-gpu $2 --steps 20 --no-cpu-baseline --no-sweep --no-parity --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$var n=$1 B=$2', round(d['value']), round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done | tee gpurun_out/r2_aes_block_size_ab.txt
unset LD_LIBRARY_PATH
echo "== default bench"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c1-260 gpurun_out/r2_bench_default.json; tail -4 gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference.json
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_aes128_n2e20_b512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > /dev/null 2>&1
grep -c dpf_eval_kernel gpurun_out/r2_launches_aes128_n2e20_b512.csv
echo "== ncu traffic per sweep config"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__inst_executed_pipe_alu.sum,sm__cycles_elapsed.max
for prf in aes128 salsa20 chacha20; do for n in 65536 262144; do
  ncu --metrics $M --clock-control none -k regex:dpf_eval_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_traffic_${prf}_n${n}.csv \
    python bench.py --prf $prf --entries $n --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > /dev/null 2>&1
done; done
ncu --metrics $M --clock-control none -k regex:dpf_eval_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_traffic_aes128_n16384_b256.csv \
    python bench.py --entries 16384 --batch-per-gpu 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-parity > /dev/null 2>&1
ls -la gpurun_out/r2_* 2>/dev/null | awk '{print $5, $9}'
