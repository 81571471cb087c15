# round 2, call G (1 GPU): ticket-range split + s_top estimator: tests, few-key-group efficiency, small-n numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2e24 and not config4 and not config5" 2>&1 | tail -3
for b in 64 128 512; do python tools/gpu_phase_timing.py 65536 $b 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"; done | tee gpurun_out/r2g_groups_vs_efficiency.txt
python tools/gpu_phase_timing.py 16384 512 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval|tree-top|barrier passed|main phase|last block" | tee -a gpurun_out/r2g_groups_vs_efficiency.txt
python tools/gpu_phase_timing.py 16384 256 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval" | tee -a gpurun_out/r2g_groups_vs_efficiency.txt
python tools/gpu_phase_timing.py 16384 64 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval" | tee -a gpurun_out/r2g_groups_vs_efficiency.txt
python tools/gpu_phase_timing.py 16384 512 1 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval" | tee -a gpurun_out/r2g_groups_vs_efficiency.txt
echo "== split on/off A/B via bench (device-timed)"
for sp in 1 0; do for cfg in "16384 512" "16384 256" "65536 512" "65536 64" "1048576 512"; do set -- $cfg
  B200DPF_SPLIT_TICKETS=$sp python bench.py --entries $1 --batch-per-gpu $2 --steps 30 --no-cpu-baseline --no-sweep --no-parity --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('split=$sp n=$1 B=$2', round(d['value']), round(d['ms_per_step'],4))"
done; done | tee gpurun_out/r2g_split_ab.txt
