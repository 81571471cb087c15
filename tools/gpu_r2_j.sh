# two item sizes: parity + sweep of the tail size
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2e24 and not config4 and not config5" 2>&1 | tail -3
for ti in 0 2 4 6 8; do for cfg in "aes128 16384 512" "aes128 16384 256" "aes128 65536 512" "aes128 65536 64" "aes128 262144 512" "salsa20 16384 512" "chacha20 65536 512" "aes128 1048576 64"; do set -- $cfg
  B200DPF_TAIL_ITEMS=$ti python bench.py --prf $1 --entries $2 --batch-per-gpu $3 --steps 30 --no-cpu-baseline --no-sweep --no-parity --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tail_items=$ti $1 n=$2 B=$3', round(d['value']), round(d['ms_per_step'],4))"
done; done | tee gpurun_out/r2_two_item_sizes.txt
python tools/gpu_phase_timing.py 16384 512 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"
