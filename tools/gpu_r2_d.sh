# round 2, call D (1 GPU): grouped evaluation, TMA-rows A/B, balanced top + ticket prefetch, chunked wide entries
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q --durations=5 -k "not 2e24 and not config4 and not config5" ) > gpurun_out/r2d_pytest.txt 2>&1; tail -12 gpurun_out/r2d_pytest.txt
for prf in 3 1; do python tools/gpu_phase_timing.py 16384 512 $prf; done 2>&1 | grep -v "balance_top=0" -A0 | tee gpurun_out/r2d_phase_timing.txt | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"
python tools/gpu_phase_timing.py 16384 256 3 2>&1 | tee -a gpurun_out/r2d_phase_timing.txt | grep -E "ms/eval"
python tools/gpu_phase_timing.py 65536 64 3 2>&1 | tee -a gpurun_out/r2d_phase_timing.txt | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"
echo "== subtree size sweep, n=2^14 / 2^16, B=512 (device-timed bench)"
for n in 16384 65536; do for prf in aes128 salsa20; do for s in 2 3 4 5; do
  python bench.py --prf $prf --entries $n --steps 30 --subtree-log2 $s --no-cpu-baseline --no-sweep --no-parity --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$prf n=$n s=$s', round(d['value']), round(d['ms_per_step'],4))"
done; done; done | tee gpurun_out/r2d_s_sweep.txt
echo "== TMA-staged rows A/B (Salsa20 / ChaCha20, n=2^24 > L2 and n=2^20)"
for prf in salsa20 chacha20; do for n in 16777216 1048576; do for tma in 0 1; do
  B200DPF_TMA_ROWS=$tma python bench.py --prf $prf --entries $n --steps 3 --no-cpu-baseline --no-sweep --no-parity --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$prf n=$n tma_rows=$tma', round(d['value'],1), 'DPFs/s', round(d['ms_per_step'],3), 'ms')"
done; done; done | tee gpurun_out/r2d_tma_rows_ab.txt
echo "== batch PIR"
python tools/gpu_batch_pir.py 2>&1 | tee gpurun_out/r2d_batch_pir.jsonl
python tools/gpu_hostpath.py 2>&1 | tee gpurun_out/r2d_hostpath.txt
echo "== wide entries on one GPU: E=128 B=8192 n=2^20 (leaf cache 32 GiB > 16 GiB cap -> 2 chunks)"
python bench.py --entry 128 --batch-per-gpu 8192 --steps 2 --no-cpu-baseline --no-sweep --no-e2e 2>/dev/null | cut -c1-330
