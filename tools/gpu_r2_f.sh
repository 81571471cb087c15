# round 2, call F (1 GPU): why is the main phase less efficient with few key groups? + grouped eval timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frontier_modes or split_lanes or matches_oracle or grouped or subtree" 2>&1 | tail -3
for b in 64 128 256 512; do python tools/gpu_phase_timing.py 65536 $b 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"; done | tee gpurun_out/r2f_groups_vs_efficiency.txt
for b in 64; do for n in 16384 262144; do python tools/gpu_phase_timing.py $n $b 3 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"; done; done | tee -a gpurun_out/r2f_groups_vs_efficiency.txt
python tools/gpu_phase_timing.py 65536 64 1 2>&1 | grep -A8 "balance_top=1" | grep -E "ms/eval|tree-top|barrier passed|main phase|last block" | tee -a gpurun_out/r2f_groups_vs_efficiency.txt
python tools/gpu_batch_pir.py 2>&1 | tee gpurun_out/r2f_batch_pir.jsonl
