for tl in 0 5 6 7; do python tools/gpu_phase_timing.py 1048576 512 3 $tl 2>&1 | grep -E "ms/eval|tree-top|barrier passed|main phase|last block"; done 2>&1 | tee gpurun_out/r2_top_item_size_n2e20.txt
for tl in 0 4 5 6; do python tools/gpu_phase_timing.py 262144 512 3 $tl 2>&1 | grep -E "ms/eval|barrier passed"; done 2>&1 | tee -a gpurun_out/r2_top_item_size_n2e20.txt
for tl in 0 4 5 6; do python tools/gpu_phase_timing.py 1048576 512 1 $tl 2>&1 | grep -E "ms/eval|barrier passed"; done 2>&1 | tee -a gpurun_out/r2_top_item_size_n2e20.txt
