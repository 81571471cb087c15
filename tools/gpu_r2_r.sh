timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "api or abi or packed or batch_edges or streams or diagnostics" 2>&1 | tail -3
python tools/gpu_hostpath.py 2>&1 | tee gpurun_out/r2_hostpath_final.txt
