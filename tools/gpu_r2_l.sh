mkdir -p gpurun_out
mkdir -p gpu-dpf_b200/variants/current && cp gpu-dpf_b200/libb200dpf.so gpu-dpf_b200/variants/current/
for rep in 1 2; do for var in r2b c_3644041 current; do python tools/ab_ctypes_timing.py gpu-dpf_b200/variants/$var/libb200dpf.so 1048576 3 2>&1 | grep "ms/eval"; done; done | tee gpurun_out/r2_kernel_regression_fix.txt
for var in r2b c_3644041 current; do for n in 16384 262144; do python tools/ab_ctypes_timing.py gpu-dpf_b200/variants/$var/libb200dpf.so $n 3 2>&1 | grep "ms/eval"; done; python tools/ab_ctypes_timing.py gpu-dpf_b200/variants/$var/libb200dpf.so 1048576 1 2>&1 | grep "ms/eval"; done | tee -a gpurun_out/r2_kernel_regression_fix.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2e24 and not config4 and not config5" 2>&1 | tail -2
