mkdir -p gpurun_out
mkdir -p gpu-dpf_b200/variants/both && cp gpu-dpf_b200/libb200dpf.so gpu-dpf_b200/variants/both/
for rep in 1 2; do for var in base quads cwlo both; do python tools/ab_ctypes_timing.py gpu-dpf_b200/variants/$var/libb200dpf.so 1048576 3 2>&1 | grep "ms/eval"; done; done | tee gpurun_out/r2_aes_quads_cwlo_ab.txt
for var in base both; do for n in 16384 65536; do python tools/ab_ctypes_timing.py gpu-dpf_b200/variants/$var/libb200dpf.so $n 3 2>&1 | grep "ms/eval"; done; done | tee -a gpurun_out/r2_aes_quads_cwlo_ab.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2e24 and not config4 and not config5" 2>&1 | tail -3
