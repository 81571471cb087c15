"""The reference's benchmark.py sweep (benchmark.py:4-7 -> dpf.test_gpu_dpf_perf, dpf.py:286-320)
run against this engine's dpf module: N in {2^14..2^20} x {AES128, SALSA20, CHACHA20}, batch 512,
entry 16, wall clock around 10 eval_gpu calls with a Python list of 512 CPU key tensors."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpu-dpf_b200"))
import dpf  # noqa: E402

for N in [16384, 65536, 262144, 1048576]:
    dpf.test_gpu_dpf_perf(N=N, prf=dpf.DPF.PRF_AES128)
    dpf.test_gpu_dpf_perf(N=N, prf=dpf.DPF.PRF_SALSA20)
    dpf.test_gpu_dpf_perf(N=N, prf=dpf.DPF.PRF_CHACHA20)
