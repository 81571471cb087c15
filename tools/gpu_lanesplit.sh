mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, time, os
sys.path.insert(0, 'gpu-dpf_b200'); sys.path.insert(0,'tests')
import numpy as np, torch, b200dpf
from common import random_table
for prf, name in ((3,'AES128'),(1,'SALSA20')):
  for n in (1<<14, 1<<20):
    table = random_table(n, 16, seed=1)
    ctx = b200dpf.Context(table)
    for B in (1, 4, 16, 32):
        ka, _ = b200dpf.gen_batch(np.arange(B) % n, n, np.arange(B)+7, prf)
        kd = torch.from_numpy(ka).cuda(); out = torch.empty((B,16), dtype=torch.int32, device='cuda')
        res = {}
        for split in ('1','0'):
            os.environ['B200DPF_LANE_SPLIT'] = split
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(3): ctx.eval_device(kd.data_ptr(), B, prf, out.data_ptr(), st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ctx.eval_device(kd.data_ptr(), B, prf, out.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            res[split] = e0.elapsed_time(e1)/10
        t0=time.perf_counter()
        for _ in range(10): ctx.eval(ka, prf)
        host_ms=(time.perf_counter()-t0)*100
        print("%s n=2^%d B=%d: device %.3f ms split / %.3f ms lane=key  (host-buffer call %.3f ms)" % (name, n.bit_length()-1, B, res['1'], res['0'], host_ms))
    ctx.close()
PY
