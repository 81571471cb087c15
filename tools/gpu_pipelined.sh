mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or wide" 2>&1 | tail -3
echo "E=128 $(python bench.py --entry 128 --steps 3 --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["ms_per_step"],2))')"
python - <<'PY'
import sys, time
sys.path.insert(0,'gpu-dpf_b200'); sys.path.insert(0,'tests')
import numpy as np, torch, b200dpf, dpf
from common import random_table
for n in (1<<14, 1<<16):
    table = torch.from_numpy(random_table(n, 16, seed=1))
    d = dpf.DPF(prf=3); d.eval_init(table)
    ka, _ = b200dpf.gen_batch(np.arange(512) % n, n, np.arange(512)+7, 3)
    keys = torch.from_numpy(ka).pin_memory()
    reps = 200
    for _ in range(5): d.eval_gpu(keys)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): d.eval_gpu(keys)
    t_sync=(time.perf_counter()-t0)/reps
    list(d.eval_gpu_pipelined(keys for _ in range(5)))
    torch.cuda.synchronize(); t0=time.perf_counter()
    outs = list(d.eval_gpu_pipelined(keys for _ in range(reps)))
    t_pipe=(time.perf_counter()-t0)/reps
    print("AES n=2^%d B=512: synchronous eval_gpu %.0f DPFs/s (%.3f ms), pipelined %.0f DPFs/s (%.3f ms)" % (n.bit_length()-1, 512/t_sync, t_sync*1e3, 512/t_pipe, t_pipe*1e3))
    d.close()
PY
