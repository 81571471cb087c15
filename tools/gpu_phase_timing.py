"""Where a single-launch evaluation spends its time (n, batch, prf from argv): per-block %globaltimer
stamps of one launch -> phase durations.  Run on the GPU box."""
import sys
sys.path.insert(0, "gpu-dpf_b200"); sys.path.insert(0, "tests")
import numpy as np, torch, b200dpf
from common import random_table

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
prf = int(sys.argv[3]) if len(sys.argv) > 3 else 3
top_log2 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
table = random_table(n, 16, seed=1)
ka, _ = b200dpf.gen_batch(np.arange(batch) % n, n, np.arange(batch) + 7, prf)
kd = torch.from_numpy(ka).cuda()
out = torch.empty((batch, 16), dtype=torch.int32, device="cuda")
ctx = b200dpf.Context(table)
ctx.set_option("top_log2", top_log2)
stream = torch.cuda.current_stream().cuda_stream
for balance in (1, 0):
    ctx.set_option("balance_top", balance)
    ctx.set_option("timing", 0)
    for _ in range(20):
        ctx.eval_device(kd.data_ptr(), batch, prf, out.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ctx.eval_device(kd.data_ptr(), batch, prf, out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    ctx.set_option("timing", 1)
    ctx.eval_device(kd.data_ptr(), batch, prf, out.data_ptr(), stream)
    torch.cuda.synchronize()
    t = ctx.read_timing().astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    us = lambda a: a / 1e3
    print("top_log2=%d " % top_log2, end="")
    print("n=%d batch=%d prf=%d balance_top=%d: %.4f ms/eval (%d DPFs/s), %d blocks" % (n, batch, prf, balance, ms, batch / ms * 1e3, len(t)))
    print("   start skew            max %.1f us" % us((t[:, 0] - t0).max()))
    print("   tables + clears       median %.1f  max %.1f us" % (us(np.median(t[:, 1] - t[:, 0])), us((t[:, 1] - t[:, 0]).max())))
    print("   tree-top work         median %.1f  max %.1f  min %.1f us" % (us(np.median(t[:, 2] - t[:, 1])), us((t[:, 2] - t[:, 1]).max()), us((t[:, 2] - t[:, 1]).min())))
    print("   barrier wait          median %.1f  max %.1f us" % (us(np.median(t[:, 3] - t[:, 2])), us((t[:, 3] - t[:, 2]).max())))
    print("   barrier passed at     %.1f us after first start" % us(t[:, 3].max() - t0))
    print("   main phase            median %.1f  max %.1f  min %.1f us" % (us(np.median(t[:, 4] - t[:, 3])), us((t[:, 4] - t[:, 3]).max()), us((t[:, 4] - t[:, 3]).min())))
    print("   last block end        %.1f us after first start; earliest block end %.1f us" % (us(t[:, 4].max() - t0), us(t[:, 4].min() - t0)))
ctx.close()
