"""One process driving every GPU of the box through ONE context (b200dpf_create_multi): strong scaling of a
512-key batch, host buffers in and out (the b200dpf_eval call), against the same call on one GPU."""
import sys, time
sys.path.insert(0, "gpu-dpf_b200"); sys.path.insert(0, "tests")
import numpy as np, torch, b200dpf
from common import random_table

ndev = torch.cuda.device_count()
devs = list(range(1 << (ndev.bit_length() - 1)))
print("devices:", devs)
def t(fn, reps):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
for prf in (3,):
    for n in (1 << 14, 1 << 16, 1 << 18, 1 << 20):
        table = random_table(n, 16, seed=1)
        batch = 512
        ka, _ = b200dpf.gen_batch(np.arange(batch) * 7 % n, n, np.arange(batch) + 7, prf)
        pinned = torch.from_numpy(ka).pin_memory().numpy()
        reps = 200 if n <= 1 << 16 else 30
        one = b200dpf.Context(table)
        ref = one.eval(pinned, prf)
        ms1 = t(lambda: one.eval(pinned, prf), reps)
        one.close()
        line = "n=2^%d AES B=512 host->host: 1 GPU %.3f ms" % (n.bit_length() - 1, ms1)
        for axis, name in ((1, "entries"), (2, "keys")):
            m = b200dpf.Context.multi(table, devs, axis)
            got = m.eval(pinned, prf)
            assert np.array_equal(got, ref), (n, name)
            ms = t(lambda: m.eval(pinned, prf), reps)
            line += " | %d GPUs %s %.3f ms (%.2fx)" % (len(devs), name, ms, ms1 / ms)
            m.close()
        print(line, flush=True)
