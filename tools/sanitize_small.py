"""Small all-PRF run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-dpf_b200"))
import b200dpf  # noqa: E402

rng = np.random.RandomState(0)
for n, entry in ((512, 16), (2048, 40)):
    table = rng.randint(-2**31, 2**31, size=(n, entry), dtype=np.int64).astype(np.int32)
    ctx = b200dpf.Context(table)
    for prf in range(4):
        alphas = rng.randint(0, n, size=37)
        ka, kb = b200dpf.gen_batch(alphas, n, np.arange(37) + 5, prf)
        a, b = ctx.eval(ka, prf), ctx.eval(kb, prf)
        assert np.array_equal((a.astype(np.uint32) - b.astype(np.uint32)).astype(np.int32), table[alphas])
    ctx.close()
print("sanitize run ok")
