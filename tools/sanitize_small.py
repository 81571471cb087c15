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
    packed = b"".join(b200dpf.key_pack(k) for k in ka)
    assert np.array_equal(ctx.eval_packed(packed, 37, 3), a)            # compact keys, last prf's batch
    assert np.array_equal(ctx.eval_gather(list(ka), 3), a)              # gather path
    if entry == 16:
        ctx.set_option("tma_rows", 1)                                      # TMA-staged rows variant (non-AES)
        ka2, kb2 = b200dpf.gen_batch(alphas, n, np.arange(37) + 5, 1)
        a2, b2 = ctx.eval(ka2, 1), ctx.eval(kb2, 1)
        assert np.array_equal((a2.astype(np.uint32) - b2.astype(np.uint32)).astype(np.int32), table[alphas])
    ctx.close()
# grouped (batch-PIR) launch
tables = [rng.randint(-2**31, 2**31, size=(m, 16), dtype=np.int64).astype(np.int32) for m in (256, 1024, 64)]
grp = b200dpf.GroupContext(tables)
bins = np.array([0, 1, 2, 1, 0, 1] * 8, np.int32)
alphas = np.array([rng.randint(0, tables[g].shape[0]) for g in bins])
ks = [b200dpf.gen(int(al), tables[g].shape[0], 77 + i, 2) for i, (al, g) in enumerate(zip(alphas, bins))]
ga = grp.eval(np.stack([k[0] for k in ks]), bins, 2)
gb = grp.eval(np.stack([k[1] for k in ks]), bins, 2)
assert np.array_equal((ga.astype(np.uint32) - gb.astype(np.uint32)).astype(np.int32), np.stack([tables[g][al] for g, al in zip(bins, alphas)]))
grp.close()
print("sanitize run ok")
