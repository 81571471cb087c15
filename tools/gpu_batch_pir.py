"""Batch-PIR throughput: Q queries x G bins, one DPF key per (query, bin), all
pairs in ONE grouped launch (b200dpf_group_eval) against the same pairs evaluated bin by bin with one ordinary
context per bin (what the reference API would force).  Shapes follow the paper's co-design space: many bins of
2^10..2^14 entries (paper/experimental/batch_pir/batch_pir_optimization.py:84-87: a query costs one DPF per bin)."""
import sys, time, json
sys.path.insert(0, "gpu-dpf_b200"); sys.path.insert(0, "tests")
import numpy as np, torch, b200dpf
from common import random_table

def t(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps

prf = 3
for G, logn, Q in ((64, 12, 256), (256, 10, 256), (16, 14, 512), (64, 12, 32), (1024, 10, 64)):
    n = 1 << logn
    tables = [random_table(n, 16, seed=g) for g in range(G)]
    alphas = np.random.RandomState(G).randint(0, n, size=Q * G)
    ka, kb = b200dpf.gen_batch(alphas, n, np.arange(Q * G) + 5, prf)
    bins = np.tile(np.arange(G, dtype=np.int32), Q)
    grp = b200dpf.GroupContext(tables)
    got = grp.eval(ka, bins, prf)
    rec = (got.astype(np.uint32) - grp.eval(kb, bins, prf).astype(np.uint32)).astype(np.int32)
    assert np.array_equal(rec, np.stack([tables[g][a] for g, a in zip(bins, alphas)]))
    s_grouped = t(lambda: grp.eval(ka, bins, prf), 10)
    dev_ms = grp.last_device_ms
    grp.close()
    ctxs = [b200dpf.Context(tb) for tb in tables[:min(G, 64)]]
    per_bin_keys = [np.ascontiguousarray(ka[g::G]) for g in range(len(ctxs))]
    def bin_by_bin():
        for c, k in zip(ctxs, per_bin_keys): c.eval(k, prf)
    s_loop = t(bin_by_bin, 3) * (G / len(ctxs))
    for c in ctxs: c.close()
    print(json.dumps({"bins": G, "bin_entries": n, "queries": Q, "pairs": Q * G, "prf": "AES128",
                      "grouped_ms": s_grouped * 1e3, "grouped_device_ms": dev_ms, "dpfs_per_s_device": Q * G / dev_ms * 1e3, "queries_per_s_grouped": Q / s_grouped, "dpfs_per_s_grouped": Q * G / s_grouped,
                      "bin_by_bin_ms": s_loop * 1e3, "queries_per_s_bin_by_bin": Q / s_loop, "speedup": s_loop / s_grouped}), flush=True)
