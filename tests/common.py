"""Helpers shared by the tests: deterministic inputs."""
import random

import numpy as np


def formula_table(n, e):
    """Closed-form table used by the golden fixtures (tests/golden/make_golden.py)."""
    i = np.arange(n, dtype=np.uint64).reshape(n, 1)
    c = np.arange(e, dtype=np.uint64).reshape(1, e)
    v = (i * np.uint64(2654435761) + c * np.uint64(40503) + (i * c) * np.uint64(97) + np.uint64(12345)) & np.uint64(0x7FFFFFFF)
    return v.astype(np.int64).astype(np.int32)


def random_table(n, e, seed=1234, full_range=True):
    rng = np.random.RandomState(seed)
    if full_range:
        return rng.randint(-2**31, 2**31, size=(n, e), dtype=np.int64).astype(np.int32)
    return rng.randint(0, 2**31, size=(n, e), dtype=np.int64).astype(np.int32)


def seeded_keys(gen_fn, n, batch, prf, seed=1234):
    """gen_fn(alpha, n, seed32, prf) -> (key_a, key_b).  Returns stacked keys and the indices."""
    r = random.Random(seed)
    ka, kb, idx = [], [], []
    for b in range(batch):
        alpha = r.randint(0, n - 1)
        a, bb = gen_fn(alpha, n, 1000 + b, prf)
        ka.append(a)
        kb.append(bb)
        idx.append(alpha)
    return np.stack(ka), np.stack(kb), np.array(idx)


def dot_u32(shares, table):
    """int32 share vectors [B,n] x int32 table [n,E] -> int32 [B,E], wrapping mod 2^32."""
    s = shares.astype(np.uint32).astype(np.uint64)
    t = table.astype(np.uint32).astype(np.uint64)
    out = np.zeros((s.shape[0], t.shape[1]), np.uint64)
    # chunk to keep the uint64 partial sums exact (each product < 2^64, so reduce mod 2^32 first)
    for e in range(t.shape[1]):
        prod = (s * t[:, e][None, :]) & np.uint64(0xFFFFFFFF)
        out[:, e] = prod.sum(axis=1) & np.uint64(0xFFFFFFFF)
    return out.astype(np.uint32).astype(np.int32)


def oracle_dot_mt(oracle, keys, prf, table, nthreads=16):
    """oracle.eval_dot over several host threads (ctypes releases the GIL): the tree-expansion
    oracle costs seconds per key at n >= 2^20, and the GPU box has cores to spare."""
    import threading
    keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 524)
    out = [None] * keys.shape[0]

    def work(t):
        for i in range(t, keys.shape[0], nthreads):
            out[i] = oracle.eval_dot(keys[i:i + 1], prf, table)[0]
    th = [threading.Thread(target=work, args=(t,)) for t in range(min(nthreads, keys.shape[0]))]
    [t.start() for t in th]
    [t.join() for t in th]
    return np.stack(out)
