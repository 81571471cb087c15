"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference CPU core.

Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Everything in the file comes out of oracle/_ref/libdpfref.so (dpf_base/dpf.h
behind oracle/ref_shim.cc): PRF known answers, keys from the reference key
generator, full share vectors from EvaluateFlat, and table inner products.  The
committed .npz is what travels; the GPU box has no reference tree.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402


def formula_table(n, e):
    """Closed-form table so fixtures do not depend on any RNG implementation."""
    i = np.arange(n, dtype=np.uint64).reshape(n, 1)
    c = np.arange(e, dtype=np.uint64).reshape(1, e)
    v = (i * np.uint64(2654435761) + c * np.uint64(40503) + (i * c) * np.uint64(97) + np.uint64(12345)) & np.uint64(0x7FFFFFFF)
    return v.astype(np.int64).astype(np.int32)


def main():
    ref = O.Ref()
    out = {}
    rng = np.random.RandomState(20260921)

    # --- PRF known answers -------------------------------------------------
    seeds = [0, 1, 0x0123456789abcdeffedcba9876543210, (1 << 128) - 1]
    seeds += [int.from_bytes(rng.bytes(16), "little") for _ in range(12)]
    kat_seed = np.zeros((len(seeds), 2), np.uint64)
    kat_out = np.zeros((4, len(seeds), 2, 2), np.uint64)   # [prf][seed][pos][lo,hi]
    for si, s in enumerate(seeds):
        kat_seed[si] = (s & (2**64 - 1), s >> 64)
        for prf in range(4):
            for pos in (0, 1):
                r = ref.prf(prf, s, pos)
                kat_out[prf, si, pos] = (r & (2**64 - 1), r >> 64)
    out["kat_seed"] = kat_seed
    out["kat_out"] = kat_out

    # --- keys, share vectors, inner products -------------------------------
    cases = []
    for prf in range(4):
        for n in (2, 128, 1024, 16384):
            for rep in range(2):
                alpha = int(rng.randint(0, n))
                seed32 = int(rng.randint(0, 2**31 - 1))
                cases.append((prf, n, alpha, seed32))
    meta = np.array(cases, np.int64)
    keys_a = np.zeros((len(cases), O.KEY_WORDS), np.int32)
    keys_b = np.zeros((len(cases), O.KEY_WORDS), np.int32)
    dots_a = np.zeros((len(cases), 16), np.int32)
    dots_b = np.zeros((len(cases), 16), np.int32)
    for ci, (prf, n, alpha, seed32) in enumerate(cases):
        ka, kb = ref.gen(alpha, n, seed32, prf)
        keys_a[ci], keys_b[ci] = ka, kb
        sa = ref.eval_full(ka, prf)
        sb = ref.eval_full(kb, prf)
        if n <= 1024:
            out["share_a_%d" % ci] = sa
            out["share_b_%d" % ci] = sb
        t = formula_table(n, 16).astype(np.uint32)
        dots_a[ci] = (sa.astype(np.uint32)[:, None] * t).sum(axis=0, dtype=np.uint32).astype(np.int32)
        dots_b[ci] = (sb.astype(np.uint32)[:, None] * t).sum(axis=0, dtype=np.uint32).astype(np.int32)
        # reference property: shares differ by beta=1 exactly at alpha
        d = (sa.astype(np.int64) - sb.astype(np.int64)) % (1 << 32)
        exp = np.zeros(n, np.int64)
        exp[alpha] = 1
        assert np.array_equal(d, exp), (prf, n, alpha)
    out["case_meta"] = meta            # columns: prf, n, alpha, seed32
    out["keys_a"] = keys_a
    out["keys_b"] = keys_b
    out["dots_a"] = dots_a             # vs formula_table(n, 16)
    out["dots_b"] = dots_b
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
