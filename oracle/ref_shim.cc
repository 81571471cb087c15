/*
 * oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" doorway onto the UNMODIFIED reference CPU core.  This file
 * contains no DPF arithmetic of its own: it #includes the reference header
 * where it lies (REF_ROOT/dpf_base/dpf.h, given on the compiler command line
 * by oracle/Makefile) and forwards to its functions.  The build product goes
 * to oracle/_ref/libdpfref.so (git-ignored; travels to the GPU box with the
 * snapshot).  Nothing from the reference tree is copied into this repository.
 *
 * Used for: (1) pinning oracle/dpf_oracle.c, (2) generating tests/golden/,
 * (3) the `--impl reference` / cpu_baseline timing legs of bench.py.
 */
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "dpf_base/dpf.h"   /* resolved through -I$(REF_ROOT) */

namespace {

/* wire format of dpf_wrapper.cu:26-46, re-expressed over raw memory because
 * the wrapper itself needs torch to compile */
void pack_key(const SeedsCodewordsFlat &f, int n, int32_t *key)
{
    uint128_t *slots = reinterpret_cast<uint128_t *>(key);
    std::memset(key, 0, 524 * sizeof(int32_t));
    slots[0] = f.depth;
    std::memcpy(&slots[1], f.cw_1, sizeof(uint128_t) * 64);
    std::memcpy(&slots[65], f.cw_2, sizeof(uint128_t) * 64);
    slots[129] = f.last_keys[0];
    slots[130] = n;
}

void unpack_key(const int32_t *key, SeedsCodewordsFlat *f, int *n)
{
    const uint128_t *slots = reinterpret_cast<const uint128_t *>(key);
    f->depth = (int)slots[0];
    std::memcpy(f->cw_1, &slots[1], sizeof(uint128_t) * 64);
    std::memcpy(f->cw_2, &slots[65], sizeof(uint128_t) * 64);
    f->last_keys[0] = slots[129];
    *n = (int)slots[130];
}

}  // namespace

extern "C" {

void ref_prf(int prf, uint64_t seed_lo, uint64_t seed_hi, uint32_t pos,
             uint64_t *out_lo, uint64_t *out_hi)
{
    uint128_t seed = ((uint128_t)seed_hi << 64) | seed_lo;
    uint128_t r = PRF_SELECT(prf)(seed, pos);
    *out_lo = (uint64_t)r;
    *out_hi = (uint64_t)(r >> 64);
}

/* dpf_wrapper.cu:49-68 with the torch tensors replaced by int32[524] buffers */
int ref_gen(int64_t alpha, int64_t n, uint32_t seed32, int prf, int32_t *key_a, int32_t *key_b)
{
    std::mt19937 g(seed32);
    SeedsCodewords *s = GenerateSeedsAndCodewordsLog((int)alpha, 1, (int)n, g, prf);
    /* value-initialise so the slots the reference leaves untouched are zero */
    SeedsCodewordsFlat *a = new SeedsCodewordsFlat();
    SeedsCodewordsFlat *b = new SeedsCodewordsFlat();
    FlattenCodewords(s, 0, a);
    FlattenCodewords(s, 1, b);
    pack_key(*a, (int)n, key_a);
    pack_key(*b, (int)n, key_b);
    delete a;
    delete b;
    return 0;   /* the SeedsCodewords chain is leaked, as in the reference's own FreeSeedsCodewords */
}

void ref_eval_flat(const int32_t *key, int64_t idx, int prf, uint64_t *out_lo, uint64_t *out_hi)
{
    SeedsCodewordsFlat f;
    int n;
    unpack_key(key, &f, &n);
    uint128_t r = EvaluateFlat(&f, (int)idx, prf);
    *out_lo = (uint64_t)r;
    *out_hi = (uint64_t)(r >> 64);
}

/* dpf_wrapper.cu:70-84 without the per-element ATen indexing */
int ref_eval_full(const int32_t *key, int prf, int32_t *out_n)
{
    SeedsCodewordsFlat f;
    int n;
    unpack_key(key, &f, &n);
    for (int i = 0; i < n; i++) out_n[i] = (int)EvaluateFlat(&f, i, prf);
    return 0;
}

/*
 * CPU baseline: the reference's evaluation path (EvaluateFlat per index, then
 * the int32 inner product of dpf.py:85-86) for `nkeys` keys spread over
 * `nthreads` host threads, restricted to indices [idx_begin, idx_begin+idx_count)
 * so a bounded sample of a large table can be timed.  out is [nkeys][entry_size].
 */
int ref_eval_dot_mt(const int32_t *keys, int64_t nkeys, int prf, const int32_t *table,
                    int64_t n, int entry_size, int64_t idx_begin, int64_t idx_count,
                    int nthreads, int32_t *out)
{
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int t) {
        for (int64_t b = t; b < nkeys; b += nthreads) {
            SeedsCodewordsFlat f;
            int kn;
            unpack_key(keys + b * 524, &f, &kn);
            std::vector<uint32_t> acc((size_t)entry_size, 0u);
            for (int64_t i = idx_begin; i < idx_begin + idx_count; i++) {
                uint32_t share = (uint32_t)EvaluateFlat(&f, (int)i, prf);
                const int32_t *row = table + i * entry_size;
                for (int e = 0; e < entry_size; e++) acc[(size_t)e] += share * (uint32_t)row[e];
            }
            for (int e = 0; e < entry_size; e++) out[b * entry_size + e] = (int32_t)acc[(size_t)e];
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
    (void)n;
    return 0;
}

int ref_sizeof_flat(void) { return (int)sizeof(SeedsCodewordsFlat); }

}  // extern "C"
