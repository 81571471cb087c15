/*
 * dpf_keygen.cu -- batched DPF key generation on the GPU ("optionally GPU").
 *
 * Key generation is the client's job in the protocol and stays a CPU function in the reference
 * (dpf_wrapper.cu:49-68 -> dpf_base/dpf.h:403-464); a benchmark or a load generator, however, needs
 * thousands of key pairs per batch (BASELINE config 5: 8192), and each costs ~6*depth PRF calls plus
 * its random draws.  One thread per key pair here; the construction, the draw order and the generator
 * (ChaCha20, RFC 8439 block function, keyed by 44 bytes of caller entropy per key) are exactly those of
 * b200dpf_gen_secure / host::gen_secure (dpf_host.cpp), so the two produce bit-identical keys for
 * identical seeds -- which is how tests/test_gpu_parity.py::test_gpu_keygen_matches_host checks it.
 *
 * Not a hot path: correction words live in local memory, AES goes through a plain 1 KiB T-table in
 * global memory.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "dpf_core.cuh"
#include "dpf_kernels.cuh"

namespace b200dpf {

namespace {

typedef unsigned __int128 u128;

struct AesGlobalTable {
    const uint32_t *te0;   /* 256 entries of Te0 in global memory */
    template <int K, int BYTE>
    __device__ __forceinline__ uint32_t te(uint32_t word) const
    {
        const uint32_t v = __ldg(te0 + ((word >> (8 * BYTE)) & 0xffu));
        return K == 0 ? v : __funnelshift_l(v, v, 8 * K);
    }
};

__device__ __forceinline__ u128 to_u128(const Seed &s)
{
    return ((u128)s.w << 96) | ((u128)s.z << 64) | ((u128)s.y << 32) | (u128)s.x;
}
__device__ __forceinline__ Seed from_u128(u128 v)
{
    return make_seed((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)(v >> 64), (uint32_t)(v >> 96));
}

template <int PRF>
__device__ __forceinline__ u128 prf128(const AesGlobalTable &ta, u128 seed, uint32_t pos)
{
    return to_u128(expand_one<PRF>(ta, from_u128(seed), pos));
}

/* ChaCha20 keystream as a deterministic random bit generator: same state layout and draw order as
 * host::ChaCha20Rng (dpf_host.cpp) */
struct ChaCha20Drbg {
    uint32_t state[16];
    uint32_t block[16];
    int used;

    __device__ void init(const uint8_t *key32, const uint8_t *nonce12)
    {
        state[0] = 0x61707865u; state[1] = 0x3320646eu; state[2] = 0x79622d32u; state[3] = 0x6b206574u;
        for (int i = 0; i < 8; i++)
            state[4 + i] = (uint32_t)key32[4 * i] | ((uint32_t)key32[4 * i + 1] << 8) | ((uint32_t)key32[4 * i + 2] << 16) |
                           ((uint32_t)key32[4 * i + 3] << 24);
        state[12] = 0;
        for (int i = 0; i < 3; i++)
            state[13 + i] = (uint32_t)nonce12[4 * i] | ((uint32_t)nonce12[4 * i + 1] << 8) |
                            ((uint32_t)nonce12[4 * i + 2] << 16) | ((uint32_t)nonce12[4 * i + 3] << 24);
        used = 16;
    }
    __device__ static void qr(uint32_t *x, int a, int b, int c, int d)
    {
        x[a] += x[b]; x[d] = __funnelshift_l(x[d] ^ x[a], x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = __funnelshift_l(x[b] ^ x[c], x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = __funnelshift_l(x[d] ^ x[a], x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = __funnelshift_l(x[b] ^ x[c], x[b] ^ x[c], 7);
    }
    __device__ void refill()
    {
        uint32_t x[16];
        for (int i = 0; i < 16; i++) x[i] = state[i];
        for (int r = 0; r < 10; r++) {
            qr(x, 0, 4, 8, 12); qr(x, 1, 5, 9, 13); qr(x, 2, 6, 10, 14); qr(x, 3, 7, 11, 15);
            qr(x, 0, 5, 10, 15); qr(x, 1, 6, 11, 12); qr(x, 2, 7, 8, 13); qr(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) block[i] = x[i] + state[i];
        state[12]++;
        used = 0;
    }
    __device__ uint32_t draw32()
    {
        if (used == 16) refill();
        return block[used++];
    }
    __device__ u128 draw128()
    {
        u128 v = 0;
        for (int i = 0; i < 4; i++) v |= (u128)draw32() << (32 * i);
        return v;
    }
    __device__ u128 draw128_odd()
    {
        u128 k = 0;
        while ((k & 1) == 0) k = draw128();
        return k;
    }
};

__device__ __forceinline__ void store_slot(uint4 *key, int slot, u128 v)
{
    key[slot] = make_uint4((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)(v >> 64), (uint32_t)(v >> 96));
}

/*
 * The log(n) construction of dpf_base/dpf.h:403-464 (levels above the base) over the N = 2 base of
 * dpf.h:290-360, written as the two loops the reference's recursion unrolls into: on the way DOWN
 * every level draws the odd payload of the level below; at the bottom the two root seeds and the
 * base correction words are made; on the way UP every level derives its correction words from the
 * PRF values of the two seeds handed up from below.
 */
template <int PRF>
__global__ void __launch_bounds__(64) keygen_kernel(const int64_t *__restrict__ alphas, const uint8_t *__restrict__ seeds44,
                                                    int64_t count, int depth, uint64_t n, const uint32_t *__restrict__ te0,
                                                    uint4 *__restrict__ keys_a, uint4 *__restrict__ keys_b)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    AesGlobalTable ta;
    ta.te0 = te0;
    const uint64_t alpha = (uint64_t)alphas[k];
    ChaCha20Drbg g;
    g.init(seeds44 + 44 * k, seeds44 + 44 * k + 32);

    u128 cw1[64], cw2[64], beta[33];
    for (int i = 0; i < 64; i++) cw1[i] = cw2[i] = 0;
    beta[0] = 1;                                       /* beta = 1, dpf_wrapper.cu:53 */
    for (int L = 0; L + 1 < depth; L++) beta[L + 1] = g.draw128_odd();

    u128 root_a, root_b, up_a, up_b;
    {
        /* N = 2 base: one seed per server, differing in the LSB (dpf.h:318-331) */
        const int L = depth - 1;
        const int row = (int)(alpha & 1u);
        u128 ka = g.draw128(), kb = g.draw128();
        ka &= ~(u128)1;
        kb = (kb & ~(u128)1) | 1;
        root_a = ka;
        root_b = kb;
        u128 delta[2];
        for (int i = 0; i < 2; i++) {
            delta[i] = prf128<PRF>(ta, ka, (uint32_t)i) - prf128<PRF>(ta, kb, (uint32_t)i);
            if (i == row) delta[i] -= beta[L];
        }
        for (int i = 0; i < 2; i++) {
            cw1[2 * L + i] = g.draw128();
            cw2[2 * L + i] = cw1[2 * L + i] + delta[i];
        }
        const u128 *bank_a = (ka & 1) ? cw2 : cw1;
        const u128 *bank_b = (kb & 1) ? cw2 : cw1;
        up_a = prf128<PRF>(ta, ka, (uint32_t)row) + bank_a[2 * L + row];
        up_b = prf128<PRF>(ta, kb, (uint32_t)row) + bank_b[2 * L + row];
    }
    for (int L = depth - 2; L >= 0; L--) {
        const int row = (int)((alpha >> (depth - 1 - L)) & 1u);
        const bool a_even = (up_a & 1) == 0;
        for (int i = 0; i < 2; i++) {
            u128 d = prf128<PRF>(ta, up_b, (uint32_t)i) - prf128<PRF>(ta, up_a, (uint32_t)i);
            if (a_even) d = (u128)0 - d;
            cw1[2 * L + i] = g.draw128();
            cw2[2 * L + i] = cw1[2 * L + i] + d;
            if (i == row) cw1[2 * L + i] += a_even ? beta[L] : (u128)0 - beta[L];
        }
        const u128 *bank_a = (up_a & 1) ? cw2 : cw1;
        const u128 *bank_b = (up_b & 1) ? cw2 : cw1;
        const u128 na = prf128<PRF>(ta, up_a, (uint32_t)row) + bank_a[2 * L + row];
        const u128 nb = prf128<PRF>(ta, up_b, (uint32_t)row) + bank_b[2 * L + row];
        up_a = na;
        up_b = nb;
    }

    /* the reference's wire format, dpf_wrapper.cu:26-35 */
    uint4 *ka = keys_a + k * 131, *kb = keys_b + k * 131;
    store_slot(ka, 0, (u128)depth);
    store_slot(kb, 0, (u128)depth);
    for (int i = 0; i < 64; i++) {
        store_slot(ka, 1 + i, cw1[i]);
        store_slot(kb, 1 + i, cw1[i]);
        store_slot(ka, 65 + i, cw2[i]);
        store_slot(kb, 65 + i, cw2[i]);
    }
    store_slot(ka, 129, root_a);
    store_slot(kb, 129, root_b);
    store_slot(ka, 130, (u128)n);
    store_slot(kb, 130, (u128)n);
}

}  // namespace

cudaError_t launch_keygen(int prf, const int64_t *alphas, const uint8_t *seeds44, int64_t count, int depth, uint64_t n,
                          const uint32_t *te0, uint4 *keys_a, uint4 *keys_b, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    const int threads = 64;
    const int grid = (int)((count + threads - 1) / threads);
    switch (prf) {
    case PRF_DUMMY: keygen_kernel<PRF_DUMMY><<<grid, threads, 0, stream>>>(alphas, seeds44, count, depth, n, te0, keys_a, keys_b); break;
    case PRF_SALSA20: keygen_kernel<PRF_SALSA20><<<grid, threads, 0, stream>>>(alphas, seeds44, count, depth, n, te0, keys_a, keys_b); break;
    case PRF_CHACHA20: keygen_kernel<PRF_CHACHA20><<<grid, threads, 0, stream>>>(alphas, seeds44, count, depth, n, te0, keys_a, keys_b); break;
    case PRF_AES128: keygen_kernel<PRF_AES128><<<grid, threads, 0, stream>>>(alphas, seeds44, count, depth, n, te0, keys_a, keys_b); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace b200dpf
