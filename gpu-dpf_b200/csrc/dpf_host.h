/*
 * dpf_host.h -- host-side (CPU) parts of the engine's API: the key wire format,
 * client key generation and the API's CPU evaluation function.  These mirror
 * CPU-only entry points of the reference extension (dpf_wrapper.cu:26-84); the
 * GPU path never routes through them.
 */
#pragma once

#include <stdint.h>

#include "dpf_core.cuh"

namespace b200dpf {
namespace host {

/* wire format, dpf_wrapper.cu:26-46: 131 little-endian 128-bit slots */
enum : int { KEY_WORDS = 524, SLOT_DEPTH = 0, SLOT_CW1 = 1, SLOT_CW2 = 65, SLOT_ROOT = 129, SLOT_N = 130 };

/* Te0 in the little-endian column convention of dpf_core.cuh (256 words),
 * built once from a computed S-box. */
const uint32_t *aes_te0();
/* The AES S-box (256 bytes). */
const uint8_t *aes_sbox();

Seed prf(int prf_id, const Seed &s, uint32_t pos);

inline Seed key_slot(const int32_t *key, int slot)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(key) + 4 * slot;
    return make_seed(p[0], p[1], p[2], p[3]);
}

/* depth / n sanity: returns depth (1..32) or -1 */
int key_depth(const int32_t *key);
int64_t key_n(const int32_t *key);

int gen(int64_t alpha, int64_t n, uint32_t seed32, int prf_id, int32_t *key_a, int32_t *key_b);
/* ChaCha20-DRBG variant: key = seed[0..31], nonce = seed[32..43] */
int gen_secure(int64_t alpha, int64_t n, const uint8_t seed[44], int prf_id, int32_t *key_a, int32_t *key_b);
int eval_cpu(const int32_t *key, int prf_id, int32_t *out_n);

}  // namespace host
}  // namespace b200dpf
