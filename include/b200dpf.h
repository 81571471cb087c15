/*
 * b200dpf.h -- C ABI of the B200-native DPF evaluation engine.
 *
 * This is the drop-in boundary for the ONE hot path of facebookresearch/GPU-DPF:
 * batched full-domain evaluation of log(n)-key distributed point functions
 * fused with the int32 inner product against the server's table.  Everything a
 * binding of that path needs is here as plain pointers and sizes (no torch, no
 * C++ types).  The reference reaches the same functionality through the five
 * pybind functions of dpf_wrapper.cu:188-204; each entry point below names the
 * reference interface it stands in for.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200DPF_E* code on
 *     failure; b200dpf_last_error() returns a thread-local message.
 *   - a key is the reference's wire format: int32[524] == 131 little-endian
 *     128-bit slots, [0]=depth, [1..64]=cw_1, [65..128]=cw_2, [129]=root seed,
 *     [130]=n                                  (dpf_wrapper.cu:26-46).
 *   - PRF ids are the reference's: 0 DUMMY, 1 SALSA20, 2 CHACHA20, 3 AES128
 *                                              (dpf_wrapper.cu:200-203).
 *   - results: out[b][e] = int32( sum_i low32(EvaluateFlat(key_b, i)) *
 *     uint32(table[i][e]) mod 2^32 )           (dpf_base/dpf.h:362-377 +
 *     dpf_wrapper.cu:178-185 + dpf.py:85-86).
 *
 * The GPU entry points require a CUDA device and fail with B200DPF_ECUDA when
 * none is usable: there is no CPU fallback behind them.
 *
 * Threading: a context is NOT re-entrant -- its device scratch (ticket counters, frontier,
 * leaf cache, key/result staging) serves one evaluation at a time, so calls on one context must
 * come from one host thread at a time.  Evaluations enqueued on DIFFERENT streams through the
 * *_device entry points are ordered by the library (each waits for the context's previous
 * evaluation).  Use one context per concurrent stream of work; contexts are independent.
 */
#ifndef B200DPF_H
#define B200DPF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DPF_KEY_WORDS 524          /* int32 words per key (2096 bytes)      */
#define B200DPF_DEFAULT_ENTRY_SIZE 16  /* dpf_wrapper.cu:18  (MM)               */
#define B200DPF_DEFAULT_BATCH_SIZE 512 /* dpf_wrapper.cu:21  (BATCH_SIZE)       */

enum {
    B200DPF_PRF_DUMMY = 0,
    B200DPF_PRF_SALSA20 = 1,
    B200DPF_PRF_CHACHA20 = 2,
    B200DPF_PRF_AES128 = 3
};

enum {
    B200DPF_OK = 0,
    B200DPF_EINVAL = -1,   /* bad argument (shape, power of two, prf id, ...)   */
    B200DPF_ECUDA = -2,    /* CUDA runtime error or no usable device            */
    B200DPF_ENOMEM = -3,
    B200DPF_ESTATE = -4    /* call out of order (e.g. eval before table upload) */
};

typedef struct b200dpf_ctx b200dpf_ctx;

/* Library identification: "b200dpf <semver> sm_100a". */
const char *b200dpf_version(void);

/* Thread-local description of the last failure on this thread. */
const char *b200dpf_last_error(void);

/* ------------------------------------------------------------------------- */
/* Client side (CPU).                                                         */
/* ------------------------------------------------------------------------- */

/*
 * Two-server key generation for the point function f(alpha) = 1.
 * Replaces: dpf_cpp.gen -> gen()                          dpf_wrapper.cu:49-68
 *           (GenerateSeedsAndCodewordsLog dpf_base/dpf.h:403-464, FlattenCodewords :239-270).
 * `seed`/`seed_len`: the caller's entropy; like the reference, the generator is
 * std::mt19937 seeded with the first 4 bytes (little endian), so identical
 * seeds give keys identical to the reference's.
 * key_a/key_b: caller-allocated int32[524] each.
 */
int b200dpf_gen(int64_t alpha, int64_t n, const uint8_t *seed, size_t seed_len,
                int prf, int32_t *key_a, int32_t *key_b);

/*
 * Key generation from a cryptographic generator (the
 * reference's own TODO at dpf.py:65 "replace with secure 128-bit RNG").  Same DPF
 * construction and wire format, but every random draw -- including the upper-level
 * correction words, which the reference draws as 32-bit values (dpf_base/dpf.h:450) --
 * is 128 bits of a ChaCha20 (RFC 8439) keystream keyed with seed[0..31] and nonce
 * seed[32..43].  seed_len must be >= 44 bytes of caller entropy.  Keys differ from
 * b200dpf_gen's for the same seed; they evaluate with the same entry points.
 */
int b200dpf_gen_secure(int64_t alpha, int64_t n, const uint8_t *seed, size_t seed_len,
                       int prf, int32_t *key_a, int32_t *key_b);

/*
 * Batched key generation: `count` independent
 * b200dpf_gen calls spread over `nthreads` host threads (0 = all cores).
 * alphas[count]; seeds32[count] (one 32-bit generator seed per key);
 * keys_a/keys_b: int32[count][524].
 */
int b200dpf_gen_batch(const int64_t *alphas, const uint32_t *seeds32, int64_t count,
                      int64_t n, int prf, int nthreads,
                      int32_t *keys_a, int32_t *keys_b);

/* Batched b200dpf_gen_secure: seeds44 = count x 44 bytes of caller entropy (key || nonce per key). */
int b200dpf_gen_batch_secure(const int64_t *alphas, const uint8_t *seeds44, int64_t count,
                             int64_t n, int prf, int nthreads,
                             int32_t *keys_a, int32_t *keys_b);

/*
 * The same batched keygen on a GPU (one thread per key pair; SURVEY.md section 8(f) rank 1, "optionally
 * GPU"): bit-identical to b200dpf_gen_batch_secure for identical seeds.  alphas and seeds44 are host
 * arrays; keys_a / keys_b may be host or device memory ([count][524] int32), so a load generator can
 * leave the keys on the device and hand them to b200dpf_eval_device.  Key generation is the CLIENT's
 * step of the protocol; this entry point exists for benchmarks and load generators.
 */
int b200dpf_gen_batch_gpu(const int64_t *alphas, const uint8_t *seeds44, int64_t count, int64_t n, int prf,
                          int device, int32_t *keys_a, int32_t *keys_b);

/*
 * One server's share vector on the CPU, natural index order: out[i] =
 * (int32) low32(EvaluateFlat(key, i)), i < n (n read from the key).
 * Replaces: dpf_cpp.eval_cpu -> eval_dpf_cpu()            dpf_wrapper.cu:70-84.
 * This is the API's CPU function, not a fallback for the GPU path.
 */
int b200dpf_eval_cpu(const int32_t *key, int prf, int32_t *out_n);

/*
 * Compact wire form of a key.  Of the 131 slots of the
 * reference format only 4*depth correction words, the root seed and the depth are live
 * (dpf_base/dpf.h:18-29 sizes the arrays for n = 2^32).  Packed layout, little endian,
 * every field 16-byte aligned so the GPU reads it in place (b200dpf_eval_packed):
 *   "DPF2" | depth (u8) | 11 zero bytes | root seed (16) | per level L = 0..depth-1:
 *   cw_1[2L], cw_1[2L+1], cw_2[2L], cw_2[2L+1] (16 bytes each)
 * = 32 + 64*depth bytes (1312 at n = 2^20, 928 at n = 2^14, instead of 2096).  unpack()
 * restores the exact int32[524] key (unused slots zero), so packed keys evaluate
 * bit-identically.
 */
size_t b200dpf_key_packed_size(int depth);
int b200dpf_key_pack(const int32_t *key, uint8_t *out, size_t out_cap, size_t *written);
int b200dpf_key_unpack(const uint8_t *in, size_t in_len, int32_t *key);

/* n stored in a key (slot 130) and its depth (slot 0); -1 if malformed. */
int64_t b200dpf_key_n(const int32_t *key);
int b200dpf_key_depth(const int32_t *key);

/* ------------------------------------------------------------------------- */
/* Server side (GPU, sm_100a).                                                */
/* ------------------------------------------------------------------------- */

/*
 * Upload a table and build an evaluation context on CUDA device `device`.
 * Replaces: dpf_cpp.eval_init -> eval_init()              dpf_wrapper.cu:93-132
 *           (+ dpf_hybrid_initialize dpf_gpu/dpf/dpf_hybrid.cu:30-36).
 *
 * table:        int32 [n][entry_size] row-major, natural index order, host or
 *               device memory (any pointer cudaMemcpy can read).
 * n:            power of two, 2 <= n <= 2^31.
 * entry_size:   1 .. 4096 int32 columns (the reference fixes 16).
 * shard_rank / shard_count: entry-range sharding.
 *               shard_count must be a power of two <= n/2; this context then owns
 *               the GGM subtree under the depth-log2(shard_count) node with
 *               breadth-first index shard_rank, copies only those n/shard_count
 *               rows, and b200dpf_eval* returns PARTIAL sums that the caller
 *               adds (mod 2^32) across shards.  Pass (0, 1) for a whole table.
 */
int b200dpf_create(b200dpf_ctx **ctx, const int32_t *table, int64_t n, int entry_size,
                   int device, int shard_rank, int shard_count);

/*
 * One table on SEVERAL GPUs of this process (the reference is
 * single-GPU: device 0 implicitly, dpf_wrapper.cu:114-129).  The returned context is used with
 * b200dpf_eval / b200dpf_eval_packed / b200dpf_destroy exactly like a single-device one, so a
 * caller of the reference API scales without a process launcher.
 *   axis B200DPF_AXIS_ENTRIES  device d owns entry-range shard (d, ndev) (ndev a power of two):
 *                              every device evaluates every key over its subtree and device
 *                              devices[0] adds the [nkeys][entry_size] partials with a kernel that
 *                              loads the peers' buffers over NVLink.
 *        B200DPF_AXIS_KEYS     every device holds the whole table and evaluates a contiguous
 *                              slice of the batch; nothing crosses GPUs.
 *        B200DPF_AXIS_AUTO     KEYS for n <= B200DPF_AUTO_KEYS_MAX_N (a shard of a small table
 *                              cannot keep a GPU busy), else ENTRIES.
 * One host thread per extra device issues that device's copies and launches.
 */
enum { B200DPF_AXIS_AUTO = 0, B200DPF_AXIS_ENTRIES = 1, B200DPF_AXIS_KEYS = 2 };
#define B200DPF_AUTO_KEYS_MAX_N (1 << 18)
int b200dpf_create_multi(b200dpf_ctx **ctx, const int32_t *table, int64_t n, int entry_size,
                         const int *devices, int ndev, int axis);

/* Devices behind a context (1 unless it came from b200dpf_create_multi) and its axis. */
int b200dpf_ctx_device_count(const b200dpf_ctx *ctx);
int b200dpf_ctx_axis(const b200dpf_ctx *ctx);

/*
 * Grouped evaluation -- the batch-PIR front end (the reference only
 * models it analytically: paper/experimental/batch_pir/batch_pir_optimization.py:84-87 charges one
 * DPF over every bin per query, :210-218 sums them).  One context holds `nbins` tables ("bins") of
 * possibly different power-of-two sizes n[g] (same entry_size); b200dpf_group_eval evaluates a mixed
 * list of (bin, key) pairs -- key b against table bins[b] -- in ONE launch: keys are sorted by bin,
 * cut into groups of <= 32 keys of one bin, and the work items of all groups form one ticket space,
 * so many small bins fill the GPU the way one big table does.
 *   out[b][e] = the same inner product b200dpf_eval would give for key b on table bins[b].
 * tables[g]: int32 [n[g]][entry_size], host or device.  keys: int32[nkeys][524]; bins: int32[nkeys].
 */
int b200dpf_group_create(b200dpf_ctx **ctx, const int32_t *const *tables, const int64_t *n, int nbins,
                         int entry_size, int device);
int b200dpf_group_eval(b200dpf_ctx *ctx, const int32_t *keys, const int32_t *bins, int64_t nkeys, int prf,
                       int32_t *out);
int b200dpf_group_bins(const b200dpf_ctx *ctx);

/* Replaces: dpf_cpp.eval_free -> eval_free()               dpf_wrapper.cu:86-91. */
int b200dpf_destroy(b200dpf_ctx *ctx);

/*
 * Evaluate `nkeys` keys against the context's table; host buffers in and out.
 * Replaces: dpf_cpp.eval_gpu -> eval_gpu()                dpf_wrapper.cu:134-186
 *           (+ dpf_hybrid<prf>() dpf_gpu/dpf/dpf_hybrid.cu:258-272).
 * keys: int32[nkeys][524] (host).  out: int32[nkeys][entry_size] (host).
 * Synchronous: the result is complete on return.  Any nkeys >= 1 (the
 * reference requires exactly 512).
 */
int b200dpf_eval(b200dpf_ctx *ctx, const int32_t *keys, int64_t nkeys, int prf, int32_t *out);

/*
 * Same evaluation with the batch given as nkeys POINTERS to int32[524] keys -- the shape the
 * reference's eval_gpu receives (a vector of 512 key tensors, dpf_wrapper.cu:134-146).  Replaces
 * its per-key marshalling loop + blocking upload (dpf_wrapper.cu:137-150): the live parts of each
 * key are gathered into pinned staging in the compact layout and uploaded chunk by chunk while the
 * host gathers the next chunk.
 */
int b200dpf_eval_gather(b200dpf_ctx *ctx, const int32_t *const *keys, int64_t nkeys, int prf, int32_t *out);

/*
 * Same evaluation with keys in the compact wire form (b200dpf_key_pack): `packed` holds nkeys
 * keys of this table's depth back to back, b200dpf_key_packed_size(depth) bytes each (host
 * memory).  The buffer goes to the device as it is and the kernel reads the packed layout, so the
 * host-to-device copy is 2.3x smaller at n = 2^14 and 1.6x at n = 2^20 than with the
 * reference's 2096-byte keys (dpf_wrapper.cu:26-46, :150).
 */
int b200dpf_eval_packed(b200dpf_ctx *ctx, const uint8_t *packed, int64_t nkeys, int prf, int32_t *out);

/*
 * Pinned (page-locked) host staging owned by the context, sized for `nkeys` keys: callers
 * that assemble a batch from scattered keys (the reference passes 512 separate tensors,
 * dpf_wrapper.cu:137-146) can pack straight into it and hand the same pointer to
 * b200dpf_eval, which then copies to the device without an intermediate staging copy.
 * The buffer stays valid until the next b200dpf_host_staging / b200dpf_destroy.
 */
int b200dpf_host_staging(b200dpf_ctx *ctx, int64_t nkeys, int32_t **keys_pinned);

/*
 * Same computation with DEVICE buffers, enqueued on `cuda_stream` (a
 * cudaStream_t, may be NULL for the default stream) and NOT synchronised:
 * for callers that keep keys and results resident (benchmarks, NCCL reduce of
 * sharded partials).  keys_dev: int32[nkeys][524]; out_dev: int32[nkeys][entry_size].
 */
int b200dpf_eval_device(b200dpf_ctx *ctx, const void *keys_dev, int64_t nkeys, int prf,
                        void *out_dev, void *cuda_stream);

/*
 * As b200dpf_eval_device, but ADDS the results into out_dev (mod 2^32) instead
 * of overwriting it: out_dev is not cleared first.  With out_dev pointing at a
 * peer GPU's buffer (NVLink peer mapping / symmetric memory) the kernel's
 * red.global.add.u32 epilogue performs the cross-shard reduction itself, so
 * sharded evaluation needs no separate collective (SURVEY.md section 8(e),
 * "fused alternative").  The caller orders the clearing of the destination
 * before, and its consumption after, all contributing launches.
 */
int b200dpf_eval_device_acc(b200dpf_ctx *ctx, const void *keys_dev, int64_t nkeys, int prf,
                            void *out_dev, void *cuda_stream);

/*
 * Non-fused full-domain expansion on the GPU (SURVEY.md section 8(f) rank 2;
 * the role of FUSES_MATMUL=0 in dpf_gpu/dpf/dpf_hybrid.cu:162-165, but in
 * natural index order): shares_dev[b][i] = low32(EvaluateFlat(key_b, i)) for
 * this context's shard rows... whole domain only (shard_count must be 1).
 * keys_dev: int32[nkeys][524]; shares_dev: int32[nkeys][n].
 */
int b200dpf_expand_device(b200dpf_ctx *ctx, const void *keys_dev, int64_t nkeys, int prf,
                          void *shares_dev, void *cuda_stream);

/* Introspection. */
int64_t b200dpf_ctx_n(const b200dpf_ctx *ctx);
int b200dpf_ctx_entry_size(const b200dpf_ctx *ctx);
int b200dpf_ctx_device(const b200dpf_ctx *ctx);

/* Device time in milliseconds of the kernels of the most recent b200dpf_eval / _eval_packed /
 * _eval_gather / _group_eval on this context (CUDA events around the launches, copies excluded);
 * negative if none. */
double b200dpf_ctx_last_device_ms(b200dpf_ctx *ctx);

/* Kernels launched by the most recent b200dpf_eval / _eval_device / _expand_device
 * call on this context (for benchmark accounting). */
int b200dpf_ctx_last_launches(const b200dpf_ctx *ctx);

/* Algorithm tuning knob, mainly for tests: log2 of the leaves one thread
 * expands depth-first per work item (0 = automatic). */
int b200dpf_ctx_set_subtree_log2(b200dpf_ctx *ctx, int s);

/*
 * Tuning knobs of a live context.  Their defaults come from the environment variables named
 * below, which are read ONCE, in b200dpf_create.
 *   "one_launch"    (B200DPF_ONE_LAUNCH, 1)     whole evaluation as one cooperative launch
 *   "frontier"      (B200DPF_FRONTIER, 1)       expand the top of the tree once per evaluation
 *   "frontier_mb"   (B200DPF_FRONTIER_MB, 256)  cap on the frontier buffer
 *   "subtree_log2"  (B200DPF_S, 0 = automatic)  as b200dpf_ctx_set_subtree_log2
 *   "lane_split"    (B200DPF_LANE_SPLIT, 1)     batches < 32 keys: lane = (key, subtree)
 *   "leaf_cache"    (B200DPF_LEAF_CACHE, 1)     entry_size > 32: expand once, MAC-only passes
 *   "leaf_cache_mb" (B200DPF_LEAF_CACHE_MB, 16384)  cap; larger batches run in chunks that fit
 *   "mac_tma"       (B200DPF_MAC_TMA, 1)        cp.async.bulk-staged MAC passes
 *   "balance_top"   (B200DPF_BALANCE_TOP, 1)    even per-block shares of the tree-top phase
 *   "top_log2"      (B200DPF_TOP_LOG2, 0 = estimate)  size of a tree-top item (log2 frontier nodes)
 *   "timing"        (B200DPF_TIMING, 0)         record per-block phase time stamps (diagnostics)
 * Results never depend on them.
 */
int b200dpf_ctx_set_option(b200dpf_ctx *ctx, const char *name, int value);

/*
 * Diagnostics ("timing" option): %globaltimer nanosecond stamps of the last single-launch
 * evaluation, 8 uint64 per block: [0] kernel start, [1] tables + clears done, [2] tree-top
 * phase done, [3] grid barrier passed, [4] main phase done, rest zero.  Synchronises the device.
 */
int b200dpf_ctx_read_timing(b200dpf_ctx *ctx, uint64_t *stamps, int max_blocks, int *nblocks);

#ifdef __cplusplus
}
#endif
#endif /* B200DPF_H */
