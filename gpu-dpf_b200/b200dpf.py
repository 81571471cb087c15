"""ctypes binding of the C ABI in include/b200dpf.h.

This is the binding a non-torch host (or the reference's maintainers, see
INTEGRATION.md) would write: plain pointers and sizes over libb200dpf.so.  The
torch-facing module is `dpf_cpp` (csrc/dpf_cpp_ext.cpp); this one is used by the
parity tests and the benchmark, which want to drive the ABI directly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200dpf.so")
KEY_WORDS = 524
PRF_DUMMY, PRF_SALSA20, PRF_CHACHA20, PRF_AES128 = 0, 1, 2, 3
PRF_NAMES = {0: "DUMMY", 1: "SALSA20", 2: "CHACHA20", 3: "AES128"}

# every symbol include/b200dpf.h declares
SYMBOLS = [
    "b200dpf_version", "b200dpf_last_error", "b200dpf_gen", "b200dpf_gen_secure", "b200dpf_gen_batch", "b200dpf_gen_batch_secure", "b200dpf_gen_batch_gpu", "b200dpf_eval_cpu",
    "b200dpf_key_packed_size", "b200dpf_key_pack", "b200dpf_key_unpack", "b200dpf_key_n", "b200dpf_key_depth", "b200dpf_create", "b200dpf_destroy", "b200dpf_eval",
    "b200dpf_eval_packed", "b200dpf_eval_gather", "b200dpf_ctx_set_option", "b200dpf_create_multi", "b200dpf_ctx_device_count", "b200dpf_ctx_axis", "b200dpf_ctx_read_timing",
    "b200dpf_group_create", "b200dpf_group_eval", "b200dpf_group_bins", "b200dpf_ctx_last_device_ms",
    "b200dpf_host_staging", "b200dpf_eval_device", "b200dpf_eval_device_acc", "b200dpf_expand_device", "b200dpf_ctx_n", "b200dpf_ctx_entry_size",
    "b200dpf_ctx_device", "b200dpf_ctx_last_launches", "b200dpf_ctx_set_subtree_log2",
]

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


class B200DPFError(RuntimeError):
    pass


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libb200dpf.so not built (run `python gpu-dpf_b200/build.py`); there is no fallback")
    L = C.CDLL(LIB_PATH)
    L.b200dpf_version.restype = C.c_char_p
    L.b200dpf_last_error.restype = C.c_char_p
    L.b200dpf_gen.argtypes = [C.c_int64, C.c_int64, C.c_char_p, C.c_size_t, C.c_int, _i32p, _i32p]
    L.b200dpf_gen_secure.argtypes = [C.c_int64, C.c_int64, C.c_char_p, C.c_size_t, C.c_int, _i32p, _i32p]
    L.b200dpf_gen_batch.argtypes = [_i64p, _u32p, C.c_int64, C.c_int64, C.c_int, C.c_int, _i32p, _i32p]
    L.b200dpf_gen_batch_secure.argtypes = [_i64p, C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int, _i32p, _i32p]
    L.b200dpf_gen_batch_gpu.argtypes = [_i64p, C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.b200dpf_eval_cpu.argtypes = [_i32p, C.c_int, _i32p]
    L.b200dpf_key_packed_size.argtypes = [C.c_int]
    L.b200dpf_key_packed_size.restype = C.c_size_t
    L.b200dpf_key_pack.argtypes = [_i32p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.b200dpf_key_unpack.argtypes = [C.c_char_p, C.c_size_t, _i32p]
    L.b200dpf_key_n.argtypes = [_i32p]
    L.b200dpf_key_n.restype = C.c_int64
    L.b200dpf_key_depth.argtypes = [_i32p]
    L.b200dpf_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.b200dpf_create_multi.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
    L.b200dpf_ctx_device_count.argtypes = [C.c_void_p]
    L.b200dpf_ctx_axis.argtypes = [C.c_void_p]
    L.b200dpf_group_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i64p, C.c_int, C.c_int, C.c_int]
    L.b200dpf_group_eval.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int64, C.c_int, _i32p]
    L.b200dpf_group_bins.argtypes = [C.c_void_p]
    L.b200dpf_ctx_last_device_ms.argtypes = [C.c_void_p]
    L.b200dpf_ctx_last_device_ms.restype = C.c_double
    L.b200dpf_ctx_read_timing.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS"), C.c_int,
                                          C.POINTER(C.c_int)]
    L.b200dpf_destroy.argtypes = [C.c_void_p]
    L.b200dpf_eval.argtypes = [C.c_void_p, _i32p, C.c_int64, C.c_int, _i32p]
    L.b200dpf_eval_packed.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, _i32p]
    L.b200dpf_eval_gather.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64, C.c_int, _i32p]
    L.b200dpf_ctx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.b200dpf_host_staging.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.POINTER(C.c_int32))]
    L.b200dpf_eval_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.b200dpf_eval_device_acc.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.b200dpf_expand_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.b200dpf_ctx_n.argtypes = [C.c_void_p]
    L.b200dpf_ctx_n.restype = C.c_int64
    L.b200dpf_ctx_entry_size.argtypes = [C.c_void_p]
    L.b200dpf_ctx_device.argtypes = [C.c_void_p]
    L.b200dpf_ctx_last_launches.argtypes = [C.c_void_p]
    L.b200dpf_ctx_set_subtree_log2.argtypes = [C.c_void_p, C.c_int]
    return L


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def _check(rc, what):
    if rc != 0:
        raise B200DPFError("%s failed (%d): %s" % (what, rc, lib().b200dpf_last_error().decode()))


def gen(alpha, n, seed32, prf):
    a = np.zeros(KEY_WORDS, np.int32)
    b = np.zeros(KEY_WORDS, np.int32)
    seed = int(seed32 & 0xFFFFFFFF).to_bytes(4, "little")
    _check(lib().b200dpf_gen(alpha, n, seed, len(seed), prf, a, b), "b200dpf_gen")
    return a, b


def gen_secure(alpha, n, seed_bytes, prf):
    a = np.zeros(KEY_WORDS, np.int32)
    b = np.zeros(KEY_WORDS, np.int32)
    _check(lib().b200dpf_gen_secure(alpha, n, bytes(seed_bytes), len(seed_bytes), prf, a, b), "b200dpf_gen_secure")
    return a, b


def gen_batch(alphas, n, seeds32, prf, nthreads=0):
    alphas = np.ascontiguousarray(alphas, np.int64)
    seeds32 = np.ascontiguousarray(np.asarray(seeds32, np.int64) & 0xFFFFFFFF, np.uint32)
    a = np.zeros((len(alphas), KEY_WORDS), np.int32)
    b = np.zeros((len(alphas), KEY_WORDS), np.int32)
    _check(lib().b200dpf_gen_batch(alphas, seeds32, len(alphas), n, prf, nthreads, a, b), "b200dpf_gen_batch")
    return a, b


def gen_batch_secure(alphas, n, seed_bytes, prf, nthreads=0):
    """seed_bytes: 44 bytes of entropy per key, concatenated"""
    alphas = np.ascontiguousarray(alphas, np.int64)
    assert len(seed_bytes) == 44 * len(alphas)
    a = np.zeros((len(alphas), KEY_WORDS), np.int32)
    b = np.zeros((len(alphas), KEY_WORDS), np.int32)
    _check(lib().b200dpf_gen_batch_secure(alphas, bytes(seed_bytes), len(alphas), n, prf, nthreads, a, b), "b200dpf_gen_batch_secure")
    return a, b


def gen_batch_gpu(alphas, n, seed_bytes, prf, device=0, out_ptrs=None):
    """GPU keygen; same keys as gen_batch_secure for the same seeds.  out_ptrs=(ptr_a, ptr_b): device
    addresses to leave the keys at (int32 [count, 524] each) instead of returning host arrays."""
    alphas = np.ascontiguousarray(alphas, np.int64)
    assert len(seed_bytes) == 44 * len(alphas)
    if out_ptrs is not None:
        _check(lib().b200dpf_gen_batch_gpu(alphas, bytes(seed_bytes), len(alphas), n, prf, device,
                                           C.c_void_p(out_ptrs[0]), C.c_void_p(out_ptrs[1])), "b200dpf_gen_batch_gpu")
        return None
    a = np.zeros((len(alphas), KEY_WORDS), np.int32)
    b = np.zeros((len(alphas), KEY_WORDS), np.int32)
    _check(lib().b200dpf_gen_batch_gpu(alphas, bytes(seed_bytes), len(alphas), n, prf, device,
                                       a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)), "b200dpf_gen_batch_gpu")
    return a, b


def key_pack(key):
    key = np.ascontiguousarray(key, np.int32)
    buf = C.create_string_buffer(lib().b200dpf_key_packed_size(32))
    n = C.c_size_t()
    _check(lib().b200dpf_key_pack(key, buf, len(buf), C.byref(n)), "b200dpf_key_pack")
    return buf.raw[:n.value]


def key_unpack(packed):
    key = np.zeros(KEY_WORDS, np.int32)
    _check(lib().b200dpf_key_unpack(bytes(packed), len(packed), key), "b200dpf_key_unpack")
    return key


def eval_cpu(key, prf):
    key = np.ascontiguousarray(key, np.int32)
    n = lib().b200dpf_key_n(key)
    if n < 0:
        raise B200DPFError("malformed key")
    out = np.zeros(n, np.int32)
    _check(lib().b200dpf_eval_cpu(key, prf, out), "b200dpf_eval_cpu")
    return out


class Context:
    """One table (or one entry-range shard of it) resident on one GPU."""

    def __init__(self, table, device=0, shard_rank=0, shard_count=1):
        table = np.ascontiguousarray(table, np.int32)
        assert table.ndim == 2
        self.n, self.entry_size = table.shape
        self.handle = C.c_void_p()
        _check(lib().b200dpf_create(C.byref(self.handle), table.ctypes.data_as(C.c_void_p), self.n, self.entry_size,
                                    device, shard_rank, shard_count), "b200dpf_create")

    @classmethod
    def multi(cls, table, devices, axis=0):
        """One table over several GPUs of this process; axis 0 auto, 1 entries, 2 keys."""
        table = np.ascontiguousarray(table, np.int32)
        self = cls.__new__(cls)
        self.n, self.entry_size = table.shape
        self.handle = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        _check(lib().b200dpf_create_multi(C.byref(self.handle), table.ctypes.data_as(C.c_void_p), self.n, self.entry_size,
                                          devs, len(devices), axis), "b200dpf_create_multi")
        return self

    @property
    def device_count(self):
        return lib().b200dpf_ctx_device_count(self.handle)

    @property
    def axis(self):
        return lib().b200dpf_ctx_axis(self.handle)

    @classmethod
    def from_device_ptr(cls, ptr, n, entry_size, device=0, shard_rank=0, shard_count=1):
        self = cls.__new__(cls)
        self.n, self.entry_size = n, entry_size
        self.handle = C.c_void_p()
        _check(lib().b200dpf_create(C.byref(self.handle), C.c_void_p(ptr), n, entry_size, device, shard_rank,
                                    shard_count), "b200dpf_create")
        return self

    def eval(self, keys, prf):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, KEY_WORDS)
        out = np.zeros((keys.shape[0], self.entry_size), np.int32)
        _check(lib().b200dpf_eval(self.handle, keys, keys.shape[0], prf, out), "b200dpf_eval")
        return out

    def eval_packed(self, packed, nkeys, prf):
        """keys in the compact wire form (key_pack), concatenated"""
        out = np.zeros((nkeys, self.entry_size), np.int32)
        _check(lib().b200dpf_eval_packed(self.handle, bytes(packed), nkeys, prf, out), "b200dpf_eval_packed")
        return out

    def eval_gather(self, key_list, prf):
        """keys as a list of separate int32[524] arrays (the reference's calling convention)"""
        arrs = [np.ascontiguousarray(k, np.int32) for k in key_list]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        out = np.zeros((len(arrs), self.entry_size), np.int32)
        _check(lib().b200dpf_eval_gather(self.handle, ptrs, len(arrs), prf, out), "b200dpf_eval_gather")
        return out

    def set_option(self, name, value):
        _check(lib().b200dpf_ctx_set_option(self.handle, name.encode(), int(value)), "b200dpf_ctx_set_option")

    def read_timing(self, max_blocks=1024):
        """[blocks, 8] uint64 %globaltimer stamps of the last evaluation (option "timing" = 1)"""
        buf = np.zeros((max_blocks, 8), np.uint64)
        n = C.c_int()
        _check(lib().b200dpf_ctx_read_timing(self.handle, buf, max_blocks, C.byref(n)), "b200dpf_ctx_read_timing")
        return buf[:n.value]

    def eval_device(self, keys_ptr, nkeys, prf, out_ptr, stream=0, accumulate=False):
        fn = lib().b200dpf_eval_device_acc if accumulate else lib().b200dpf_eval_device
        _check(fn(self.handle, C.c_void_p(keys_ptr), nkeys, prf, C.c_void_p(out_ptr),
                                         C.c_void_p(stream)), "b200dpf_eval_device")

    def expand_device(self, keys_ptr, nkeys, prf, out_ptr, stream=0):
        _check(lib().b200dpf_expand_device(self.handle, C.c_void_p(keys_ptr), nkeys, prf, C.c_void_p(out_ptr),
                                           C.c_void_p(stream)), "b200dpf_expand_device")

    def set_subtree_log2(self, s):
        _check(lib().b200dpf_ctx_set_subtree_log2(self.handle, s), "b200dpf_ctx_set_subtree_log2")

    @property
    def last_device_ms(self):
        return lib().b200dpf_ctx_last_device_ms(self.handle)

    @property
    def last_launches(self):
        return lib().b200dpf_ctx_last_launches(self.handle)

    def close(self):
        if self.handle:
            lib().b200dpf_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupContext(Context):
    """Many tables ("bins") of one entry size behind one context: the batch-PIR front end
    (b200dpf_group_create / b200dpf_group_eval)."""

    def __init__(self, tables, device=0):
        self.tables = [np.ascontiguousarray(t, np.int32) for t in tables]
        assert all(t.ndim == 2 and t.shape[1] == self.tables[0].shape[1] for t in self.tables)
        self.entry_size = self.tables[0].shape[1]
        self.n = max(t.shape[0] for t in self.tables)
        sizes = np.array([t.shape[0] for t in self.tables], np.int64)
        ptrs = (C.c_void_p * len(self.tables))(*[t.ctypes.data for t in self.tables])
        self.handle = C.c_void_p()
        _check(lib().b200dpf_group_create(C.byref(self.handle), ptrs, sizes, len(self.tables), self.entry_size, device),
               "b200dpf_group_create")

    @property
    def nbins(self):
        return lib().b200dpf_group_bins(self.handle)

    def eval(self, keys, bins, prf):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, KEY_WORDS)
        bins = np.ascontiguousarray(bins, np.int32)
        assert bins.shape == (keys.shape[0],)
        out = np.zeros((keys.shape[0], self.entry_size), np.int32)
        _check(lib().b200dpf_group_eval(self.handle, keys, bins, keys.shape[0], prf, out), "b200dpf_group_eval")
        return out
