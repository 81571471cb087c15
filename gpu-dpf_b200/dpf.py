"""dpf.DPF -- the reference's Python API (dpf.py:35-137) on the B200-native engine.

Same class, method names, argument meaning and exceptions as the reference, so
`sample.py` and `benchmark.py` written against it keep working:

    d = DPF(prf=DPF.PRF_AES128)
    k1, k2 = d.gen(index, n)        # client: two int32[524] CPU key tensors
    d.eval_init(table)              # server: upload [n, entry_size] table
    shares = d.eval_gpu([k1, ...])  # server: int32 [len(keys), entry_size] on the CPU

What differs is behind the boundary: `dpf_cpp` here is the pybind shim over
include/b200dpf.h (hand-written sm_100a kernels), tables may have any entry
size, batches may have any length (no pad-to-512 round trip, no pad-to-16
columns), and one process can own an entry-range shard of the table
(`DPF(prf, device=d, shard=(rank, count))`) whose partial results add mod 2^32.

The module-level test_* / *_perf helpers mirror the reference's self-tests
(dpf.py:139-356): same names and defaults, same pass criteria.
"""
import os
import random
import sys
import time

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

try:
    import dpf_cpp
except ImportError as exc:  # fail loudly: there is no Python/CPU stand-in for the kernels
    raise ImportError(
        "dpf_cpp extension not built (run `python gpu-dpf_b200/build.py`); "
        "the DPF engine has no fallback implementation: %s" % exc)


class DPF(object):

    PRF_CHACHA20 = dpf_cpp.PRF_CHACHA20
    PRF_DUMMY = dpf_cpp.PRF_DUMMY
    PRF_SALSA20 = dpf_cpp.PRF_SALSA20
    PRF_AES128 = dpf_cpp.PRF_AES128

    ENTRY_SIZE = dpf_cpp.ENTRY_SIZE
    BATCH_SIZE = dpf_cpp.BATCH_SIZE

    DEFAULT_PRF = dpf_cpp.PRF_AES128

    _PRF_NAMES = {
        dpf_cpp.PRF_CHACHA20: "CHACHA20",
        dpf_cpp.PRF_DUMMY: "DUMMY",
        dpf_cpp.PRF_SALSA20: "SALSA20",
        dpf_cpp.PRF_AES128: "AES128",
    }

    def __init__(self, prf=None, device=0, shard=(0, 1), allow_non_pow2=False, devices=None, axis="auto"):
        # devices: None (one GPU: `device`), "all", or a list of device ids -- one process drives them all
        # (b200dpf_create_multi); axis "auto" | "entries" | "keys".  B200DPF_DEVICES / B200DPF_AXIS in the
        # environment set the defaults, so an unchanged `dpf.DPF()` caller (benchmark.py) can scale too.
        if devices is None and os.environ.get("B200DPF_DEVICES"):
            devices = os.environ["B200DPF_DEVICES"]
            axis = os.environ.get("B200DPF_AXIS", axis)
        if isinstance(devices, str):
            devices = dpf_cpp.parse_devices(devices)
        self.devices = list(devices) if devices is not None and len(devices) > 1 else None
        if devices is not None and len(devices) == 1:
            device = devices[0]
        self.axis = axis
        # allow_non_pow2: tables / domains whose size is not a power of two are padded with zero
        # rows up to the next one (a reference TODO, dpf.py:21; off by default so the reference's
        # "must be a power of two" errors stay as they are)
        self.allow_non_pow2 = allow_non_pow2
        self.buffers = None
        self.table_num_entries = None
        self.table_effective_entry_size = None
        self.table = None
        self.device = device
        self.shard = tuple(shard)
        self.prf_method = prf if prf is not None else self.DEFAULT_PRF
        self.prf_method_string = self._PRF_NAMES[self.prf_method]

    # ---- client -----------------------------------------------------------
    def gen(self, k, n, seed=None, secure=None):
        """Two keys for the point function at index k over a domain of n (dpf.py:63-74).

        By default (no seed given) every random word comes from a ChaCha20 DRBG keyed with fresh
        os.urandom entropy.  The reference seeds std::mt19937 with 32 bits of its seed
        (dpf_wrapper.cu:52; its own TODO, dpf.py:65), so a server could enumerate the 2^32 possible
        generator states of a key it holds and recover the index: that generator is kept only
        for callers that pass an explicit `seed` (deterministic keys, bit-identical to the
        reference's for the same seed: golden and parity tests) and do not ask for secure=True."""
        if secure is None:
            secure = seed is None
        if seed is None:
            seed = os.urandom(128)
        if n & (n - 1) != 0 and not self.allow_non_pow2:
            raise Exception("Table num entries (%d) must be a power of two" % (n))
        if k >= n:
            raise Exception("k (%d), the selected element, must be less than n (%d), the number of entries in the table"
                            % (k, n))
        n = _next_pow2(n)
        if secure:
            return dpf_cpp.gen_secure(k, n, seed, self.prf_method)
        return dpf_cpp.gen(k, n, seed, self.prf_method)

    def gen_batch(self, indices, n, seeds=None, nthreads=0):
        """Keys for many indices at once: two int32 [B, 524] tensors (multi-threaded keygen).
        seeds=None: ChaCha20-DRBG keys from os.urandom entropy (44 bytes per key); explicit integer
        seeds select the reference's 32-bit mt19937 generator (deterministic, for tests)."""
        if n & (n - 1) != 0:
            raise Exception("Table num entries (%d) must be a power of two" % (n))
        idx = torch.as_tensor(list(indices), dtype=torch.int64)
        if idx.numel() and int(idx.max()) >= n:
            raise Exception("k (%d), the selected element, must be less than n (%d), the number of entries in the table"
                            % (int(idx.max()), n))
        if seeds is None:
            return dpf_cpp.gen_batch_secure(idx, n, os.urandom(44 * idx.numel()), self.prf_method, nthreads)
        return dpf_cpp.gen_batch(idx, n, torch.as_tensor(seeds, dtype=torch.int64), self.prf_method, nthreads)

    # ---- server -----------------------------------------------------------
    def eval_cpu(self, keys, one_hot_only=False):
        """CPU evaluation (dpf.py:76-86): share vectors, or shares @ table."""
        if one_hot_only:
            return torch.stack([dpf_cpp.eval_cpu(k, self.prf_method) for k in keys])
        if self.table is None:
            raise Exception("Must call `eval_init` before `eval_cpu` with one_hot_only=False")
        one_hots = torch.stack([dpf_cpp.eval_cpu(k, self.prf_method) for k in keys])
        return torch.matmul(one_hots, self.table)

    def eval_init(self, table):
        """Upload the table (dpf.py:88-113).  Any entry size; no column padding."""
        if self.buffers is not None:
            dpf_cpp.eval_free(self.buffers)
            self.buffers = None
        if table.shape[0] < 128:
            raise Exception("Table (%d) must have at least 128 elements" % table.shape[0])
        if table.shape[0] & (table.shape[0] - 1) != 0:
            if not self.allow_non_pow2:
                raise Exception("Table num entries (%d) must be a power of two" % (table.shape[0]))
            table = _pad_rows_to_pow2(table)
        self.table = table
        self.table_num_entries = table.shape[0]
        self.table_effective_entry_size = table.shape[1]
        if self.devices:
            self.buffers = dpf_cpp.eval_init_multi(table, self.devices, {"auto": 0, "entries": 1, "keys": 2}[self.axis])
        else:
            self.buffers = dpf_cpp.eval_init_sharded(table, self.device, self.shard[0], self.shard[1])

    def eval_gpu(self, keys):
        """Evaluate a batch on the GPU (dpf.py:115-131): int32 [len(keys), entry_size] on the CPU.

        `keys` is a list of int32[524] tensors, or one int32 [B, 524] tensor."""
        if self.buffers is None:
            raise Exception("Must call `eval_init` before `eval_gpu`")
        if isinstance(keys, torch.Tensor):
            if keys.is_cuda:        # keys already on the device: no staging at all
                return self.eval_gpu_device(keys.contiguous()).cpu()
            return dpf_cpp.eval_gpu_packed(keys.contiguous(), self.buffers, self.prf_method)
        if len(keys) == 0:
            return torch.zeros((0, self.table_effective_entry_size), dtype=torch.int32)
        return dpf_cpp.eval_gpu_list(list(keys), self.buffers, self.prf_method)

    def pack_keys(self, keys):
        """Compact wire form of a batch (24% to 56% smaller than int32[524] keys, depending on n): one
        contiguous uint8 tensor [B, 32 + 64*depth], pinned when CUDA is available, for eval_gpu_compact."""
        blobs = [dpf_cpp.key_pack(k) for k in keys]
        t = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).reshape(len(blobs), -1)
        return t.pin_memory() if torch.cuda.is_available() else t

    def eval_gpu_compact(self, packed):
        """eval_gpu for keys in the compact wire form (pack_keys): the packed bytes are what crosses PCIe."""
        if self.buffers is None:
            raise Exception("Must call `eval_init` before `eval_gpu`")
        return dpf_cpp.eval_gpu_compact(packed.contiguous(), packed.shape[0], self.buffers, self.prf_method)

    def eval_gpu_device(self, keys_dev, out_dev=None, out_ptr=None, accumulate=False):
        """Device-resident, asynchronous variant: keys_dev int32 [B,524] CUDA tensor ->
        int32 [B, entry_size] CUDA tensor, enqueued on the current torch stream.
        accumulate=True adds into the destination instead of overwriting it; out_ptr may
        then be a raw (e.g. peer-mapped) device address."""
        if self.buffers is None:
            raise Exception("Must call `eval_init` before `eval_gpu`")
        if self.devices:
            raise Exception("a multi-device DPF takes host keys (eval_gpu); device-resident evaluation is per GPU "
                            "(DPF(device=d, shard=(rank, count)) or sharded.ShardedDPF)")
        if not (keys_dev.is_cuda and keys_dev.dtype == torch.int32 and keys_dev.is_contiguous() and keys_dev.dim() == 2
                and keys_dev.shape[1] == 524 and keys_dev.device.index == self.device):
            raise Exception("keys_dev must be a contiguous int32 [B, 524] tensor on cuda:%d" % self.device)
        stream = torch.cuda.current_stream(keys_dev.device).cuda_stream
        if out_ptr is None:
            if out_dev is None:
                out_dev = torch.empty((keys_dev.shape[0], self.table_effective_entry_size), dtype=torch.int32,
                                      device=keys_dev.device)
            elif not (out_dev.is_cuda and out_dev.dtype == torch.int32 and out_dev.is_contiguous()
                      and out_dev.device == keys_dev.device
                      and tuple(out_dev.shape) == (keys_dev.shape[0], self.table_effective_entry_size)):
                raise Exception("out_dev must be a contiguous int32 [%d, %d] tensor on %s"
                                % (keys_dev.shape[0], self.table_effective_entry_size, keys_dev.device))
            out_ptr = out_dev.data_ptr()
        dpf_cpp.eval_gpu_device(keys_dev.data_ptr(), keys_dev.shape[0], self.buffers, self.prf_method,
                                out_ptr, stream, accumulate)
        return out_dev

    def eval_gpu_pipelined(self, batches, depth=2):
        """Server loop over a stream of batches: yields one CPU
        int32 [B, entry_size] tensor per input batch, in order.  Each batch is an int32 [B, 524]
        CPU tensor (pinned memory makes the copies truly asynchronous).  The host-to-device copy of
        batch i+1 and the device-to-host copy of batch i-1 run on their own streams while batch i
        is being evaluated, so in steady state only the kernel time is exposed."""
        if self.buffers is None:
            raise Exception("Must call `eval_init` before `eval_gpu`")
        dev = torch.device("cuda", self.device)
        compute = torch.cuda.current_stream(dev)
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        inflight = []                      # (out_host, done_event)

        def drain(limit):
            while len(inflight) > limit:
                out_host, done = inflight.pop(0)
                done.synchronize()
                yield out_host

        for keys in batches:
            keys = keys.contiguous()
            with torch.cuda.stream(h2d):
                keys_dev = keys.to(dev, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(h2d)
            compute.wait_event(copied)
            keys_dev.record_stream(compute)
            out_dev = self.eval_gpu_device(keys_dev)
            computed = torch.cuda.Event()
            computed.record(compute)
            with torch.cuda.stream(d2h):
                d2h.wait_event(computed)
                out_dev.record_stream(d2h)
                out_host = torch.empty(out_dev.shape, dtype=torch.int32, pin_memory=True)
                out_host.copy_(out_dev, non_blocking=True)
                done = torch.cuda.Event()
                done.record(d2h)
            inflight.append((out_host, done))
            yield from drain(depth)
        yield from drain(0)

    def expand_gpu_device(self, keys_dev, out_dev=None):
        """Share vectors on the GPU (the eval_cpu(one_hot_only=True) quantity): int32 [B, n] CUDA tensor."""
        if self.buffers is None:
            raise Exception("Must call `eval_init` before `expand_gpu_device`")
        assert keys_dev.is_cuda and keys_dev.dtype == torch.int32 and keys_dev.is_contiguous()
        if out_dev is None:
            out_dev = torch.empty((keys_dev.shape[0], self.table_num_entries), dtype=torch.int32,
                                  device=keys_dev.device)
        stream = torch.cuda.current_stream(keys_dev.device).cuda_stream
        dpf_cpp.expand_gpu_device(keys_dev.data_ptr(), keys_dev.shape[0], self.buffers, self.prf_method,
                                  out_dev.data_ptr(), stream)
        return out_dev

    def close(self):
        if self.buffers is not None:
            dpf_cpp.eval_free(self.buffers)
            self.buffers = None

    def __del__(self):
        # sample.py's server() builds a DPF per call and drops it: release the device table with the object
        # (the reference never frees unless eval_init is called again, dpf.py:93-94)
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):
        if self.buffers is None:
            return "DPF(_uninitialized_, prf_method=%s)" % self.prf_method_string
        return "DPF(entries=%d, entry_size=%d, prf_method=%s)" % (
            self.table_num_entries, self.table_effective_entry_size, self.prf_method_string)


class BinnedDPF(DPF):
    """Batch PIR (the paper's co-design front end, paper/experimental/batch_pir/batch_pir_optimization.py:84-87,
    210-218, which the reference only costs analytically): the table is split into bins, a query carries one DPF
    key per bin it touches, and the server evaluates all (bin, key) pairs of a batch of queries in ONE launch.

        d = BinnedDPF(prf=DPF.PRF_AES128)
        d.eval_init([bin0, bin1, ...])           # tables [n_g, entry_size], power-of-two n_g >= 2
        k1, k2 = d.gen_in_bin(g, index)          # keys for entry `index` of bin g
        shares = d.eval_gpu(keys, bins)          # int32 [len(keys), entry_size]
    """

    def eval_init(self, tables):
        if self.buffers is not None:
            dpf_cpp.eval_free(self.buffers)
            self.buffers = None
        tables = list(tables)
        for t in tables:
            if t.shape[0] & (t.shape[0] - 1) != 0 or t.shape[0] < 2:
                raise Exception("Table num entries (%d) must be a power of two" % (t.shape[0]))
        self.bin_sizes = [int(t.shape[0]) for t in tables]
        self.table = None
        self.table_num_entries = sum(self.bin_sizes)
        self.table_effective_entry_size = tables[0].shape[1]
        self.buffers = dpf_cpp.group_init(tables, self.device)

    def gen_in_bin(self, g, k, seed=None, secure=False):
        return self.gen(k, self.bin_sizes[g], seed=seed, secure=secure)

    def eval_gpu(self, keys, bins):
        if self.buffers is None:
            raise Exception("Must call `eval_init` before `eval_gpu`")
        if not isinstance(keys, torch.Tensor):
            if len(keys) == 0:
                return torch.zeros((0, self.table_effective_entry_size), dtype=torch.int32)
            keys = torch.stack(list(keys))
        return dpf_cpp.group_eval(keys.contiguous(), torch.as_tensor(bins), self.buffers, self.prf_method)

    def __repr__(self):
        if self.buffers is None:
            return "BinnedDPF(_uninitialized_, prf_method=%s)" % self.prf_method_string
        return "BinnedDPF(bins=%d, entries=%d, entry_size=%d, prf_method=%s)" % (
            len(self.bin_sizes), self.table_num_entries, self.table_effective_entry_size, self.prf_method_string)


def _next_pow2(n):
    return 1 << max(1, (n - 1).bit_length())


def _pad_rows_to_pow2(table):
    """Zero rows up to the next power of two: they contribute nothing to any inner product."""
    pad = _next_pow2(table.shape[0]) - table.shape[0]
    return torch.cat([table, torch.zeros((pad, table.shape[1]), dtype=table.dtype, device=table.device)])


# ---------------------------------------------------------------------------
# self-tests and perf helpers, mirroring dpf.py:139-356 of the reference
# ---------------------------------------------------------------------------

def _gen_keys(dpf, batch, N):
    k1s, k2s, indices = [], [], []
    for _ in range(batch):
        indx = random.randint(0, N - 1)
        indices.append(indx)
        k1, k2 = dpf.gen(indx, N)
        k1s.append(k1)
        k2s.append(k2)
    return k1s, k2s, indices


def _index_table(N, cols=16, dtype=torch.int32):
    return (torch.arange(N).reshape(N, 1) * cols + torch.arange(cols).reshape(1, cols)).to(dtype)


def test_cpu_dpf_one_hot(N=1024):
    dpf = DPF()
    K = 42
    k1, k2 = dpf.gen(K, N)
    v1 = dpf.eval_cpu([k1], one_hot_only=True)
    v2 = dpf.eval_cpu([k2], one_hot_only=True)
    rec = (v1 - v2).numpy()
    gt = np.zeros(rec.shape)
    gt[:, K] = 1
    assert np.linalg.norm(rec - gt) <= 1e-8
    print("Pass CPU (one-hot only) check.")


def test_cpu_dpf(N=1024):
    dpf = DPF()
    k1s, k2s, gt_indices = _gen_keys(dpf, 64, N)
    dpf.table = _index_table(N)   # CPU-only: no device upload needed for eval_cpu
    a = dpf.eval_cpu(k1s)
    b = dpf.eval_cpu(k2s)
    rec = (a - b).numpy()
    gt = dpf.table[gt_indices, :].numpy()
    assert np.linalg.norm(rec - gt) <= 1e-8
    print("Pass CPU check.")


def test_gpu_dpf(N=8192):
    dpf = DPF()
    k1s, k2s, gt_indices = _gen_keys(dpf, 64, N)
    table = _index_table(N, dtype=torch.float32)   # the reference feeds a float table here
    dpf.eval_init(table)
    a = dpf.eval_gpu(k1s)
    b = dpf.eval_gpu(k2s)
    rec = (a - b).numpy()
    gt = table[gt_indices, :].numpy()
    assert np.linalg.norm(rec - gt) <= 1e-8
    print("Pass GPU check.")


def test_gpu_dpf_nopad(N=8192, batch=42, entrysize=13):
    dpf = DPF()
    k1s, k2s, gt_indices = _gen_keys(dpf, batch, N)
    table = torch.randint(2 ** 31, (N, entrysize)).int()
    dpf.eval_init(table)
    a = dpf.eval_gpu(k1s)
    b = dpf.eval_gpu(k2s)
    rec = (a - b).numpy()
    gt = table[gt_indices, :].numpy()
    assert np.linalg.norm(rec - gt) <= 1e-8
    print("Pass GPU (nopad) check.")


def test_gpu_dpf_sweep():
    for n in [128, 256, 512, 1024, 8192]:
        test_gpu_dpf_nopad(n, batch=random.randint(1, dpf_cpp.BATCH_SIZE * 5 - 1), entrysize=random.randint(1, 16 - 1))
    print("Pass GPU (sweep) check.")


def _perf(kind, N, batch, entrysize, prf):
    dpf = DPF(prf=prf)
    k1s, _, _ = _gen_keys(dpf, batch, N)
    table = torch.rand(N, entrysize).int()
    dpf.eval_init(table)
    fn = dpf.eval_gpu if kind == "gpu" else dpf.eval_cpu
    tstart = time.time()
    reps = 10
    for _ in range(reps):
        fn(k1s)
    elapsed = time.time() - tstart
    dpfs_per_sec = batch * reps / elapsed
    keysize = np.prod(k1s[0].shape) * 4
    print("%s Key Size: %d bytes, Perf: %d dpfs/sec" % (dpf, keysize, dpfs_per_sec))
    return dpfs_per_sec


def test_gpu_dpf_perf(N=2048, batch=dpf_cpp.BATCH_SIZE, entrysize=16, prf=DPF.DEFAULT_PRF):
    return _perf("gpu", N, batch, entrysize, prf)


def test_cpu_dpf_perf(N=2048, batch=dpf_cpp.BATCH_SIZE, entrysize=16, prf=DPF.DEFAULT_PRF):
    return _perf("cpu", N, batch, entrysize, prf)


if __name__ == "__main__":
    random.seed(time.time())
    test_cpu_dpf()
    test_cpu_dpf_one_hot()
    test_gpu_dpf()
    test_gpu_dpf_nopad()
    test_gpu_dpf_sweep()
    test_gpu_dpf_perf()
