"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 4) process groups over
gloo.  There is no GPU here, so a TEST subclass of ShardedDPF replaces the per-rank GPU
evaluation with the oracle's shard partial; what is under test is the product's shard index
math, process-group plumbing and the int32 wrapping reduce (ShardedDPF._reduce_to_rank0)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, prf, ret):
    for sub in ("gpu-dpf_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from common import random_table, seeded_keys
    from sharded import ShardedDPF, shard_indices
    orc = O.Oracle()
    table = random_table(n, 16, seed=3)
    ka, kb, idx = seeded_keys(orc.gen, n, 6, prf, seed=5)

    # the rows this rank would upload, in device order, must be exactly the oracle's shard
    rows = shard_indices(n, rank, world)
    depth = n.bit_length() - 1
    assert rows == [orc.bitrev(rank * (n // world) + q, depth) for q in range(n // world)]

    class OracleBackedShards(ShardedDPF):
        def eval_init(self, t):
            self.n, self.entry_size = t.shape
            return self

        def _evaluate_and_combine(self, packed):
            part = np.stack([orc.eval_dot_shard(k, prf, table, self.rank * (n // self.world), n // self.world)
                             for k in packed.numpy()])
            return self._reduce_to_rank0(torch.from_numpy(part))

    d = OracleBackedShards(prf=prf)
    d.eval_init(torch.from_numpy(table))
    got_a = d.eval_gpu([torch.from_numpy(k) for k in ka])
    got_b = d.eval_gpu(torch.from_numpy(kb))
    if rank == 0:
        assert np.array_equal(got_a.numpy(), orc.eval_dot(ka, prf, table))
        rec = (got_a.numpy().astype(np.uint32) - got_b.numpy().astype(np.uint32)).astype(np.int32)
        assert np.array_equal(rec, table[idx])
        ret.put("ok")
    else:
        assert got_a is None and got_b is None
    dist.barrier()
    dist.destroy_process_group()


def _worker_keys_axis(rank, world, port, ret):
    """axis="keys": batch slices, in-place gather into one [world*per, E] buffer on rank 0, buffers reused
    between calls -- the product's eval_gpu_device code path with the per-rank GPU evaluation replaced by the
    oracle (CPU tensors over gloo)."""
    for sub in ("gpu-dpf_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from common import random_table, seeded_keys
    from sharded import ShardedDPF, key_slice
    orc = O.Oracle()
    n, prf = 512, 1
    table = random_table(n, 16, seed=9)

    class OracleDPF:
        def eval_gpu_device(self, keys, out):
            out.copy_(torch.from_numpy(orc.eval_dot(keys.numpy(), prf, table)))
            return out

        def close(self):
            pass

    class KeysAxis(ShardedDPF):
        def eval_init(self, t):
            self.n, self.entry_size = t.shape
            if self.axis == "auto":
                self.axis = "keys" if self.n <= self.AUTO_KEYS_MAX_N else "entries"
            self._dpf = OracleDPF()
            return self

    d = KeysAxis(prf=prf, axis="auto")
    d.eval_init(torch.from_numpy(table))
    assert d.axis == "keys"
    for batch in (70, 70, 5, 64):                    # same size twice: the kept buffers are reused
        ka, _, _ = seeded_keys(orc.gen, n, batch, prf, seed=batch)
        got = d.eval_gpu_device(torch.from_numpy(ka))
        b, e = key_slice(batch, rank, world)
        assert 0 <= b <= e <= batch
        if rank == 0:
            assert np.array_equal(got.numpy(), orc.eval_dot(ka, prf, table)), batch
        else:
            assert got is None
    if rank == 0:
        ret.put("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_keys_axis_gather_over_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_keys_axis, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) == "ok"


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_reduce_over_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1024, 2, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) == "ok"


def test_shard_indices_partition():
    sys.path.insert(0, os.path.join(ROOT, "gpu-dpf_b200"))
    from sharded import shard_indices
    n = 256
    for world in (1, 2, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
        assert seen == list(range(n))
    # residue class of the LOW bits: the root of the tree consumes the index LSB
    assert all(i % 4 == 0b10 for i in shard_indices(n, 1, 4))
