"""Entry-range sharding across real GPUs: one process per GPU, NCCL reduce of the partials.
Needs >= 2 CUDA devices (skipped on a 1-GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    for sub in ("gpu-dpf_b200", "oracle", "tests"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import b200dpf
    import oracle as O
    from common import random_table, seeded_keys
    from sharded import ShardedDPF
    orc = O.Oracle()
    n = 1 << 14
    table = random_table(n, 16, seed=13)
    for prf, axis, reduce in ((3, "entries", "nccl"), (2, "entries", "nccl"), (3, "entries", "fused"),
                              (1, "entries", "fused"), (3, "keys", "nccl")):
        ka, kb, idx = seeded_keys(b200dpf.gen, n, 70, prf, seed=17)
        d = ShardedDPF(prf=prf, axis=axis, reduce=reduce)
        d.eval_init(torch.from_numpy(table))
        got_a = d.eval_gpu(torch.from_numpy(ka))
        got_b = d.eval_gpu([torch.from_numpy(k) for k in kb])
        if rank == 0:
            assert np.array_equal(got_a.numpy()[:8], orc.eval_dot(ka[:8], prf, table))
            rec = (got_a.numpy().astype(np.uint32) - got_b.numpy().astype(np.uint32)).astype(np.int32)
            assert np.array_equal(rec, table[idx])
        else:
            assert got_a is None
        d.close()
    # BASELINE config 5 shape (scaled down): entry_size 128, AES128, sharded, batch not a multiple of 32
    n, entry, prf = 1 << 16, 128, 3
    table = random_table(n, entry, seed=5)
    ka, kb, idx = seeded_keys(b200dpf.gen, n, 45, prf, seed=6)
    d = ShardedDPF(prf=prf)
    d.eval_init(torch.from_numpy(table))
    got_a, got_b = d.eval_gpu(torch.from_numpy(ka)), d.eval_gpu(torch.from_numpy(kb))
    if rank == 0:
        rec = (got_a.numpy().astype(np.uint32) - got_b.numpy().astype(np.uint32)).astype(np.int32)
        assert np.array_equal(rec, table[idx])
        assert np.array_equal(got_a.numpy()[:2], orc.eval_dot(ka[:2], prf, table))
    d.close()
    if rank == 0:
        ret.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_nccl():
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (world.bit_length() - 1)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) == "ok"
