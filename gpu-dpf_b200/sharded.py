"""Entry-range sharding of one table across the GPUs of a box (one process per GPU).

The reference is single-GPU.  A DPF evaluation is a sum
over leaves, so disjoint leaf ranges give partial sums that add mod 2^32: rank r
of G = 2^g owns the GGM subtree under the depth-g node with breadth-first index
r (natural indices i with bitrev_g(i mod G) == r), holds only those n/G table
rows, evaluates EVERY key over its subtree, and the [B, E] int32 partials meet in
one NCCL reduce over NVLink (wrapping int32 add is exactly the required
arithmetic).  No other data-path collective exists.

    dist.init_process_group("nccl")            # torchrun, one rank per GPU
    d = ShardedDPF(prf=DPF.PRF_AES128)         # device = LOCAL_RANK
    d.eval_init(table)                         # every rank passes the same table
    out = d.eval_gpu(keys)                     # rank 0: int32 [B, E] CPU tensor; others: None
"""
import os

import torch
import torch.distributed as dist


def shard_indices(n, rank, world):
    """Natural table indices owned by `rank`, in the breadth-first leaf order they are
    stored in on the device: position q of the shard holds index
    bitrev_{d-g}(q) * G + bitrev_g(rank).  (Pure index math; used by the tests.)"""
    d = n.bit_length() - 1
    g = world.bit_length() - 1
    assert 1 << d == n and 1 << g == world and world <= n // 2

    def brev(v, bits):
        r = 0
        for i in range(bits):
            r |= ((v >> i) & 1) << (bits - 1 - i)
        return r

    c = brev(rank, g)
    return [brev(q, d - g) * world + c for q in range(n // world)]


class ShardedDPF(object):
    """dpf.DPF semantics over a process group; rank 0 receives the reduced result."""

    def __init__(self, prf=None, group=None, device=None, partial_fn=None):
        # partial_fn(keys_packed_cpu_int32[B,524]) -> int32 [B,E] partial tensor on this rank's
        # device.  Default: the CUDA engine.  Tests inject a CPU stand-in to exercise the
        # process-group plumbing with the gloo backend.
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world & (self.world - 1):
            raise Exception("number of shards (%d) must be a power of two" % self.world)
        self.device = int(os.environ.get("LOCAL_RANK", self.rank)) if device is None else device
        self._partial_fn = partial_fn
        self._dpf = None
        self.prf = prf
        self.entry_size = None
        self.n = None

    def eval_init(self, table):
        self.n, self.entry_size = table.shape[0], table.shape[1]
        if self._partial_fn is None:
            import dpf
            self._dpf = dpf.DPF(prf=self.prf, device=self.device, shard=(self.rank, self.world))
            self._dpf.eval_init(table)
        return self

    # -- device-resident path (benchmarks, servers that keep keys on the GPU) --
    def eval_gpu_device(self, keys_dev, out_dev=None):
        """Partial on this rank's stream followed by the reduce; returns the device tensor
        (complete on rank 0 once the stream is synchronised)."""
        out_dev = self._dpf.eval_gpu_device(keys_dev, out_dev)
        if self.world > 1:
            dist.reduce(out_dev, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        return out_dev

    # -- host-buffer path (the dpf.DPF.eval_gpu contract) --
    def eval_gpu(self, keys):
        if isinstance(keys, torch.Tensor):
            packed = keys.contiguous()
        else:
            packed = torch.stack(list(keys))
        if self._partial_fn is not None:
            part = self._partial_fn(packed)
            if self.world > 1:
                dist.reduce(part, dst=0, op=dist.ReduceOp.SUM, group=self.group)
            return part.cpu() if self.rank == 0 else None
        dev = torch.device("cuda", self.device)
        keys_dev = packed.to(dev, non_blocking=True)
        out_dev = self.eval_gpu_device(keys_dev)
        if self.rank == 0:
            return out_dev.cpu()
        torch.cuda.current_stream(dev).synchronize()
        return None

    def close(self):
        if self._dpf is not None:
            self._dpf.close()
