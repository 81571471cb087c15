"""Driver for the reference's own GPU extension (oracle/_ref/ref_dpf_cpp.so) -- TEST
INFRASTRUCTURE ONLY.  Replays what the reference's dpf.py does around its extension
(dpf.py:88-131: pad the table to 16 columns, pad every batch to 512 keys with the last
key, slice the result) so the reference KERNEL can be run without the reference tree."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "ref_dpf_cpp.so")


def available():
    return os.path.exists(SO)


def module():
    d = os.path.dirname(SO)
    if d not in sys.path:
        sys.path.insert(0, d)
    import ref_dpf_cpp
    return ref_dpf_cpp


class RefGpuDPF:
    def __init__(self, prf):
        self.m = module()
        self.prf = prf
        self.buffers = None

    def eval_init(self, table):
        table = torch.as_tensor(table).to(torch.int32)
        self.n, self.e = table.shape
        assert self.e <= self.m.ENTRY_SIZE and self.n >= 128
        padded = torch.nn.functional.pad(table, (0, self.m.ENTRY_SIZE - self.e, 0, 0))
        if self.buffers is not None:
            self.m.eval_free(self.buffers)
        self.buffers = self.m.eval_init(padded)      # n*16 .item() calls on the host: slow for big n

    def eval_gpu(self, keys):
        """keys: list of int32[524] CPU tensors (any length) -> int32 [len, e] CPU tensor."""
        bs = self.m.BATCH_SIZE
        outs = []
        for i in range(0, len(keys), bs):
            cur = list(keys[i:i + bs])
            cur = cur + [cur[-1]] * (bs - len(cur))
            outs.append(self.m.eval_gpu(cur, self.buffers, self.n, self.prf)[:, :self.e])
        return torch.cat(outs)[:len(keys)]

    def close(self):
        if self.buffers is not None:
            self.m.eval_free(self.buffers)
            self.buffers = None
