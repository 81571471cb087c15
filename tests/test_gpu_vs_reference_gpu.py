"""GPU kernel against GPU kernel: this engine vs the reference's own dpf_hybrid_kernel,
compiled unmodified for sm_100a into oracle/_ref/ref_dpf_cpp.so (build container only; the
.so travels with the snapshot).  Same keys, same table, results must be bit-identical."""
import numpy as np
import pytest
import torch

import b200dpf
from common import random_table, seeded_keys

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refgpu():
    import refgpu as R
    if not R.available():
        pytest.skip("oracle/_ref/ref_dpf_cpp.so not built (reference tree absent at build time)")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return R


@pytest.mark.parametrize("prf", [0, 1, 2, 3])
def test_same_results_as_reference_kernel(refgpu, prf):
    n, batch = 4096, 70
    table = random_table(n, 16, seed=prf, full_range=True)
    ka, kb, idx = seeded_keys(b200dpf.gen, n, batch, prf, seed=40 + prf)
    ref = refgpu.RefGpuDPF(prf)
    ref.eval_init(torch.from_numpy(table))
    want_a = ref.eval_gpu([torch.from_numpy(k) for k in ka]).numpy()
    want_b = ref.eval_gpu([torch.from_numpy(k) for k in kb]).numpy()
    ref.close()
    ctx = b200dpf.Context(table)
    got_a, got_b = ctx.eval(ka, prf), ctx.eval(kb, prf)
    ctx.close()
    assert np.array_equal(got_a, want_a)
    assert np.array_equal(got_b, want_b)
    assert np.array_equal((want_a.astype(np.uint32) - want_b.astype(np.uint32)).astype(np.int32), table[idx])


def test_entry_size_and_short_batch_like_reference(refgpu):
    n, batch, entry, prf = 1024, 5, 7, 3
    table = random_table(n, entry, seed=9)
    ka, _, _ = seeded_keys(b200dpf.gen, n, batch, prf, seed=9)
    ref = refgpu.RefGpuDPF(prf)
    ref.eval_init(torch.from_numpy(table))
    want = ref.eval_gpu([torch.from_numpy(k) for k in ka]).numpy()
    ref.close()
    ctx = b200dpf.Context(table)
    assert np.array_equal(ctx.eval(ka, prf), want)
    ctx.close()
