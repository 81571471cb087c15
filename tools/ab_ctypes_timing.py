"""Time b200dpf_eval_device of an arbitrary build of libb200dpf.so (path in argv[1]) through bare ctypes:
for A/B runs against older builds whose symbol set differs from the current Python bindings."""
import ctypes as C, sys
sys.path.insert(0, "gpu-dpf_b200"); sys.path.insert(0, "tests")
import numpy as np, torch
from common import random_table
import b200dpf                       # current build, used for key generation only

lib = C.CDLL(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
batch, prf = 512, int(sys.argv[3]) if len(sys.argv) > 3 else 3
table = random_table(n, 16, seed=1)
ka, _ = b200dpf.gen_batch(np.arange(batch) * 7 % n, n, np.arange(batch) + 7, prf)
ctx = C.c_void_p()
lib.b200dpf_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
lib.b200dpf_eval_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
assert lib.b200dpf_create(C.byref(ctx), table.ctypes.data_as(C.c_void_p), n, 16, 0, 0, 1) == 0
kd = torch.from_numpy(ka).cuda(); out = torch.empty((batch, 16), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
run = lambda: lib.b200dpf_eval_device(ctx, C.c_void_p(kd.data_ptr()), batch, prf, C.c_void_p(out.data_ptr()), C.c_void_p(st))
for _ in range(5): assert run() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("%s n=%d prf=%d B=512: %.4f ms/eval  checksum %d" % (sys.argv[1].split("/")[-2], n, prf, e0.elapsed_time(e1) / 20, int(out.to(torch.int64).sum())), flush=True)
