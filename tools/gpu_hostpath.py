import sys, time, os
sys.path.insert(0,'gpu-dpf_b200'); sys.path.insert(0,'tests')
import numpy as np, torch, b200dpf, dpf, dpf_cpp
from common import random_table
n=1<<14
d=dpf.DPF(prf=3); d.eval_init(torch.from_numpy(random_table(n,16,seed=1)))
ka,_=b200dpf.gen_batch(np.arange(512)%n, n, np.arange(512)+7, 3)
lst=[torch.from_numpy(k.copy()) for k in ka]
packed=torch.from_numpy(ka.copy()); pinned=packed.pin_memory()
def t(fn, reps=300):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
kd=pinned.cuda(); out=torch.empty((512,16),dtype=torch.int32,device='cuda')
print("device only            %.3f ms" % t(lambda: d.eval_gpu_device(kd,out)))
print("packed pinned          %.3f ms" % t(lambda: dpf_cpp.eval_gpu_packed(pinned,d.buffers,3)))
print("packed pageable        %.3f ms" % t(lambda: dpf_cpp.eval_gpu_packed(packed,d.buffers,3)))
print("list (eval_gpu_list)   %.3f ms" % t(lambda: dpf_cpp.eval_gpu_list(lst,d.buffers,3)))
print("list reference-style   %.3f ms" % t(lambda: dpf_cpp.eval_gpu(lst,d.buffers,n,3)))
print("DPF.eval_gpu(list)     %.3f ms" % t(lambda: d.eval_gpu(lst)))
print("torch.stack only       %.3f ms" % t(lambda: torch.stack(lst)))
compact=d.pack_keys(lst)
print("compact keys (%d B/key) %.3f ms" % (compact.shape[1], t(lambda: d.eval_gpu_compact(compact))))
