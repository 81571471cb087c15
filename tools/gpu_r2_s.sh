timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keygen or matches_oracle or smallest" 2>&1 | tail -5
python - <<'PY'
import sys, time
sys.path.insert(0, "gpu-dpf_b200")
import numpy as np, b200dpf
rng = np.random.RandomState(1)
for n, count in ((1 << 14, 8192), (1 << 20, 8192), (1 << 24, 4096)):
    alphas = rng.randint(0, n, size=count); seeds = rng.bytes(44 * count)
    b200dpf.gen_batch_gpu(alphas[:64], n, seeds[:44 * 64], 3)
    t0 = time.perf_counter(); b200dpf.gen_batch_gpu(alphas, n, seeds, 3); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); b200dpf.gen_batch_secure(alphas, n, seeds, 3); tc = time.perf_counter() - t0
    print("keygen AES n=2^%d, %d key pairs: GPU %.1f ms (host arrays out), CPU all cores %.1f ms" % (n.bit_length() - 1, count, tg * 1e3, tc * 1e3))
PY
