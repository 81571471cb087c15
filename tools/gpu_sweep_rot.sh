# experiment: FMA-pipe rotates for ChaCha/Salsa; subtree size for AES
mkdir -p gpurun_out
for m in 0 8 4 2 1 10 12 9; do
  B200DPF_EXTRA_NVCC_FLAGS="-DDPF_FMA_ROT_MASK=$m" python -c "
import sys; sys.path.insert(0,'gpu-dpf_b200'); import build; build.build_lib(force=True)"
  for prf in chacha20 salsa20; do
    echo "mask=$m $prf $(python bench.py --prf $prf --no-cpu-baseline --no-e2e --steps 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])')"
  done
done 2>&1 | tee gpurun_out/sweep_rot.txt
python -c "
import sys; sys.path.insert(0,'gpu-dpf_b200'); import build; build.build_lib(force=True)"
for s in 7 8 9 10; do
  echo "aes s=$s $(python bench.py --subtree-log2 $s --no-cpu-baseline --no-e2e --steps 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])')"
done 2>&1 | tee gpurun_out/sweep_aes_s.txt
