mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run() { # label, env..., bench args
  label=$1; shift
  echo "$label $(env "$@" 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["gpu_launches"])')"
}
{
for s in 4 5 6 7 8; do run "aes n=2^20 frontier s=$s" B200DPF_S=$s python bench.py --no-cpu-baseline --no-e2e --steps 5; done
run "aes n=2^20 nofrontier s=8" B200DPF_FRONTIER=0 python bench.py --no-cpu-baseline --no-e2e --steps 5
for s in 3 4 5 6; do run "aes n=2^14 frontier s=$s" B200DPF_S=$s python bench.py --entries 16384 --no-cpu-baseline --no-e2e --steps 20; done
run "aes n=2^14 nofrontier" B200DPF_FRONTIER=0 python bench.py --entries 16384 --no-cpu-baseline --no-e2e --steps 20
for s in 4 5 6; do run "aes n=2^16 frontier s=$s" B200DPF_S=$s python bench.py --entries 65536 --no-cpu-baseline --no-e2e --steps 20; done
run "aes n=2^16 nofrontier" B200DPF_FRONTIER=0 python bench.py --entries 65536 --no-cpu-baseline --no-e2e --steps 20
for s in 5 6 7 8 10; do run "chacha n=2^20 frontier s=$s" B200DPF_S=$s python bench.py --prf chacha20 --no-cpu-baseline --no-e2e --steps 5; done
run "chacha n=2^20 nofrontier" B200DPF_FRONTIER=0 python bench.py --prf chacha20 --no-cpu-baseline --no-e2e --steps 5
run "aes n=2^20 E=128 B=512" python bench.py --entry 128 --no-cpu-baseline --no-e2e --steps 3
run "aes n=2^20 E=32 B=512" python bench.py --entry 32 --no-cpu-baseline --no-e2e --steps 3
run "aes n=2^20 E=64 B=512" python bench.py --entry 64 --no-cpu-baseline --no-e2e --steps 3
} 2>&1 | tee gpurun_out/sweep_frontier.txt
