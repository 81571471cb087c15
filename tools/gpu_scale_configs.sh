# BASELINE.json configs 4 and 5 and the small-n case on 8 GPUs, default on 4 GPUs
mkdir -p gpurun_out
run() { # nproc, label, args...
  N=$1; L=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/cfg_$L.json 2> gpurun_out/cfg_$L.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cfg_$L.json").read())
    print("$L", d["config"]["workload"], "|", round(d["value"],1), "DPFs/s", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["value"],1), d["clocks"])
    open("gpurun_out/cfg_all.jsonl","a").write(json.dumps(d)+"\n")
except Exception as e:
    print("$L FAILED", e); print(open("gpurun_out/cfg_$L.err").read()[-1200:])
PY
}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
run 8 c4_n2e24_salsa_b4096 --entries 16777216 --prf salsa20 --steps 3
run 8 c5_n2e20_e128_aes_b8192 --entry 128 --batch-per-gpu 1024 --steps 3
run 8 small_n2e14_aes_b4096 --entries 16384 --steps 50
run 8 small_n2e16_aes_b4096 --entries 65536 --steps 30
run 4 default_4gpu
