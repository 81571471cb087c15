mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi_pytest.txt 2>&1; tail -25 gpurun_out/multi_pytest.txt
for r in nccl fused; do
  for extra in "" "--entries 16384 --steps 50"; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --reduce $r --no-cpu-baseline $extra > gpurun_out/multi_${r}.json 2> gpurun_out/multi_${r}.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/multi_${r}.json").read())
    print("$r $extra", d["config"]["parallelism"], d["value"], d["ms_per_step"], d["e2e"]["value"])
    open("gpurun_out/multi_all.jsonl","a").write(json.dumps(d)+"\n")
except Exception as e:
    print("$r $extra FAILED", e); print(open("gpurun_out/multi_${r}.err").read()[-1500:])
PY
  done
done
