mkdir -p gpurun_out
for extra in "--strong" "--strong --entries 65536 --steps 50"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --no-cpu-baseline $extra > gpurun_out/strong.json 2> gpurun_out/strong.err
python -c "
import json; d=json.loads(open('gpurun_out/strong.json').read()); print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1)); open('gpurun_out/strong_all.jsonl','a').write(json.dumps(d)+'\n')"
done
