mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/dbg.out 2> gpurun_out/dbg.err; echo rc=$?
echo STDOUT; cat gpurun_out/dbg.out | cut -c1-400; echo STDERR; tail -30 gpurun_out/dbg.err
