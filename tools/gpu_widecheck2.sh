mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or entry" 2>&1 | tail -4
for e in 64 128 256; do for t in 1 0; do
echo "E=$e mac_tma=$t $(B200DPF_MAC_TMA=$t timeout 300 python bench.py --entry $e --steps 3 --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["ms_per_step"],2), d["gpu_launches"], round(d["e2e"]["value"],1))')"
done; done 2>&1 | tee gpurun_out/wide2.txt
