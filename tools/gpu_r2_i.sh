# full GPU suite on the final code + host-path timings
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2i_pytest.txt 2>&1; tail -12 gpurun_out/r2i_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/gpu_hostpath.py 2>&1 | tee gpurun_out/r2i_hostpath.txt
