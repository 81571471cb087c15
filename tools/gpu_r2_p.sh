mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2_final_pytest.txt 2>&1; tail -10 gpurun_out/r2_final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
