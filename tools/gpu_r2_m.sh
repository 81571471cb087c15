timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 2e24 and not config4 and not config5" 2>&1 | tail -40
