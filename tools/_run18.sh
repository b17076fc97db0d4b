python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden" 2>&1 | grep -v Warning | tail -15
