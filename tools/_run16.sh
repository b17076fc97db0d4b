set -x
python -m pytest tests/test_gpu_front_back.py -m gpu -q 2>&1 | tail -15
python tools/time_frontback.py 2>&1 | tail -5
