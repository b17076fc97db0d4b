python -m pytest tests/test_gpu_front_back.py -m gpu -q -x 2>&1 | grep -v Warning | grep -B5 -A25 "def test_posenet\|^E " | head -80
