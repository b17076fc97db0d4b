mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_front_back.py -m gpu -q -x 2>&1 | tail -2; done
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n 'passed\|failed\|^FAILED' gpurun_out/r2_pytest_gpu.log | head
