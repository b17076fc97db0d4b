mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n 'passed\|failed\|^FAILED' gpurun_out/r2_pytest_gpu.log | head
bash tools/_run13.sh
