set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -60 gpurun_out/r2_pytest_gpu.log
