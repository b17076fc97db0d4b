set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n '^E  \|^FAILED\|passed\|failed' gpurun_out/r2_pytest_gpu.log | head -40
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/ncu_forward.py 2 256 train > gpurun_out/r2_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd.csv python tools/ncu_forward.py 2 > gpurun_out/r2_ncu1.log 2>&1
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_train_b.json 2> gpurun_out/r2_bench_train_b.err
head -c 700 gpurun_out/r2_bench_train_b.json
