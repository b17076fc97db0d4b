set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n '^E  \|^FAILED\|passed\|failed' gpurun_out/r2_pytest_gpu.log | head -40
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_train_c.json 2> gpurun_out/r2_bench_train_c.err
head -c 400 gpurun_out/r2_bench_train_c.json; echo
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_(conv_umma|t1)' -s 77 -c 6 -o gpurun_out/r2_l17 -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
