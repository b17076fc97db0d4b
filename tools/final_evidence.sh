set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n '^E  \|^FAILED\|passed\|failed' gpurun_out/r2_pytest_gpu.log | head -20
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_train_n1.json 2> gpurun_out/r2_bench_train.err
python bench.py --mesh mano --batch 1024 --steps 10 --warmup 3 --cpu-sample 0 --train-steps 0 > gpurun_out/r2_bench_mano_fwd.json 2> gpurun_out/r2_bench_mano.err
python bench.py --mesh mano --batch 1024 --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_mano_train.json 2>> gpurun_out/r2_bench_mano.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd.csv python tools/ncu_forward.py 2 > gpurun_out/r2_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/ncu_forward.py 2 256 train > gpurun_out/r2_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_(conv_umma|t1)' -s 77 -c 6 -o gpurun_out/r2_l17 -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu2.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke > gpurun_out/r2_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r2_sanitizer_memcheck.log
for f in n1 reference_arm train_n1 mano_fwd mano_train; do head -c 260 gpurun_out/r2_bench_$f.json; echo; done
